"""file_parsing mirror: `wav.parse`, `aiff.parse`, `AudioFile`, batch decode.

Same names, argument meaning and error behaviour as the reference
(blast/src/file_parsing/{wav,aiff,decode_helpers}.rs); the sample loops run on the GPU through
libblast_cuda (blast_pcm_decode_batch / blast_pcm_plan_*), the ~20-field header walks through
blast_wav_probe / blast_aiff_probe.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .context import Context, DevBuf
from .errors import Io, check


@dataclass
class AudioFile:
    """decode_helpers.rs:17-38"""
    file_name: str
    format: str
    sample_rate: int
    num_channels: int
    bits_per_sample: int
    samples: np.ndarray          # int16, interleaved (host) — Vec<i16>


@dataclass
class DeviceTrack:
    """AudioFile whose samples stay resident in HBM (their only consumer is the render kernel)."""
    file_name: str
    format: str
    sample_rate: int
    num_channels: int
    bits_per_sample: int
    n_samples: int
    buf: DevBuf                  # int16 words on the device

    def to_host(self) -> np.ndarray:
        return self.buf.download(np.int16, self.n_samples)


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def probe(kind: str, image) -> _lib.PcmDesc:
    """Header walk of one in-memory file image; raises the DecodeError the reference would return."""
    L = _lib.load()
    a = _as_u8(image)
    d = _lib.PcmDesc()
    fn = L.blast_wav_probe if kind == "wav" else L.blast_aiff_probe
    check(fn(a.ctypes.data if a.size else None, a.size, C.byref(d)))
    return d


def file_name(path: str) -> str:
    L = _lib.load()
    buf = C.create_string_buffer(max(16, len(path.encode()) + 1))
    check(L.blast_file_name(path.encode(), buf, len(buf)))
    return buf.value.decode()


def decode_batch(ctx: Context, images, descs, *, to_host=True, keep_on_device=False):
    """blast_pcm_decode_batch over host file images.

    Returns (host_arrays | None, device_bufs | None); host arrays are int16 numpy arrays.
    """
    L = ctx.lib
    n = len(images)
    arrs = [_as_u8(im) for im in images]
    files = (C.c_void_p * max(1, n))(*[a.ctypes.data if a.size else None for a in arrs])
    lens = (C.c_size_t * max(1, n))(*[a.size for a in arrs])
    dd = (_lib.PcmDesc * max(1, n))(*descs)
    outs = [np.empty(L.blast_pcm_out_len(C.byref(d)), dtype=np.int16) for d in descs] if to_host else None
    host_out = (C.c_void_p * max(1, n))(*[o.ctypes.data if o.size else None for o in outs]) if to_host else None
    dev = [ctx.alloc(max(2, 2 * L.blast_pcm_out_len(C.byref(d)))) for d in descs] if keep_on_device else None
    d_out = (C.c_void_p * max(1, n))(*[b.ptr for b in dev]) if keep_on_device else None
    check(L.blast_pcm_decode_batch(ctx.h, n, files, lens, dd, host_out, d_out))
    return outs, dev


class _Format:
    def __init__(self, kind: str, fmt: str):
        self.kind, self.fmt = kind, fmt

    def parse_bytes(self, image, path: str = "assets/memory." + "bin", ctx: Context | None = None) -> AudioFile:
        own = ctx is None
        ctx = ctx or Context()
        try:
            d = probe(self.kind, image)
            outs, _ = decode_batch(ctx, [image], [d])
            name = file_name(path)                      # checked after decoding, like the reference
            return AudioFile(name, self.fmt, d.sample_rate, d.num_channels, d.bits_per_sample, outs[0])
        finally:
            if own:
                ctx.close()

    def parse(self, path: str, ctx: Context | None = None) -> AudioFile:
        """`pub fn parse(path: &str) -> DecodeResult<AudioFile>` (wav.rs:69 / aiff.rs:99)."""
        try:
            with open(path, "rb") as f:
                image = f.read()
        except OSError as e:                            # DecodeError::Io via From<io::Error>
            raise Io(_lib.ERR_IO, str(e)) from e
        return self.parse_bytes(image, path, ctx)

    def parse_to_device(self, ctx: Context, image, path: str) -> DeviceTrack:
        d = probe(self.kind, image)
        _, dev = decode_batch(ctx, [image], [d], to_host=False, keep_on_device=True)
        name = file_name(path)
        return DeviceTrack(name, self.fmt, d.sample_rate, d.num_channels, d.bits_per_sample,
                           ctx.lib.blast_pcm_out_len(C.byref(d)), dev[0])


wav = _Format("wav", "wav")
aiff = _Format("aiff", "aiff")


def decode_jobs_dev(ctx: Context, jobs):
    """blast_pcm_decode_dev over device-resident payloads: jobs = [(d_src, d_dst, n_words, big_endian)]."""
    arr = (_lib.PcmJob * max(1, len(jobs)))(*[_lib.PcmJob(s, d, n, int(be), 0) for s, d, n, be in jobs])
    check(ctx.lib.blast_pcm_decode_dev(ctx.h, arr, len(jobs)))


class PcmPlan:
    """blast_pcm_plan_*: a batch whose launch is the only thing inside the timed region."""

    def __init__(self, ctx: Context, jobs):
        self.ctx = ctx
        arr = (_lib.PcmJob * max(1, len(jobs)))(*[_lib.PcmJob(s, d, n, int(be), 0) for s, d, n, be in jobs])
        p = C.c_void_p()
        check(ctx.lib.blast_pcm_plan_create(ctx.h, arr, len(jobs), C.byref(p)))
        self.h = p.value

    @property
    def words(self) -> int:
        return self.ctx.lib.blast_pcm_plan_words(self.h)

    def run(self):
        check(self.ctx.lib.blast_pcm_plan_run_dev(self.ctx.h, self.h))

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.blast_pcm_plan_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pcm24_unpack_dev(ctx: Context, jobs):
    """Extension: jobs = [(d_src, d_dst, n_samples, big_endian, out_kind)]"""
    arr = (_lib.Pcm24Job * max(1, len(jobs)))(*[_lib.Pcm24Job(s, d, n, int(be), k) for s, d, n, be, k in jobs])
    check(ctx.lib.blast_pcm24_unpack_dev(ctx.h, arr, len(jobs)))
