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


def asset_consensus(descs) -> tuple[int, int]:
    """(mutual sample rate, channel count) of an asset set: main.rs:79-120"""
    arr = (_lib.PcmDesc * max(1, len(descs)))(*descs)
    rate, ch = C.c_uint32(), C.c_uint32()
    check(_lib.load().blast_asset_consensus(arr, len(descs), C.byref(rate), C.byref(ch)))
    return rate.value, ch.value


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


_default_ctx: Context | None = None


def default_context() -> Context:
    """The process-wide context of the parse() drop-ins (the reference calls them in a loop over an asset directory,
    main.rs:18-89): created on first use and kept, so that loop pays for streams and staging slabs once, not per file."""
    global _default_ctx
    if _default_ctx is None or not _default_ctx.h:
        _default_ctx = Context()
    return _default_ctx


class _Format:
    def __init__(self, kind: str, fmt: str):
        self.kind, self.fmt = kind, fmt

    def parse_bytes(self, image, path: str = "assets/memory." + "bin", ctx: Context | None = None) -> AudioFile:
        ctx = ctx or default_context()
        d = probe(self.kind, image)
        outs, _ = decode_batch(ctx, [image], [d])
        name = file_name(path)                      # checked after decoding, like the reference
        return AudioFile(name, self.fmt, d.sample_rate, d.num_channels, d.bits_per_sample, outs[0])

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


# ---------------------------------------------------------------------------------- mpeg
def mpeg_header_info(header: int) -> _lib.MpegHeader:
    """parse_header + Header::format + compute_frame_len for one header word (mpeg.rs:154-234, 367-496)"""
    o = _lib.MpegHeader()
    check(_lib.load().blast_mpeg_header_info(header & 0xFFFFFFFF, C.byref(o)))
    return o


class _Mpeg:
    """file_parsing::mpeg (blast/src/file_parsing/mpeg.rs)"""

    @staticmethod
    def scan_dev(ctx: Context, d_bytes: int, length: int, cap: int | None = None):
        """greedy sync scan of device-resident bytes -> (positions uint64, headers uint32) in file order"""
        cap = length // 32 + 4096 if cap is None else cap
        d_pos, d_hdr = ctx.alloc(max(16, 8 * cap)), ctx.alloc(max(16, 4 * cap))
        n = C.c_uint64()
        check(ctx.lib.blast_mpeg_scan_dev(ctx.h, d_bytes, length, d_pos.ptr, d_hdr.ptr, cap, C.byref(n)))
        return d_pos.download(np.uint64, n.value), d_hdr.download(np.uint32, n.value)

    @staticmethod
    def index_dev(ctx: Context, d_bytes: int, length: int, reference_compat: bool = True):
        """-> dict(offsets uint64 (frames[*].file_pos, sorted), ref_header, n_candidates)"""
        n, ncand, ref = C.c_uint64(), C.c_uint64(), C.c_uint32()
        check(ctx.lib.blast_mpeg_index_dev(ctx.h, d_bytes, length, int(reference_compat), None, 0, C.byref(n),
                                           C.byref(ref), C.byref(ncand)))
        d_off = ctx.alloc(max(16, 8 * n.value))
        check(ctx.lib.blast_mpeg_index_dev(ctx.h, d_bytes, length, int(reference_compat), d_off.ptr, n.value,
                                           C.byref(n), C.byref(ref), C.byref(ncand)))
        return dict(offsets=d_off.download(np.uint64, n.value), ref_header=ref.value, n_candidates=ncand.value,
                    d_offsets=d_off)

    @staticmethod
    def shard_walk_dev(ctx: Context, d_bytes: int, own_len: int, halo_len: int):
        """phase 1 of the sharded scan -> (exit_state[4], count[4]) of this byte range"""
        agg = _lib.MpegShardAgg()
        check(ctx.lib.blast_mpeg_shard_walk_dev(ctx.h, d_bytes, own_len, halo_len, C.byref(agg)))
        return [int(x) for x in agg.exit_state], [int(x) for x in agg.count]

    @staticmethod
    def shard_emit_dev(ctx: Context, d_bytes: int, own_len: int, halo_len: int, entry_state: int, pos_offset: int,
                       count: int):
        """phase 2 -> device buffers (positions uint64 with pos_offset added, headers uint32) of `count` candidates"""
        d_pos, d_hdr = ctx.alloc(max(16, 8 * count)), ctx.alloc(max(16, 4 * count))
        n = C.c_uint64()
        check(ctx.lib.blast_mpeg_shard_emit_dev(ctx.h, d_bytes, own_len, halo_len, entry_state, pos_offset, d_pos.ptr,
                                                d_hdr.ptr, count, C.byref(n)))
        assert n.value == count, (n.value, count)
        return d_pos, d_hdr

    @staticmethod
    def gather_dev(ctx: Context, d_bytes: int, length: int, d_offsets: int, n_offsets: int) -> np.ndarray:
        plen = C.c_uint64()
        check(ctx.lib.blast_mpeg_gather_dev(ctx.h, d_bytes, length, d_offsets, n_offsets, None, 0, C.byref(plen)))
        d_pay = ctx.alloc(max(16, plen.value))
        check(ctx.lib.blast_mpeg_gather_dev(ctx.h, d_bytes, length, d_offsets, n_offsets, d_pay.ptr, plen.value,
                                            C.byref(plen)))
        return d_pay.download(np.uint8, plen.value)

    @staticmethod
    def parse_bytes(image, ctx: Context | None = None, reference_compat: bool = True, want_payload: bool = True):
        """blast_mpeg_parse on a host buffer -> dict(offsets, ref_header, n_candidates, payload)"""
        ctx = ctx or default_context()
        a = _as_u8(image)
        n, ncand, ref, plen = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_uint64()
        ptr = a.ctypes.data if a.size else None
        check(ctx.lib.blast_mpeg_parse(ctx.h, ptr, a.size, int(reference_compat), None, 0, C.byref(n), C.byref(ref),
                                       C.byref(ncand), None, 0, C.byref(plen) if want_payload else None))
        offs = np.empty(n.value, dtype=np.uint64)
        pay = np.empty(plen.value if want_payload else 0, dtype=np.uint8)
        check(ctx.lib.blast_mpeg_parse(ctx.h, ptr, a.size, int(reference_compat), offs.ctypes.data, offs.size,
                                       C.byref(n), C.byref(ref), C.byref(ncand),
                                       pay.ctypes.data if want_payload and pay.size else None, pay.size,
                                       C.byref(plen) if want_payload else None))
        return dict(offsets=offs, ref_header=ref.value, n_candidates=ncand.value, payload=pay)

    def parse(self, path: str, ctx: Context | None = None) -> np.ndarray:
        """`pub fn parse(path: &str) -> DecodeResult<Vec<u8>>` (mpeg.rs:7): the concatenated frame payloads"""
        try:
            with open(path, "rb") as f:
                image = f.read()
        except OSError as e:
            raise Io(_lib.ERR_IO, str(e)) from e
        return self.parse_bytes(image, ctx)["payload"]


mpeg = _Mpeg()
