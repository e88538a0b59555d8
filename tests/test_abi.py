"""CPU-side checks of the C ABI: the library loads, exports every symbol the header declares,
refuses to run without a GPU, and its host logic (header walks, file-name rule) matches the oracle."""
import os
import re
import struct

import numpy as np
import pytest

import audio_decoder_b200 as blast
from audio_decoder_b200 import _lib, file_parsing as fp
import oracle
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(blast_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    declared = _declared_symbols()
    assert len(declared) > 20
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    # and the ctypes table covers the whole header (no unbound entry points)
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.blast_abi_version() == 3


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(blast.BlastError) as e:
        blast.Context()
    assert e.value.code == _lib.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def _same_probe(kind, image):
    """product probe == oracle probe, including the error variant"""
    o_fn = oracle.wav_probe if kind == "wav" else oracle.aiff_probe
    try:
        o = o_fn(image)
        o_err = None
    except oracle.OracleError as e:
        o, o_err = None, e.code
    try:
        p = fp.probe(kind, image)
        p_err = None
    except blast.BlastError as e:
        p, p_err = None, e.code
    assert o_err == p_err, (kind, o_err, p_err)
    if o is not None:
        for f, _ in _lib.PcmDesc._fields_:
            assert getattr(o, f) == getattr(p, f), f
    return p, p_err


def test_probe_canonical_headers():
    p, _ = _same_probe("wav", synth.wav_image(1, 4000))
    assert (p.data_off, p.data_len, p.sample_rate, p.num_channels, p.bits_per_sample) == (44, 4000, 44100, 2, 16)
    p, _ = _same_probe("aiff", synth.aiff_image(2, 6000))
    assert (p.data_off, p.data_len, p.sample_rate, p.num_channels, p.bits_per_sample, p.big_endian) == \
        (54, 6000, 48000, 2, 24, 1)
    assert _lib.load().blast_pcm_out_len(p) == 3000


def test_probe_truncations_and_quirks():
    w = synth.wav_image(3, 101)            # odd payload, no byte after it -> EOF
    _, err = _same_probe("wav", w)
    assert err == _lib.ERR_UNEXPECTED_EOF
    _same_probe("wav", np.concatenate([w, np.zeros(1, np.uint8)]))
    a = synth.aiff_image(4, 64)
    for cut in range(0, len(a) + 1):
        _same_probe("aiff", a[:cut])
    w = synth.wav_image(5, 64)
    for cut in range(0, len(w) + 1):
        _same_probe("wav", w[:cut])
    # extensible fmt: +91 skip
    ext = struct.pack("<HHIH", 22, 16, 3, 1) + bytes(91)
    img = (b"RIFF" + struct.pack("<I", 0) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 40, 0xFFFE, 2, 48000, 0, 4, 16) +
           ext + b"data" + struct.pack("<I", 32) + bytes(32))
    p, err = _same_probe("wav", img)
    assert err is None and p.data_off == 145
    for cut in range(30, len(img)):
        _same_probe("wav", img[:cut])
    # bad tag, bad COMM size, SSND size < 8
    bad = bytearray(synth.wav_image(6, 16)); bad[20] = 2
    assert _same_probe("wav", bytes(bad))[1] == _lib.ERR_UNSUPPORTED_FORMAT
    bad = bytearray(synth.aiff_image(7, 16)); bad[19] = 20
    assert _same_probe("aiff", bytes(bad))[1] == _lib.ERR_INVALID_DATA
    bad = bytearray(synth.aiff_image(8, 16)); bad[42:46] = struct.pack(">I", 4)
    assert _same_probe("aiff", bytes(bad))[1] == _lib.ERR_UNEXPECTED_EOF


def test_probe_fuzz_matches_oracle():
    rng = np.random.default_rng(99)
    base_w = synth.wav_image(9, 40)
    base_a = synth.aiff_image(10, 40)
    for _ in range(1500):
        for kind, base in (("wav", base_w), ("aiff", base_a)):
            img = base.copy()
            k = int(rng.integers(1, 5))
            pos = rng.integers(0, 60, size=k)
            img[pos] = rng.integers(0, 256, size=k, dtype=np.uint8)
            cut = int(rng.integers(0, len(img) + 1)) if rng.random() < 0.3 else len(img)
            _same_probe(kind, img[:cut])


def test_aiff_rate_conversion_matches_oracle():
    rng = np.random.default_rng(5)
    L = _lib.load()
    for _ in range(2000):
        ext = rng.integers(0, 256, size=10, dtype=np.uint8)
        if rng.random() < 0.5:            # plausible exponents around 2^0 .. 2^40
            e = 16383 + int(rng.integers(-4, 40))
            ext[0] = (e >> 8) & 0x7F
            ext[1] = e & 0xFF
        img = synth.aiff_image(1, 8)
        img[28:38] = ext
        _same_probe("aiff", img)
    assert L is not None


def test_file_name_rule():
    for path in ["blast/assets/fairies.wav", "a/b.c/d.e.aif", "x/.wav", "/abs/path/t.aiff"]:
        assert fp.file_name(path) == oracle.file_name(path)
    for bad in ["fairies.wav", "noext", ".wav", "dir/name.", "a.b/c"]:
        with pytest.raises(blast.InvalidData) as e:
            fp.file_name(bad)
        with pytest.raises(oracle.OracleError) as oe:
            oracle.file_name(bad)
        assert oe.value.code == oracle.INVALID_DATA
        assert oracle.lib().orc_last_error().decode() in str(e.value)


def test_parse_missing_file_is_io_error():
    with pytest.raises(blast.Io):
        fp.wav.parse("/nonexistent/dir/file.wav")


def test_x128p_seed_and_host_jump_match_oracle():
    from audio_decoder_b200 import blast_rand as br
    for seed in (0, 1, 42, 0xDEADBEEFCAFEBABE, 2**64 - 1):
        assert br.seed_state(seed) == oracle.Rng(seed).state
    base = br.seed_state(42)
    for n in (0, 1, 2, 3, 65536, 1_000_000, 12_345_678):
        g = oracle.Rng(42)
        g.discard(n)
        assert br.advance(base, n) == g.state, n
    assert br.advance(base, 65536) == (0x7642b3b57ffb2a57, 0xc7c3ecc7c44024b3)
    # composition: jump(a) then jump(b) == jump(a + b), for jumps far beyond what can be walked
    assert br.advance(br.advance(base, 2**40), 2**40) == br.advance(base, 2**41)
    # against the Python GF(2) model
    import pyref
    T = pyref.transition_columns()
    st = pyref.mat_vec(pyref.mat_pow(T, 2**40 + 7), base[0] | (base[1] << 64))
    assert br.advance(base, 2**40 + 7) == (st & pyref.M64, st >> 64)


def test_mpeg_header_info_matches_oracle_exhaustively():
    """all 2^21 header values with the 11 sync bits set: host classifier == oracle parse_header/format/frame_len"""
    L = _lib.load()
    o = _lib.MpegHeader()
    import ctypes as C
    ol = oracle.lib()
    oh = oracle.MpegHeader()
    for idx in list(range(0, 1 << 21, 7)) + list(range(0x1B9000, 0x1B9400)):
        h = 0xFFE00000 | idx
        L.blast_mpeg_header_info(h, C.byref(o))
        ol.orc_mpeg_parse_header(h, C.byref(oh))
        assert o.ok == oh.ok, hex(h)
        if not oh.ok:
            assert o.status == oh.err, hex(h)
            continue
        assert (o.layer, o.is_protected, o.padded, o.channel_mode, o.bitrate, o.sample_rate, o.skip) == \
            (oh.layer, int(oh.not_protected == 0), oh.padded, oh.channel_mode, oh.bitrate, oh.sr, oh.skip), hex(h)
        assert o.frame_len_ok == oh.frame_len_ok, hex(h)
        if oh.frame_len_ok:
            assert o.payload_len == oh.payload_len, hex(h)


def test_struct_layouts_match_the_header(tmp_path):
    """every ABI struct: sizeof and field offsets of the ctypes mirror == what gcc computes from include/blast_cuda.h"""
    import ctypes as C
    import subprocess
    structs = {"blast_pcm_desc": _lib.PcmDesc, "blast_pcm_job": _lib.PcmJob, "blast_pcm24_job": _lib.Pcm24Job,
               "blast_track": _lib.Track, "blast_voice": _lib.Voice, "blast_x128p": _lib.X128PState,
               "blast_mpeg_header": _lib.MpegHeader, "blast_tempo_repr": _lib.TempoRepr, "blast_command": _lib.Command,
               "blast_timed_command": _lib.TimedCommand, "blast_voice_state": _lib.VoiceState, "blast_mpeg_shard_agg": _lib.MpegShardAgg}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "blast_cuda.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_asset_consensus():
    """main.rs:79-120: most frequent rate, largest channel count; 44100 / 2 without assets"""
    def d(rate, ch):
        return _lib.PcmDesc(rate, ch, 16, 0, 44, 100)
    assert fp.asset_consensus([]) == (44100, 2)
    assert fp.asset_consensus([d(48000, 1)]) == (48000, 1)
    assert fp.asset_consensus([d(48000, 2), d(44100, 1), d(44100, 6), d(22050, 2)]) == (44100, 6)
    assert fp.asset_consensus([d(48000, 2), d(44100, 2)]) == (44100, 2)          # tie: smallest (reference: HashMap order)


def test_rust_sys_crate_names_exist_in_the_header():
    """rust/blast-cuda-sys cannot be compiled here (no rustc): at least every symbol it binds must be declared by the
    header, and the entry points a BLAST maintainer needs must be bound"""
    rs = open(os.path.join(ROOT, "rust", "blast-cuda-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (blast_[a-z0-9_]+)\s*\(", rs))
    declared = _declared_symbols()
    assert bound and bound <= declared, sorted(bound - declared)
    for need in ("blast_wav_probe", "blast_aiff_probe", "blast_pcm_decode_batch", "blast_file_name", "blast_conductor_apply",
                 "blast_conductor_coordinate", "blast_render", "blast_x128p_seed", "blast_mpeg_parse", "blast_asset_consensus"):
        assert need in bound, need


def test_rust_patch_is_self_consistent():
    """rust/blast_patch cannot be compiled here: every `sys::` item it uses must be bound (or defined) by the sys crate,
    every `crate::file_parsing::` item must be a `pub` item of the file_parsing patch, and what engine.rs needs of X128P
    must be added by the blast_rand patch."""
    d = os.path.join(ROOT, "rust", "blast_patch")
    sys_rs = open(os.path.join(ROOT, "rust", "blast-cuda-sys", "src", "lib.rs")).read()
    sys_items = set(re.findall(r"pub (?:fn|const|struct) ([A-Za-z0-9_]+)", sys_rs))
    fp_rs = open(os.path.join(d, "file_parsing.rs")).read()
    fp_pub = set(re.findall(r"pub (?:fn|struct|mod) ([A-Za-z0-9_]+)", fp_rs))
    for fn in ("engine.rs", "file_parsing.rs"):
        text = re.sub(r"//.*", "", open(os.path.join(d, fn)).read())
        used = set(re.findall(r"\bsys::([A-Za-z0-9_]+)", text))
        assert used and used <= sys_items, (fn, sorted(used - sys_items))
        for item in re.findall(r"crate::file_parsing::([A-Za-z0-9_]+)", text):
            assert item in fp_pub | {"decode_helpers"}, (fn, item)
    eng = re.sub(r"//.*", "", open(os.path.join(d, "engine.rs")).read())
    assert "DeviceTrack" in fp_pub and "use crate::file_parsing::DeviceTrack" in eng
    if ".state()" in eng:
        assert re.search(r"pub fn state\(&self\)", open(os.path.join(d, "blast_rand.rs")).read())
