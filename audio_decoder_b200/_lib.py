"""ctypes loader for libblast_cuda.so (the C ABI declared in include/blast_cuda.h).

There is no fallback: if the shared library is missing this module raises ImportError telling
the user to build it (python -c "import __graft_entry__ as g; g.build()" or make -C
audio_decoder_b200/csrc).  No compute ever happens in Python.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BLAST_CUDA_LIB: development only (A/B timing of two builds of the same ABI on one box)
SO_PATH = os.environ.get("BLAST_CUDA_LIB") or os.path.join(_HERE, "libblast_cuda.so")

OK = 0
ERR_IO, ERR_UNSUPPORTED_FORMAT, ERR_UNEXPECTED_EOF, ERR_INVALID_DATA, ERR_REF_PANIC = 1, 2, 3, 4, 5
ERR_CUDA, ERR_ARG, ERR_NO_DEVICE, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_TIMEOUT = 100, 101, 102, 103, 104, 105


class PcmDesc(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("num_channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("big_endian", C.c_uint32), ("data_off", C.c_uint64), ("data_len", C.c_uint64)]


class PcmJob(C.Structure):
    _fields_ = [("d_src", C.c_void_p), ("d_dst", C.c_void_p), ("n_words", C.c_uint64), ("big_endian", C.c_uint32),
                ("reserved", C.c_uint32)]


class Pcm24Job(C.Structure):
    _fields_ = [("d_src", C.c_void_p), ("d_dst", C.c_void_p), ("n_samples", C.c_uint64), ("big_endian", C.c_uint32),
                ("out_kind", C.c_uint32)]


class Track(C.Structure):
    _fields_ = [("d_samples", C.c_void_p), ("n_samples", C.c_uint64), ("num_channels", C.c_uint32),
                ("sample_rate", C.c_uint32)]


class Voice(C.Structure):
    _fields_ = [("track", C.c_uint32), ("active", C.c_uint32), ("position", C.c_float), ("velocity", C.c_float),
                ("gain", C.c_float), ("reserved", C.c_uint32)]


class X128PState(C.Structure):
    _fields_ = [("s0", C.c_uint64), ("s1", C.c_uint64)]


class MpegHeader(C.Structure):
    _fields_ = [("ok", C.c_uint32), ("status", C.c_uint32), ("version_id", C.c_uint32), ("layer", C.c_uint32),
                ("is_protected", C.c_uint32), ("padded", C.c_uint32), ("channel_mode", C.c_uint32),
                ("bitrate", C.c_uint32), ("sample_rate", C.c_double), ("frame_len_ok", C.c_uint32),
                ("payload_len", C.c_uint32), ("skip", C.c_uint32), ("reserved", C.c_uint32)]


class TempoRepr(C.Structure):
    _fields_ = [("idx", C.c_uint64), ("owned", C.c_uint32), ("mode", C.c_uint32), ("unit", C.c_uint32),
                ("interval", C.c_float)]


class Command(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("idx_kind", C.c_uint32), ("idx", C.c_uint64), ("val", C.c_float),
                ("reserved", C.c_uint32), ("tempo", TempoRepr),
                ("n_members", C.c_uint32), ("reserved2", C.c_uint32), ("member_voice", C.POINTER(C.c_uint64)),
                ("member_update_tempo", C.POINTER(C.c_uint8)), ("member_n_procs", C.POINTER(C.c_uint32)),
                ("member_proc_ids", C.POINTER(C.c_uint64)),
                ("period", C.c_uint64), ("n_steps", C.c_uint32), ("reserved3", C.c_uint32),
                ("steps", C.POINTER(C.c_float)), ("chance", C.POINTER(C.c_float)),
                ("rng_s0", C.c_uint64), ("rng_s1", C.c_uint64)]


class TimedCommand(C.Structure):
    _fields_ = [("frame", C.c_uint64), ("cmd", Command)]


class VoiceState(C.Structure):
    _fields_ = [("active", C.c_uint32), ("position", C.c_float), ("velocity", C.c_float), ("gain", C.c_float),
                ("end", C.c_uint64), ("channels", C.c_uint32), ("tempo_current", C.c_uint32),
                ("tempo_active", C.c_uint32), ("n_processes", C.c_uint32)]


class MpegShardAgg(C.Structure):
    _fields_ = [("exit_state", C.c_uint32 * 4), ("count", C.c_uint64 * 4)]


# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
_vp, _u64, _u32, _sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t
SIGNATURES = {
    "blast_abi_version": (C.c_int, []),
    "blast_last_error": (C.c_char_p, []),
    "blast_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "blast_ctx_destroy": (None, [_vp]),
    "blast_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "blast_ctx_stream": (_vp, [_vp]),
    "blast_ctx_sync": (C.c_int, [_vp]),
    "blast_ctx_device": (C.c_int, [_vp]),
    "blast_ctx_sm_count": (C.c_int, [_vp]),
    "blast_ctx_launch_count": (_u64, [_vp]),
    "blast_ctx_trim": (C.c_int, [_vp]),
    "blast_dev_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "blast_dev_free": (C.c_int, [_vp, _vp]),
    "blast_host_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "blast_host_free": (C.c_int, [_vp, _vp]),
    "blast_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, _sz]),
    "blast_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, _sz]),
    "blast_memcpy_d2d": (C.c_int, [_vp, _vp, _vp, _sz]),
    "blast_memset_dev": (C.c_int, [_vp, _vp, C.c_int, _sz]),
    "blast_event_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "blast_event_destroy": (None, [_vp]),
    "blast_event_record": (C.c_int, [_vp, _vp]),
    "blast_event_elapsed_ms": (C.c_int, [_vp, _vp, C.POINTER(C.c_float)]),
    "blast_wav_probe": (C.c_int, [_vp, _sz, C.POINTER(PcmDesc)]),
    "blast_aiff_probe": (C.c_int, [_vp, _sz, C.POINTER(PcmDesc)]),
    "blast_pcm_out_len": (_sz, [C.POINTER(PcmDesc)]),
    "blast_file_name": (C.c_int, [C.c_char_p, C.c_char_p, _sz]),
    "blast_asset_consensus": (C.c_int, [C.POINTER(PcmDesc), _u32, C.POINTER(_u32), C.POINTER(_u32)]),
    "blast_pcm_plan_create": (C.c_int, [_vp, C.POINTER(PcmJob), _u32, C.POINTER(_vp)]),
    "blast_pcm_plan_run_dev": (C.c_int, [_vp, _vp]),
    "blast_pcm_plan_destroy": (None, [_vp, _vp]),
    "blast_pcm_plan_words": (_u64, [_vp]),
    "blast_pcm_decode_dev": (C.c_int, [_vp, C.POINTER(PcmJob), _u32]),
    "blast_pcm_decode_batch": (C.c_int, [_vp, _u32, C.POINTER(_vp), C.POINTER(_sz), C.POINTER(PcmDesc),
                                         C.POINTER(_vp), C.POINTER(_vp)]),
    "blast_pcm24_unpack_dev": (C.c_int, [_vp, C.POINTER(Pcm24Job), _u32]),
    "blast_scene_create": (C.c_int, [_vp, C.POINTER(Track), _u32, C.POINTER(Voice), _u32, _u32, C.POINTER(_vp)]),
    "blast_scene_destroy": (None, [_vp, _vp]),
    "blast_scene_set_voices": (C.c_int, [_vp, _vp, C.POINTER(Voice), _u32]),
    "blast_scene_restore_dev": (C.c_int, [_vp, _vp]),
    "blast_scene_get_voices": (C.c_int, [_vp, _vp, C.POINTER(Voice), _u32]),
    "blast_scene_render_dev": (C.c_int, [_vp, _vp, _u64, _vp]),
    "blast_scene_reserve": (C.c_int, [_vp, _vp, _u64]),
    "blast_scene_check": (C.c_int, [_vp, _vp]),
    "blast_bus_finalize_dev": (C.c_int, [_vp, _vp, _vp, _u64]),
    "blast_x128p_seed": (None, [_u64, C.POINTER(X128PState)]),
    "blast_x128p_advance": (C.c_int, [C.POINTER(X128PState), _u64, C.POINTER(X128PState)]),
    "blast_x128p_jump_dev": (C.c_int, [_vp, C.POINTER(X128PState), _u64, _u64, _vp]),
    "blast_x128p_fill_dev": (C.c_int, [_vp, _vp, _u64, _u64, C.c_int64, C.c_int64, _vp, _vp, _vp]),
    "blast_x128p_fill": (C.c_int, [_vp, _u64, _u64, _u64, _u64, C.c_int64, C.c_int64, _vp, _vp, _vp]),
    "blast_mpeg_header_info": (C.c_int, [_u32, C.POINTER(MpegHeader)]),
    "blast_mpeg_scan_dev": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(_u64)]),
    "blast_mpeg_index_dev": (C.c_int, [_vp, _vp, _u64, C.c_int, _vp, _u64, C.POINTER(_u64), C.POINTER(_u32),
                                       C.POINTER(_u64)]),
    "blast_mpeg_hist_dev": (C.c_int, [_vp, _vp, _u64, _vp]),
    "blast_mpeg_pick_ref_dev": (C.c_int, [_vp, _vp, C.POINTER(_u32)]),
    "blast_mpeg_first_pos_dev": (C.c_int, [_vp, _vp, _vp, _u64, _u32, _vp]),
    "blast_mpeg_classify_dev": (C.c_int, [_vp, _vp, _vp, _u64, _u32, _vp, _u64, _vp, _u64, C.POINTER(_u64)]),
    "blast_mpeg_shard_walk_dev": (C.c_int, [_vp, _vp, _u64, _u64, C.POINTER(MpegShardAgg)]),
    "blast_mpeg_shard_emit_dev": (C.c_int, [_vp, _vp, _u64, _u64, _u32, _u64, _vp, _vp, _u64, C.POINTER(_u64)]),
    "blast_mpeg_gather_dev": (C.c_int, [_vp, _vp, _u64, _vp, _u64, _vp, _u64, C.POINTER(_u64)]),
    "blast_mpeg_parse": (C.c_int, [_vp, _vp, _u64, C.c_int, _vp, _u64, C.POINTER(_u64), C.POINTER(_u32),
                                   C.POINTER(_u64), _vp, _u64, C.POINTER(_u64)]),
    "blast_convert_interval": (C.c_float, [_u32, _u32, C.c_float]),
    "blast_conductor_create": (C.c_int, [_vp, _u32, _u32, C.POINTER(Track), _u32, C.POINTER(_vp)]),
    "blast_conductor_destroy": (None, [_vp, _vp]),
    "blast_conductor_apply": (C.c_int, [_vp, _vp, C.POINTER(Command)]),
    "blast_conductor_set_shard": (C.c_int, [_vp, _u32, _u32]),
    "blast_conductor_reserve": (C.c_int, [_vp, _vp, _u64]),
    "blast_conductor_render_dev": (C.c_int, [_vp, _vp, _u64, _vp]),
    "blast_conductor_coordinate": (C.c_int, [_vp, _vp, _u64, _vp]),
    "blast_conductor_render_timeline_dev": (C.c_int, [_vp, _vp, C.POINTER(TimedCommand), _u32, _u64, _vp]),
    "blast_conductor_render_timeline": (C.c_int, [_vp, _vp, C.POINTER(TimedCommand), _u32, _u64, _vp]),
    "blast_conductor_n_voices": (C.c_int, [_vp, C.c_int]),
    "blast_conductor_n_groups": (C.c_int, [_vp]),
    "blast_conductor_get_voice": (C.c_int, [_vp, C.c_int, _u32, C.POINTER(VoiceState)]),
    "blast_conductor_set_voice": (C.c_int, [_vp, C.c_int, _u32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                            C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "blast_conductor_clock": (_u64, [_vp]),
    "blast_peer_bus_create": (C.c_int, [_vp, _u64, _u32, _u32, _u32, C.POINTER(_vp)]),
    "blast_peer_bus_destroy": (None, [_vp, _vp]),
    "blast_peer_bus_export": (C.c_int, [_vp, _vp, _vp]),
    "blast_peer_bus_connect_ipc": (C.c_int, [_vp, _vp, _vp]),
    "blast_peer_bus_connect_local": (C.c_int, [C.POINTER(_vp), _u32]),
    "blast_peer_bus_set_fused": (C.c_int, [_vp, C.c_int]),
    "blast_peer_bus_partial": (_vp, [_vp]),
    "blast_peer_bus_bus": (_vp, [_vp]),
    "blast_scene_render_reduce_dev": (C.c_int, [_vp, _vp, _u64, _vp]),
    "blast_peer_bus_begin_dev": (C.c_int, [_vp, _vp]),
    "blast_peer_bus_reduce_dev": (C.c_int, [_vp, _vp, _u64]),
    "blast_peer_bus_wait_dev": (C.c_int, [_vp, _vp]),
    "blast_peer_bus_flags": (C.c_int, [_vp, _vp, _vp, _u32]),
    "blast_peer_bus_check": (C.c_int, [_vp, _vp]),
    "blast_conductor_set_shard_by_track": (C.c_int, [_vp, _u32, _u32]),
    "blast_group_create": (C.c_int, [C.POINTER(_vp), C.POINTER(C.c_int), _u32]),
    "blast_group_destroy": (None, [_vp]),
    "blast_group_set_fused": (C.c_int, [_vp, C.c_int]),
    "blast_group_size": (_u32, [_vp]),
    "blast_group_ctx": (_vp, [_vp, _u32]),
    "blast_group_pcm_decode_batch": (C.c_int, [_vp, _u32, C.POINTER(_vp), C.POINTER(_sz), C.POINTER(PcmDesc), C.POINTER(_vp),
                                               C.POINTER(Track)]),
    "blast_group_free_tracks": (C.c_int, [_vp]),
    "blast_group_render": (C.c_int, [_vp, C.POINTER(Track), _u32, C.POINTER(Voice), _u32, _u32, _u64, _vp]),
    "blast_group_conductor_create": (C.c_int, [_vp, _u32, _u32, C.POINTER(Track), _u32, C.POINTER(_vp)]),
    "blast_group_conductor_destroy": (None, [_vp]),
    "blast_group_conductor_apply": (C.c_int, [_vp, C.POINTER(Command)]),
    "blast_group_conductor_coordinate": (C.c_int, [_vp, _u64, _vp]),
    "blast_group_conductor_member": (_vp, [_vp, _u32]),
    "blast_group_x128p_fill": (C.c_int, [_vp, _u64, _u64, _u64, _u64, C.c_int64, C.c_int64, _vp, _vp, _vp]),
    "blast_group_mpeg_index": (C.c_int, [_vp, _vp, _u64, C.c_int, _vp, _u64, C.POINTER(_u64), C.POINTER(_u32),
                                         C.POINTER(_u64)]),
    "blast_render": (C.c_int, [_vp, C.POINTER(Track), _u32, C.POINTER(Voice), _u32, _u32, _u64, _vp,
                               C.POINTER(Voice)]),
}

_lib = None


def load():
    """Load libblast_cuda.so and set the prototypes.  Raises ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build the CUDA extension first (make -C audio_decoder_b200/csrc, or "
            "__graft_entry__.build()).  audio_decoder_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
