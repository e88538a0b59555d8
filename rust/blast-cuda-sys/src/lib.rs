//! Raw bindings of include/blast_cuda.h.  Written by hand (no bindgen in the image) and NOT compiled here:
//! the build image has no Rust toolchain.  Every item mirrors the C declaration one to one.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const BLAST_OK: c_int = 0;
pub const BLAST_ERR_IO: c_int = 1;
pub const BLAST_ERR_UNSUPPORTED_FORMAT: c_int = 2;
pub const BLAST_ERR_UNEXPECTED_EOF: c_int = 3;
pub const BLAST_ERR_INVALID_DATA: c_int = 4;
pub const BLAST_ERR_REF_PANIC: c_int = 5;
pub const BLAST_ERR_TIMEOUT: c_int = 105;

#[repr(C)] pub struct blast_ctx { _p: [u8; 0] }
#[repr(C)] pub struct blast_scene { _p: [u8; 0] }
#[repr(C)] pub struct blast_pcm_plan { _p: [u8; 0] }
#[repr(C)] pub struct blast_conductor { _p: [u8; 0] }
#[repr(C)] pub struct blast_peer_bus { _p: [u8; 0] }
#[repr(C)] pub struct blast_group { _p: [u8; 0] }
#[repr(C)] pub struct blast_group_conductor { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct blast_pcm_desc {
    pub sample_rate: u32, pub num_channels: u32, pub bits_per_sample: u32, pub big_endian: u32,
    pub data_off: u64, pub data_len: u64,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct blast_pcm_job { pub d_src: *const u8, pub d_dst: *mut i16, pub n_words: u64, pub big_endian: u32, pub reserved: u32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct blast_track { pub d_samples: *const i16, pub n_samples: u64, pub num_channels: u32, pub sample_rate: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct blast_voice { pub track: u32, pub active: u32, pub position: f32, pub velocity: f32, pub gain: f32, pub reserved: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct blast_x128p { pub s0: u64, pub s1: u64 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct blast_mpeg_shard_agg { pub exit_state: [u32; 4], pub count: [u64; 4] }

// ---- Conductor (engine.rs:36-248, commands.rs:86-234): TempoMode / TempoUnit / Command / Idx discriminants
pub const BLAST_TM_PROCESS: u32 = 0; pub const BLAST_TM_VOICE: u32 = 1; pub const BLAST_TM_GROUP: u32 = 2;
pub const BLAST_TM_CONTEXT: u32 = 3; pub const BLAST_TM_TBD: u32 = 4;
pub const BLAST_TU_SAMPLES: u32 = 0; pub const BLAST_TU_MILLIS: u32 = 1; pub const BLAST_TU_BPM: u32 = 2;
pub const BLAST_CMD_LOAD: u32 = 0; pub const BLAST_CMD_START: u32 = 1; pub const BLAST_CMD_PAUSE: u32 = 2;
pub const BLAST_CMD_RESUME: u32 = 3; pub const BLAST_CMD_STOP: u32 = 4; pub const BLAST_CMD_UNLOAD: u32 = 5;
pub const BLAST_CMD_VELOCITY: u32 = 6; pub const BLAST_CMD_GROUP: u32 = 7; pub const BLAST_CMD_TC: u32 = 8;
pub const BLAST_CMD_SEQ: u32 = 9; pub const BLAST_CMD_QUIT: u32 = 10;
pub const BLAST_IDX_TEMPO: u32 = 0; pub const BLAST_IDX_VOICE: u32 = 1; pub const BLAST_IDX_PROCESS: u32 = 2;
pub const BLAST_IDX_GROUP: u32 = 3;

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct blast_tempo_repr { pub idx: u64, pub owned: u32, pub mode: u32, pub unit: u32, pub interval: f32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct blast_command {
    pub kind: u32, pub idx_kind: u32, pub idx: u64, pub val: f32, pub reserved: u32,
    pub tempo: blast_tempo_repr,
    pub n_members: u32, pub reserved2: u32,
    pub member_voice: *const u64, pub member_update_tempo: *const u8, pub member_n_procs: *const u32,
    pub member_proc_ids: *const u64,
    pub period: u64, pub n_steps: u32, pub reserved3: u32, pub steps: *const f32, pub chance: *const f32,
    pub rng_s0: u64, pub rng_s1: u64,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct blast_timed_command { pub frame: u64, pub cmd: blast_command }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct blast_voice_state {
    pub active: u32, pub position: f32, pub velocity: f32, pub gain: f32, pub end: u64, pub channels: u32,
    pub tempo_current: u32, pub tempo_active: u32, pub n_processes: u32,
}

extern "C" {
    pub fn blast_conductor_create(ctx: *mut blast_ctx, out_channels: u32, sample_rate: u32, tracks: *const blast_track,
                                  n_tracks: u32, out: *mut *mut blast_conductor) -> c_int;
    pub fn blast_conductor_destroy(ctx: *mut blast_ctx, c: *mut blast_conductor);
    pub fn blast_conductor_apply(ctx: *mut blast_ctx, c: *mut blast_conductor, cmd: *const blast_command) -> c_int;
    pub fn blast_conductor_set_shard(c: *mut blast_conductor, rank: u32, world: u32) -> c_int;
    pub fn blast_conductor_reserve(ctx: *mut blast_ctx, c: *mut blast_conductor, frames: u64) -> c_int;
    pub fn blast_conductor_render_dev(ctx: *mut blast_ctx, c: *mut blast_conductor, frames: u64, d_partial_bus: *mut i32) -> c_int;
    pub fn blast_conductor_coordinate(ctx: *mut blast_ctx, c: *mut blast_conductor, frames: u64, host_bus_out: *mut i16) -> c_int;
    pub fn blast_conductor_render_timeline(ctx: *mut blast_ctx, c: *mut blast_conductor, events: *const blast_timed_command,
                                           n_events: u32, total_frames: u64, host_bus_out: *mut i16) -> c_int;
    pub fn blast_conductor_get_voice(c: *const blast_conductor, group: c_int, idx: u32, out: *mut blast_voice_state) -> c_int;
    pub fn blast_conductor_set_voice(c: *mut blast_conductor, group: c_int, idx: u32, position: *const f32,
                                     velocity: *const f32, gain: *const f32, active: *const c_int) -> c_int;
    pub fn blast_conductor_clock(c: *const blast_conductor) -> u64;
    pub fn blast_convert_interval(sample_rate: u32, unit: u32, interval: f32) -> f32;
}

extern "C" {
    pub fn blast_last_error() -> *const c_char;
    pub fn blast_ctx_create(out: *mut *mut blast_ctx, device: c_int) -> c_int;
    pub fn blast_ctx_destroy(ctx: *mut blast_ctx);
    pub fn blast_ctx_sync(ctx: *mut blast_ctx) -> c_int;
    pub fn blast_dev_alloc(ctx: *mut blast_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn blast_dev_free(ctx: *mut blast_ctx, p: *mut c_void) -> c_int;
    pub fn blast_host_alloc(ctx: *mut blast_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn blast_host_free(ctx: *mut blast_ctx, p: *mut c_void) -> c_int;
    pub fn blast_memcpy_h2d(ctx: *mut blast_ctx, d_dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn blast_memcpy_d2h(ctx: *mut blast_ctx, dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;

    pub fn blast_wav_probe(file: *const u8, len: usize, out: *mut blast_pcm_desc) -> c_int;
    pub fn blast_aiff_probe(file: *const u8, len: usize, out: *mut blast_pcm_desc) -> c_int;
    pub fn blast_pcm_out_len(desc: *const blast_pcm_desc) -> usize;
    pub fn blast_file_name(path: *const c_char, out: *mut c_char, cap: usize) -> c_int;
    pub fn blast_pcm_decode_batch(ctx: *mut blast_ctx, n: u32, files: *const *const u8, lens: *const usize,
                                  descs: *const blast_pcm_desc, host_out: *const *mut i16, d_out: *const *mut i16) -> c_int;
    pub fn blast_pcm_decode_dev(ctx: *mut blast_ctx, jobs: *const blast_pcm_job, n_jobs: u32) -> c_int;

    pub fn blast_scene_create(ctx: *mut blast_ctx, tracks: *const blast_track, n_tracks: u32, voices: *const blast_voice,
                              n_voices: u32, out_channels: u32, out: *mut *mut blast_scene) -> c_int;
    pub fn blast_scene_destroy(ctx: *mut blast_ctx, scene: *mut blast_scene);
    pub fn blast_scene_set_voices(ctx: *mut blast_ctx, scene: *mut blast_scene, voices: *const blast_voice, n: u32) -> c_int;
    pub fn blast_scene_get_voices(ctx: *mut blast_ctx, scene: *mut blast_scene, out: *mut blast_voice, n: u32) -> c_int;
    pub fn blast_scene_render_dev(ctx: *mut blast_ctx, scene: *mut blast_scene, frames: u64, d_partial_bus: *mut i32) -> c_int;
    pub fn blast_scene_reserve(ctx: *mut blast_ctx, scene: *mut blast_scene, frames: u64) -> c_int;
    pub fn blast_scene_check(ctx: *mut blast_ctx, scene: *mut blast_scene) -> c_int;
    pub fn blast_bus_finalize_dev(ctx: *mut blast_ctx, d_partial: *const i32, d_bus: *mut i16, n_slots: u64) -> c_int;
    pub fn blast_render(ctx: *mut blast_ctx, tracks: *const blast_track, n_tracks: u32, voices: *const blast_voice,
                        n_voices: u32, out_channels: u32, frames: u64, host_bus_out: *mut i16,
                        voices_after: *mut blast_voice) -> c_int;

    pub fn blast_x128p_seed(seed: u64, out: *mut blast_x128p);
    pub fn blast_x128p_advance(state: *const blast_x128p, n_draws: u64, out: *mut blast_x128p) -> c_int;
    pub fn blast_x128p_fill(ctx: *mut blast_ctx, seed: u64, stride: u64, n_streams: u64, draws_per_stream: u64,
                            lower: i64, upper: i64, raw_out: *mut u64, ranged_out: *mut i64, checks_out: *mut u64) -> c_int;

    pub fn blast_asset_consensus(descs: *const blast_pcm_desc, n: u32, sample_rate_out: *mut u32, num_channels_out: *mut u32) -> c_int;
    pub fn blast_ctx_trim(ctx: *mut blast_ctx) -> c_int;

    // MPEG in steps / over several GPUs (device pointers)
    pub fn blast_mpeg_scan_dev(ctx: *mut blast_ctx, d_bytes: *const u8, len: u64, d_pos_out: *mut u64, d_hdr_out: *mut u32,
                               cap: u64, n_out: *mut u64) -> c_int;
    pub fn blast_mpeg_index_dev(ctx: *mut blast_ctx, d_bytes: *const u8, len: u64, reference_compat: c_int, d_offsets_out: *mut u64,
                                cap: u64, n_offsets_out: *mut u64, ref_header_out: *mut u32, n_candidates_out: *mut u64) -> c_int;
    pub fn blast_mpeg_gather_dev(ctx: *mut blast_ctx, d_bytes: *const u8, len: u64, d_offsets: *const u64, n_offsets: u64,
                                 d_payload_out: *mut u8, cap: u64, payload_len_out: *mut u64) -> c_int;
    pub fn blast_mpeg_hist_dev(ctx: *mut blast_ctx, d_hdr: *const u32, n: u64, d_hist: *mut u32) -> c_int;
    pub fn blast_mpeg_pick_ref_dev(ctx: *mut blast_ctx, d_hist: *const u32, ref_header_out: *mut u32) -> c_int;
    pub fn blast_mpeg_first_pos_dev(ctx: *mut blast_ctx, d_pos: *const u64, d_hdr: *const u32, n: u64, ref_header: u32,
                                    d_first: *mut u64) -> c_int;
    pub fn blast_mpeg_classify_dev(ctx: *mut blast_ctx, d_pos: *const u64, d_hdr: *const u32, n: u64, ref_header: u32,
                                   d_first: *const u64, stream_len: u64, d_offsets_out: *mut u64, cap: u64,
                                   n_offsets_out: *mut u64) -> c_int;
    pub fn blast_mpeg_shard_walk_dev(ctx: *mut blast_ctx, d_bytes: *const u8, own_len: u64, halo_len: u64,
                                     agg_out: *mut blast_mpeg_shard_agg) -> c_int;
    pub fn blast_mpeg_shard_emit_dev(ctx: *mut blast_ctx, d_bytes: *const u8, own_len: u64, halo_len: u64, entry_state: u32,
                                     pos_offset: u64, d_pos_out: *mut u64, d_hdr_out: *mut u32, cap: u64, n_out: *mut u64) -> c_int;

    // the mix reduction over peer memory: tile protocol inside the render kernel (ranks = processes or group members)
    pub fn blast_peer_bus_create(ctx: *mut blast_ctx, n_slots: u64, rank: u32, world: u32, root: u32, out: *mut *mut blast_peer_bus) -> c_int;
    pub fn blast_peer_bus_destroy(ctx: *mut blast_ctx, pb: *mut blast_peer_bus);
    pub fn blast_peer_bus_export(ctx: *mut blast_ctx, pb: *mut blast_peer_bus, handle_out: *mut u8) -> c_int;
    pub fn blast_peer_bus_connect_ipc(ctx: *mut blast_ctx, pb: *mut blast_peer_bus, handles: *const u8) -> c_int;
    pub fn blast_peer_bus_connect_local(all: *const *mut blast_peer_bus, world: u32) -> c_int;
    pub fn blast_peer_bus_set_fused(pb: *mut blast_peer_bus, fused: c_int) -> c_int;
    pub fn blast_peer_bus_partial(pb: *mut blast_peer_bus) -> *mut i32;
    pub fn blast_peer_bus_bus(pb: *mut blast_peer_bus) -> *mut i16;
    pub fn blast_scene_render_reduce_dev(ctx: *mut blast_ctx, scene: *mut blast_scene, frames: u64, pb: *mut blast_peer_bus) -> c_int;
    pub fn blast_peer_bus_begin_dev(ctx: *mut blast_ctx, pb: *mut blast_peer_bus) -> c_int;
    pub fn blast_peer_bus_reduce_dev(ctx: *mut blast_ctx, pb: *mut blast_peer_bus, n_slots_used: u64) -> c_int;
    pub fn blast_peer_bus_wait_dev(ctx: *mut blast_ctx, pb: *mut blast_peer_bus) -> c_int;
    pub fn blast_peer_bus_flags(ctx: *mut blast_ctx, pb: *mut blast_peer_bus, out: *mut u32, cap: u32) -> c_int;
    pub fn blast_peer_bus_check(ctx: *mut blast_ctx, pb: *mut blast_peer_bus) -> c_int;
    pub fn blast_conductor_set_shard_by_track(c: *mut blast_conductor, rank: u32, world: u32) -> c_int;

    // several GPUs driven by this one process (main.rs is one process): decode by file, render by track, RNG by stream
    pub fn blast_group_create(out: *mut *mut blast_group, device_ids: *const c_int, n_devices: u32) -> c_int;
    pub fn blast_group_destroy(g: *mut blast_group);
    pub fn blast_group_set_fused(g: *mut blast_group, fused: c_int) -> c_int;
    pub fn blast_group_size(g: *const blast_group) -> u32;
    pub fn blast_group_ctx(g: *mut blast_group, member: u32) -> *mut blast_ctx;
    pub fn blast_group_pcm_decode_batch(g: *mut blast_group, n: u32, files: *const *const u8, lens: *const usize,
                                        descs: *const blast_pcm_desc, host_out: *const *mut i16, tracks_out: *mut blast_track) -> c_int;
    pub fn blast_group_free_tracks(g: *mut blast_group) -> c_int;
    pub fn blast_group_render(g: *mut blast_group, tracks: *const blast_track, n_tracks: u32, voices: *const blast_voice,
                              n_voices: u32, out_channels: u32, frames: u64, host_bus_out: *mut i16) -> c_int;
    pub fn blast_group_conductor_create(g: *mut blast_group, out_channels: u32, sample_rate: u32, tracks: *const blast_track,
                                        n_tracks: u32, out: *mut *mut blast_group_conductor) -> c_int;
    pub fn blast_group_conductor_destroy(gc: *mut blast_group_conductor);
    pub fn blast_group_conductor_apply(gc: *mut blast_group_conductor, cmd: *const blast_command) -> c_int;
    pub fn blast_group_conductor_coordinate(gc: *mut blast_group_conductor, frames: u64, host_bus_out: *mut i16) -> c_int;
    pub fn blast_group_conductor_member(gc: *mut blast_group_conductor, member: u32) -> *mut blast_conductor;
    pub fn blast_group_x128p_fill(g: *mut blast_group, seed: u64, stride: u64, n_streams: u64, draws_per_stream: u64,
                                  lower: i64, upper: i64, raw_out: *mut u64, ranged_out: *mut i64, checks_out: *mut u64) -> c_int;
    pub fn blast_group_mpeg_index(g: *mut blast_group, bytes: *const u8, len: u64, reference_compat: c_int, offsets_out: *mut u64,
                                  cap: u64, n_offsets_out: *mut u64, ref_header_out: *mut u32, n_candidates_out: *mut u64) -> c_int;

    pub fn blast_mpeg_parse(ctx: *mut blast_ctx, bytes: *const u8, len: u64, reference_compat: c_int,
                            offsets_out: *mut u64, offsets_cap: u64, n_offsets_out: *mut u64, ref_header_out: *mut u32,
                            n_candidates_out: *mut u64, payload_out: *mut u8, payload_cap: u64,
                            payload_len_out: *mut u64) -> c_int;
}
