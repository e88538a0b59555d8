/*
 * blast_cuda.h — C ABI of libblast_cuda.so, the B200 (sm_100a) implementation of BLAST's
 * data-parallel hot path (gitxandert/audio_decoder).  This is the drop-in boundary: a
 * Rust (or any FFI) host binds exactly these symbols; INTEGRATION.md shows the
 * `extern "C"` block and the replacement bodies of file_parsing::{wav,aiff,mpeg}::parse,
 * Conductor::coordinate and X128P that sit on top of it.
 *
 * Each entry point cites the reference interface it replaces (file:line relative to the
 * reference root).  The reference has no FFI of its own (its only extern "C" item is the
 * SIGTERM handler, blast/src/audio_processing/runtime.rs:400), so the seam is the Rust
 * function/struct surface; this header is its C projection.
 *
 * Conventions
 *  - every function returns a status code (0 = OK); codes 1..4 are the reference's
 *    DecodeError variants (blast/src/file_parsing/decode_helpers.rs:1-7); code 5 marks inputs
 *    on which the reference PANICS (index out of bounds, usize underflow) — never UB here;
 *    codes >= 100 are this library's own.  blast_last_error() gives the message
 *    (thread-local).
 *  - all buffers are caller-owned.  "d_" parameters are device pointers on the context's
 *    GPU, everything else is host memory.
 *  - one blast_ctx is bound to one GPU and may be driven by one host thread at a time
 *    (the reference's callers are single-threaded: main.rs:18-89, runtime.rs:320-380).
 *    Work is enqueued on the context's CUDA stream; functions named *_dev are
 *    asynchronous, the host-buffer entry points synchronise before returning.
 *  - there is NO CPU compute fallback: without a usable CUDA device blast_ctx_create fails
 *    with BLAST_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef BLAST_CUDA_H
#define BLAST_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLAST_ABI_VERSION 3

enum {
    BLAST_OK = 0,
    BLAST_ERR_IO = 1,                 /* DecodeError::Io */
    BLAST_ERR_UNSUPPORTED_FORMAT = 2, /* DecodeError::UnsupportedFormat */
    BLAST_ERR_UNEXPECTED_EOF = 3,     /* DecodeError::UnexpectedEof */
    BLAST_ERR_INVALID_DATA = 4,       /* DecodeError::InvalidData */
    BLAST_ERR_REF_PANIC = 5,          /* the reference would panic on this input */
    BLAST_ERR_CUDA = 100,
    BLAST_ERR_ARG = 101,
    BLAST_ERR_NO_DEVICE = 102,
    BLAST_ERR_CAPACITY = 103,
    BLAST_ERR_UNSUPPORTED = 104,
    BLAST_ERR_TIMEOUT = 105           /* a bounded device-side wait for a peer GPU gave up */
};

typedef struct blast_ctx blast_ctx;

/* ------------------------------------------------------------------ lifecycle */
int  blast_abi_version(void);
const char* blast_last_error(void);
/* Binds a context to CUDA device `device` and creates its stream.  Fails loudly
 * (BLAST_ERR_NO_DEVICE) if there is no sm_100 device: there is no CPU path. */
int  blast_ctx_create(blast_ctx** out, int device);
void blast_ctx_destroy(blast_ctx* ctx);
/* Use an existing cudaStream_t (e.g. torch's current stream) instead of the context's own. */
int  blast_ctx_set_stream(blast_ctx* ctx, void* cuda_stream);
void* blast_ctx_stream(blast_ctx* ctx);
int  blast_ctx_sync(blast_ctx* ctx);
int  blast_ctx_device(const blast_ctx* ctx);
int  blast_ctx_sm_count(const blast_ctx* ctx);
/* Releases the context's grow-only device scratch (staging slabs, candidate lists, tile tables).  The hot entry
 * points never cudaMalloc / cudaFree once warm; this gives the memory back between workloads. */
int  blast_ctx_trim(blast_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t blast_ctx_launch_count(const blast_ctx* ctx);

/* memory helpers (device allocations are 256-byte aligned and padded to 256 bytes) */
int blast_dev_alloc(blast_ctx* ctx, size_t bytes, void** out);
int blast_dev_free(blast_ctx* ctx, void* d_ptr);
int blast_host_alloc(blast_ctx* ctx, size_t bytes, void** out);   /* pinned */
int blast_host_free(blast_ctx* ctx, void* ptr);
int blast_memcpy_h2d(blast_ctx* ctx, void* d_dst, const void* src, size_t bytes);   /* async */
int blast_memcpy_d2h(blast_ctx* ctx, void* dst, const void* d_src, size_t bytes);   /* async */
int blast_memcpy_d2d(blast_ctx* ctx, void* d_dst, const void* d_src, size_t bytes); /* async */
int blast_memset_dev(blast_ctx* ctx, void* d_dst, int value, size_t bytes);         /* async */

/* CUDA-event timing on the context's stream (the stream the kernels are launched on) */
typedef struct blast_event blast_event;
int  blast_event_create(blast_ctx* ctx, blast_event** out);
void blast_event_destroy(blast_event* ev);
int  blast_event_record(blast_ctx* ctx, blast_event* ev);
int  blast_event_elapsed_ms(blast_event* start, blast_event* stop, float* ms_out);  /* syncs on stop */

/* ------------------------------------------------------------------ L0: PCM decode
 * replaces file_parsing::wav::parse (blast/src/file_parsing/wav.rs:69-167) and
 * file_parsing::aiff::parse (blast/src/file_parsing/aiff.rs:99-183). */
typedef struct {
    uint32_t sample_rate;      /* AudioFile.sample_rate  (decode_helpers.rs:21) */
    uint32_t num_channels;     /* AudioFile.num_channels */
    uint32_t bits_per_sample;  /* AudioFile.bits_per_sample (reported, never used to unpack) */
    uint32_t big_endian;       /* 0 = "wav" (i16::from_le_bytes), 1 = "aiff" (i16::from_be_bytes) */
    uint64_t data_off;         /* offset of the first payload byte in the file image */
    uint64_t data_len;         /* declared payload bytes (data_size / ssnd_size) */
} blast_pcm_desc;

/* Host header walks: wav.rs:69-138 and aiff.rs:99-154, every quirk kept (ids never
 * compared, the +91 extensible skip, COMM size must be 18, 80-bit rate -> f64 -> u32).
 * They also apply the sample loop's bounds rule (wav.rs:143-151, aiff.rs:159-167): if any
 * byte pair of the declared payload is missing the whole parse is UnexpectedEof. */
int    blast_wav_probe(const uint8_t* file, size_t len, blast_pcm_desc* out);
int    blast_aiff_probe(const uint8_t* file, size_t len, blast_pcm_desc* out);
size_t blast_pcm_out_len(const blast_pcm_desc* desc);   /* ceil(data_len / 2) i16 words */
/* file-name rule applied after decoding (wav.rs:156-164, aiff.rs:172-180) */
int    blast_file_name(const char* path, char* out, size_t cap);

/* Asset-set consensus of main() (blast/src/main.rs:79-120): the most frequent sample rate (ties: the smallest; the
 * reference follows HashMap order) and the largest channel count of the successfully decoded files; 44100 / 2
 * when there are none.  These are what run_blast() opens the output device with. */
int    blast_asset_consensus(const blast_pcm_desc* descs, uint32_t n, uint32_t* sample_rate_out, uint32_t* num_channels_out);

/* One decode job on device-resident bytes: n_words byte pairs at d_src (ANY byte
 * alignment) -> int16 at d_dst (2-byte aligned).  d_src must be readable up to the next
 * 16-byte boundary past its last byte (true for blast_dev_alloc / cudaMalloc buffers). */
typedef struct {
    const uint8_t* d_src;
    int16_t*       d_dst;
    uint64_t       n_words;
    uint32_t       big_endian;
    uint32_t       reserved;
} blast_pcm_job;

/* A plan holds the device-side job + tile tables so that a batch can be (re)launched as a
 * single kernel with nothing but the launch inside the timed region. */
typedef struct blast_pcm_plan blast_pcm_plan;
int  blast_pcm_plan_create(blast_ctx* ctx, const blast_pcm_job* jobs, uint32_t n_jobs, blast_pcm_plan** out);
int  blast_pcm_plan_run_dev(blast_ctx* ctx, blast_pcm_plan* plan);          /* async, 1 launch */
void blast_pcm_plan_destroy(blast_ctx* ctx, blast_pcm_plan* plan);
uint64_t blast_pcm_plan_words(const blast_pcm_plan* plan);
/* one-shot: create + run + destroy */
int  blast_pcm_decode_dev(blast_ctx* ctx, const blast_pcm_job* jobs, uint32_t n_jobs);

/* The parse() drop-in for a batch of in-memory file images (host buffers, ideally pinned):
 * copies each payload to the GPU, decodes, and delivers AudioFile.samples to
 *   host_out[i]  (nullable array / nullable entries; blast_pcm_out_len(descs[i]) words) and/or
 *   d_out[i]     (nullable array / nullable entries; device, stays resident for the render).
 * Copies and kernels are pipelined over internal streams; returns after everything landed. */
int  blast_pcm_decode_batch(blast_ctx* ctx, uint32_t n, const uint8_t* const* files, const size_t* lens,
                            const blast_pcm_desc* descs, int16_t* const* host_out, int16_t* const* d_out);

/* Extension (not in the reference): true packed 24-bit unpack.
 * out_kind 0: sign-extended int32 per sample; 1: top 16 bits as int16. */
typedef struct {
    const uint8_t* d_src;      /* 3 bytes per sample, any alignment */
    void*          d_dst;      /* int32* or int16*, naturally aligned */
    uint64_t       n_samples;
    uint32_t       big_endian;
    uint32_t       out_kind;
} blast_pcm24_job;
int  blast_pcm24_unpack_dev(blast_ctx* ctx, const blast_pcm24_job* jobs, uint32_t n_jobs);

/* ------------------------------------------------------------------ L1: voice render / mix-down
 * replaces Conductor::coordinate (blast/src/audio_processing/engine.rs:46-81) and Voice::process
 * (engine.rs:386-448) for a set of voices over `frames` output frames.  Output is bit-exact with
 * the reference's release-build semantics (saturating `as i16`, wrapping i16 accumulate). */
/* d_samples must be readable from the 16-byte boundary at or below it up to the 16-byte boundary at or above its last
 * sample (the render stages source spans with 16-byte aligned bulk copies): true for blast_dev_alloc / cudaMalloc
 * buffers and for tracks sub-allocated from them. */
typedef struct {
    const int16_t* d_samples;  /* device, interleaved, 4-byte aligned (AudioFile.samples) */
    uint64_t n_samples;        /* samples.len() */
    uint32_t num_channels;     /* AudioFile.num_channels */
    uint32_t sample_rate;      /* informational */
} blast_track;

typedef struct {               /* VoiceState (engine.rs:279-286), flattened; `end` is derived like */
    uint32_t track;            /* Voice::new does: samples.len()/channels - 1 (engine.rs:302)      */
    uint32_t active;
    float    position;
    float    velocity;
    float    gain;             /* no Command sets it; exposed here as a plain field */
    uint32_t reserved;
} blast_voice;

typedef struct blast_scene blast_scene;
/* Uploads the voice table.  BLAST_ERR_REF_PANIC for inputs on which Voice::new / load panic
 * (track index out of range, 0 channels, empty track). out_channels 1..8. */
int  blast_scene_create(blast_ctx* ctx, const blast_track* tracks, uint32_t n_tracks, const blast_voice* voices,
                        uint32_t n_voices, uint32_t out_channels, blast_scene** out);
void blast_scene_destroy(blast_ctx* ctx, blast_scene* scene);
int  blast_scene_set_voices(blast_ctx* ctx, blast_scene* scene, const blast_voice* voices, uint32_t n_voices);
/* async: puts every voice back to the state uploaded by the last create / set_voices (rewind) */
int  blast_scene_restore_dev(blast_ctx* ctx, blast_scene* scene);
/* reads the voice states back (positions as left by the last render); synchronises */
int  blast_scene_get_voices(blast_ctx* ctx, blast_scene* scene, blast_voice* out, uint32_t n_voices);
/* Renders `frames` frames of all active voices into the int32 partial bus
 * d_partial_bus[frames * out_channels] (overwritten) and advances the voices, exactly as `frames`
 * iterations of coordinate()'s outer loop would.  Async: position scan + render/mix launches.
 * Partial buses of several GPUs may be summed (int32) before blast_bus_finalize_dev. */
int  blast_scene_render_dev(blast_ctx* ctx, blast_scene* scene, uint64_t frames, int32_t* d_partial_bus);
/* Allocates now what a render of `frames` frames would allocate on first use (grow-only scratch): the renders that
 * follow neither allocate nor synchronise. */
int  blast_scene_reserve(blast_ctx* ctx, blast_scene* scene, uint64_t frames);
/* synchronises and reports deferred device-side errors of the last render (BLAST_ERR_CAPACITY) */
int  blast_scene_check(blast_ctx* ctx, blast_scene* scene);
/* S16 bus = low 16 bits of the int32 partial sums (== i16 wrapping accumulate, engine.rs:441) */
int  blast_bus_finalize_dev(blast_ctx* ctx, const int32_t* d_partial, int16_t* d_bus, uint64_t n_slots);
/* ---- The mix reduction across the GPUs of one box WITHOUT a collective library (SURVEY.md §8 e: the render's one
 * exchange step).  Every rank (one per GPU; ranks may be processes — CUDA IPC — or contexts of one process — peer
 * access) owns a WINDOW: its int32 partial bus, an S16 bus (the result, complete on the root) and flag tables, mapped
 * into every peer's address space.  The bus is cut into tiles; tile t is reduced by rank t mod world.  Per step, on
 * every rank, in the same order on every rank:
 *     blast_scene_render_reduce_dev    render + exchange.  The exchange: every rank raises, per tile, a system-scope
 *                                      flag at the tile's owner; the owner waits for the tile's flags, sums the peers'
 *                                      int32 tiles with NVLink loads, wraps to S16 (engine.rs:441) and stores into the
 *                                      ROOT's bus.  Either two small kernels after the render (default) or inside the
 *                                      render kernel's work queue, tile by tile as the tiles complete
 *                                      (blast_peer_bus_set_fused).
 *  or blast_peer_bus_begin_dev, <stream work that fills blast_peer_bus_partial>, blast_peer_bus_reduce_dev
 *                                      the exchange on its own (Conductor spans).
 *     blast_peer_bus_wait_dev          joins the exchange: it runs on a stream of its own behind the render, BESIDE what
 *                                      the caller enqueues next (the next batch's decode), so a rank waits for its slowest
 *                                      peer there and not in the caller's stream.  After it, on the root, the bus of the
 *                                      step is complete in stream order.  Call it before the bus is consumed (and at the
 *                                      end of a run); the next step orders itself behind the exchange on its own.
 * A rank may overwrite its partial bus again only after every rank has finished reading it: the next step's first
 * kernel waits for those acknowledgements on the device.  All waits are bounded (BLAST_PEER_TIMEOUT_MS, default
 * 20,000): a rank that never publishes costs its peers the timeout and BLAST_ERR_TIMEOUT from blast_peer_bus_check,
 * not a hung GPU.  world == 1 is the single-GPU case: no exchange, the bus is finalized in place. */
typedef struct blast_peer_bus blast_peer_bus;
#define BLAST_PEER_HANDLE_BYTES 64
int  blast_peer_bus_create(blast_ctx* ctx, uint64_t n_slots, uint32_t rank, uint32_t world, uint32_t root, blast_peer_bus** out);
void blast_peer_bus_destroy(blast_ctx* ctx, blast_peer_bus* pb);
/* multi-process: export this rank's window, exchange the handles (any transport), connect with all of them in rank order */
int  blast_peer_bus_export(blast_ctx* ctx, blast_peer_bus* pb, uint8_t handle_out[BLAST_PEER_HANDLE_BYTES]);
int  blast_peer_bus_connect_ipc(blast_ctx* ctx, blast_peer_bus* pb, const uint8_t* handles /* world x 64 bytes */);
/* one process: all[r] = rank r's peer bus (peer access between the GPUs is enabled here); call once, for all ranks */
int  blast_peer_bus_connect_local(blast_peer_bus* const* all, uint32_t world);
/* blast_scene_render_reduce_dev does the exchange inside the render kernel (1) or as launches of its own after it (0,
 * the default: the faster one on B200s).  Same value on every rank. */
int  blast_peer_bus_set_fused(blast_peer_bus* pb, int fused);
int32_t* blast_peer_bus_partial(blast_peer_bus* pb);     /* this rank's int32 partial bus [n_slots] */
int16_t* blast_peer_bus_bus(blast_peer_bus* pb);         /* this rank's S16 bus [n_slots]: the result on the root */
int  blast_scene_render_reduce_dev(blast_ctx* ctx, blast_scene* scene, uint64_t frames, blast_peer_bus* pb);
int  blast_peer_bus_begin_dev(blast_ctx* ctx, blast_peer_bus* pb);
int  blast_peer_bus_reduce_dev(blast_ctx* ctx, blast_peer_bus* pb, uint64_t n_slots_used);
int  blast_peer_bus_wait_dev(blast_ctx* ctx, blast_peer_bus* pb);
/* diagnostic: {step, world, then per rank r: ready[r][tile 0..3], done[r], ack[r]} of this rank's window, read beside
 * whatever may still be waiting */
int  blast_peer_bus_flags(blast_ctx* ctx, blast_peer_bus* pb, uint32_t* out, uint32_t cap);
/* synchronises; BLAST_ERR_TIMEOUT if a device-side wait of this rank gave up since the last check */
int  blast_peer_bus_check(blast_ctx* ctx, blast_peer_bus* pb);
/* one-shot with a host bus (interleaved S16_LE like the ALSA area, runtime.rs:272-276) */
int  blast_render(blast_ctx* ctx, const blast_track* tracks, uint32_t n_tracks, const blast_voice* voices,
                  uint32_t n_voices, uint32_t out_channels, uint64_t frames, int16_t* host_bus_out,
                  blast_voice* voices_after /* nullable */);

/* ------------------------------------------------------------------ L1b: the Conductor (Command timeline, tempo, Seq)
 * replaces Conductor::{prepare, apply, coordinate} (blast/src/audio_processing/engine.rs:36-248), Voice /
 * Group transport (engine.rs:318-384, 477-528), TempoState (blast_time.rs:58-161) and the Seq process
 * (processes.rs:52-99).  The Command payloads are the reference's (commands.rs:86-234) with names already
 * resolved to indices, exactly what CmdProcessor hands to the audio thread.  The control state machine runs
 * on the host (it is a few scalar updates per command); every sample, every position step and every Seq
 * event / RNG draw is computed on the GPU. */
enum { BLAST_TM_PROCESS = 0, BLAST_TM_VOICE = 1, BLAST_TM_GROUP = 2, BLAST_TM_CONTEXT = 3, BLAST_TM_TBD = 4 };  /* TempoMode */
enum { BLAST_TU_SAMPLES = 0, BLAST_TU_MILLIS = 1, BLAST_TU_BPM = 2 };                                          /* TempoUnit */
enum { BLAST_CMD_LOAD = 0, BLAST_CMD_START, BLAST_CMD_PAUSE, BLAST_CMD_RESUME, BLAST_CMD_STOP, BLAST_CMD_UNLOAD,
       BLAST_CMD_VELOCITY, BLAST_CMD_GROUP, BLAST_CMD_TC, BLAST_CMD_SEQ, BLAST_CMD_QUIT };                   /* Command */
enum { BLAST_IDX_TEMPO = 0, BLAST_IDX_VOICE = 1, BLAST_IDX_PROCESS = 2, BLAST_IDX_GROUP = 3 };                 /* Idx */

typedef struct {               /* TempoRepr (commands.rs:187-234) */
    uint64_t idx;              /* borrowed tempo: index into voices / groups / tempo contexts (by mode) */
    uint32_t owned;            /* 1: a new TempoState initialised with (mode, unit, interval) */
    uint32_t mode;             /* BLAST_TM_* */
    uint32_t unit;             /* BLAST_TU_* */
    float    interval;         /* in `unit`; converted like convert_interval (blast_time.rs:151-161) */
} blast_tempo_repr;

typedef struct {               /* Command + its Args struct (commands.rs:86-161), flattened */
    uint32_t kind;             /* BLAST_CMD_* */
    uint32_t idx_kind;         /* BLAST_IDX_* for Start / Pause / Resume / Stop / Seq */
    uint64_t idx;              /* Idx payload; LoadArgs.track_idx; UnloadArgs.idx; VelocityArgs.idx */
    float    val;              /* VelocityArgs.val */
    uint32_t reserved;
    blast_tempo_repr tempo;    /* LoadArgs.tempo_repr / GroupArgs.tempo / TcArgs.tempo / SeqArgs.tempo */
    /* GroupArgs.vs_fs_ps: member m = (voice index at the time it is removed, update_tempo, process ids) */
    uint32_t n_members;
    uint32_t reserved2;
    const uint64_t* member_voice;
    const uint8_t*  member_update_tempo;   /* nullable = all false */
    const uint32_t* member_n_procs;        /* nullable = all 0 */
    const uint64_t* member_proc_ids;       /* concatenated over members */
    /* SeqArgs (jit is carried by the reference but never read: processes.rs:61) */
    uint64_t period;
    uint32_t n_steps;
    uint32_t reserved3;
    const float* steps;
    const float* chance;       /* n_steps entries, like steps (commands.rs:944-945) */
    uint64_t rng_s0, rng_s1;   /* SeqArgs.rng state: blast_x128p_seed(seed) replaces fast_seed() (commands.rs:838) */
} blast_command;

typedef struct {               /* one entry of an offline timeline: `cmd` is applied before frame `frame` is rendered */
    uint64_t frame;
    blast_command cmd;
} blast_timed_command;

typedef struct {               /* VoiceState (engine.rs:279-286) + what Voice::new derives */
    uint32_t active;
    float    position;
    float    velocity;
    float    gain;
    uint64_t end;
    uint32_t channels;
    uint32_t tempo_current;    /* state.tempo.current */
    uint32_t tempo_active;
    uint32_t n_processes;
} blast_voice_state;

typedef struct blast_conductor blast_conductor;
/* convert_interval (blast_time.rs:151-161) with sample_rate::get() = sample_rate */
float blast_convert_interval(uint32_t sample_rate, uint32_t unit, float interval);
/* Conductor::prepare (engine.rs:36-44) + sample_rate::set (runtime.rs:37).  Tracks stay where they are in HBM;
 * voices refer to them by index (the reference clones the samples per voice, engine.rs:309). */
int  blast_conductor_create(blast_ctx* ctx, uint32_t out_channels, uint32_t sample_rate, const blast_track* tracks,
                            uint32_t n_tracks, blast_conductor** out);
void blast_conductor_destroy(blast_ctx* ctx, blast_conductor* c);
/* Conductor::apply (engine.rs:83-248).  Out-of-range indices (the reference `unwrap()`s / indexes: panic) return
 * BLAST_ERR_REF_PANIC and leave the state untouched.  Quit is accepted and ignored (the reference raises SIGTERM). */
int  blast_conductor_apply(blast_ctx* ctx, blast_conductor* c, const blast_command* cmd);
/* Multi-GPU: this rank renders the voices whose load order number is congruent to rank mod world; every rank
 * applies every command and tracks every tempo.  Partial buses are summed (int32) before finalising. */
int  blast_conductor_set_shard(blast_conductor* c, uint32_t rank, uint32_t world);
/* The same when the TRACKS are sharded (file i decoded on GPU i mod world, blast_group): a voice is rendered on the
 * rank that holds its track, track t on rank t mod world; tracks of other ranks are never dereferenced here. */
int  blast_conductor_set_shard_by_track(blast_conductor* c, uint32_t rank, uint32_t world);
/* `frames` iterations of coordinate()'s frame loop (engine.rs:46-81) into the int32 partial bus
 * d_partial_bus[frames * out_channels] (overwritten); voices, Seqs, tempi and the clock advance.  Returns after
 * the device work has finished (the Seq / position state is read back). BLAST_ERR_REF_PANIC where a Seq would index
 * an empty step list (processes.rs:79). */
int  blast_conductor_render_dev(blast_ctx* ctx, blast_conductor* c, uint64_t frames, int32_t* d_partial_bus);
/* Allocates now what a span of `frames` frames of the current scene would allocate on first use (grow-only): the spans
 * that follow neither allocate nor wait for an allocation. */
int  blast_conductor_reserve(blast_ctx* ctx, blast_conductor* c, uint64_t frames);
/* coordinate() with a host bus: interleaved S16_LE like the ALSA area (runtime.rs:272-276) */
int  blast_conductor_coordinate(blast_ctx* ctx, blast_conductor* c, uint64_t frames, int16_t* host_bus_out);
/* Offline render of a whole Command timeline (the reference applies queued commands between periods,
 * runtime.rs:326-328): events sorted by frame, frame <= total_frames.  Commands at frame f take effect before
 * frame f is rendered. */
int  blast_conductor_render_timeline_dev(blast_ctx* ctx, blast_conductor* c, const blast_timed_command* events,
                                         uint32_t n_events, uint64_t total_frames, int32_t* d_partial_bus);
int  blast_conductor_render_timeline(blast_ctx* ctx, blast_conductor* c, const blast_timed_command* events,
                                     uint32_t n_events, uint64_t total_frames, int16_t* host_bus_out);
/* state access: group < 0 addresses Conductor.voices, else Conductor.groups[group].voices */
int  blast_conductor_n_voices(const blast_conductor* c, int group);       /* -1: no such group */
int  blast_conductor_n_groups(const blast_conductor* c);
int  blast_conductor_get_voice(const blast_conductor* c, int group, uint32_t idx, blast_voice_state* out);
/* VoiceState's fields are pub (engine.rs:279-286); gain in particular has no Command.  Nullable = keep. */
int  blast_conductor_set_voice(blast_conductor* c, int group, uint32_t idx, const float* position,
                               const float* velocity, const float* gain, const int* active);
uint64_t blast_conductor_clock(const blast_conductor* c);                 /* clock::current (blast_time.rs:29-31) */

/* ------------------------------------------------------------------ RNG parameter streams
 * replaces X128P (blast/src/audio_processing/blast_rand.rs:4-60): xoroshiro128+ with the 55/14/36
 * constants, SplitMix64 seeding, Lemire multiply-shift range WITHOUT rejection. */
typedef struct { uint64_t s0, s1; } blast_x128p;
/* X128P::new(seed) (blast_rand.rs:10-24) */
void blast_x128p_seed(uint64_t seed, blast_x128p* out);
/* host-side jump of one generator by n_draws (table construction only, no GPU needed) */
int  blast_x128p_advance(const blast_x128p* in, uint64_t n_draws, blast_x128p* out);
/* Jump-ahead (not in the reference; GF(2) matrix powers): d_states_out[s] = `base` advanced by
 * s * stride draws, so draw j of stream s is draw s*stride + j of the CPU's sequential sequence. */
int  blast_x128p_jump_dev(blast_ctx* ctx, const blast_x128p* base, uint64_t stride, uint64_t n_streams,
                          blast_x128p* d_states_out);
/* draws_per_stream draws from every stream, states advanced in place.  Nullable outputs:
 *   d_raw    [n_streams * draws]  next_u64()                      (blast_rand.rs:31-39)
 *   d_ranged [n_streams * draws]  next_i64_range(lower, upper)    (blast_rand.rs:50-59)
 *   d_checks [n_streams * 4]      {xor raw, sum raw, xor ranged, sum ranged} per stream (wrapping) */
int  blast_x128p_fill_dev(blast_ctx* ctx, blast_x128p* d_states, uint64_t n_streams, uint64_t draws_per_stream,
                          int64_t lower, int64_t upper, uint64_t* d_raw, int64_t* d_ranged, uint64_t* d_checks);
/* host one-shot: seed -> jump -> fill -> copy back (outputs nullable) */
int  blast_x128p_fill(blast_ctx* ctx, uint64_t seed, uint64_t stride, uint64_t n_streams, uint64_t draws_per_stream,
                      int64_t lower, int64_t upper, uint64_t* raw_out, int64_t* ranged_out, uint64_t* checks_out);

/* ------------------------------------------------------------------ MPEG frame-sync scan / index
 * replaces file_parsing::mpeg::parse (blast/src/file_parsing/mpeg.rs:7-128). */
typedef struct {               /* parse_header + Header::format + compute_frame_len for one 32-bit header */
    uint32_t ok;               /* parse_header returned Ok (mpeg.rs:367-496) */
    uint32_t status;           /* otherwise BLAST_ERR_UNSUPPORTED_FORMAT / BLAST_ERR_INVALID_DATA */
    uint32_t version_id;       /* the 2-bit id as the reference computes it (low bit = protection bit) */
    uint32_t layer;            /* 1 / 2 / 3 */
    uint32_t is_protected;
    uint32_t padded;
    uint32_t channel_mode;
    uint32_t bitrate;          /* kbit/s, always column 4 of BITRATES (mpeg.rs:273-284) */
    double   sample_rate;
    uint32_t frame_len_ok;     /* compute_frame_len returned Ok (mpeg.rs:207-234) */
    uint32_t payload_len;      /* its value */
    uint32_t skip;             /* payload starts at pos + skip (6 if protected else 4, mpeg.rs:86-89) */
    uint32_t reserved;
} blast_mpeg_header;
int  blast_mpeg_header_info(uint32_t header, blast_mpeg_header* out);        /* host */

/* Greedy non-overlapping sync scan (mpeg.rs:17-50) over device-resident bytes (16-byte aligned):
 * candidates (position, big-endian header) in file order, WITHOUT the duplicate-first quirk.
 * *n_out is the number found even when it exceeds cap (then BLAST_ERR_CAPACITY).
 * BLAST_ERR_REF_PANIC if the scan reaches a trailing 0xFF (the reference indexes out of bounds). */
int  blast_mpeg_scan_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, uint64_t* d_pos_out,
                         uint32_t* d_hdr_out, uint64_t cap, uint64_t* n_out);
/* The frame-offset index = frames[*].file_pos after the sort (mpeg.rs:53-116): reference header = most
 * frequent header value that parses (ties: smallest value; the reference follows HashMap order), frames =
 * candidates whose header parses, match_ref()s and has a valid length.  reference_compat != 0 duplicates
 * the first position of every distinct header value (mpeg.rs:39) and reports payloads past EOF as
 * BLAST_ERR_REF_PANIC.  d_offsets_out nullable (count only). */
int  blast_mpeg_index_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, int reference_compat,
                          uint64_t* d_offsets_out, uint64_t cap, uint64_t* n_offsets_out, uint32_t* ref_header_out,
                          uint64_t* n_candidates_out);
/* The same in separate steps (blast_mpeg_index_dev = scan + these four).  The histogram has BLAST_MPEG_HDR_BINS
 * uint32 bins indexed by the low 21 header bits (the 11 sync bits are fixed) and is ACCUMULATED into (headers that
 * parse_header rejects cannot win the vote and are not counted: their bins stay as they were); the
 * first-position table has BLAST_MPEG_HDR_BINS uint64 entries, initialised by the caller to all-ones, and is
 * min-reduced into.  Both may be summed / min-reduced across GPUs between the steps (one NCCL all-reduce each). */
#define BLAST_MPEG_HDR_BINS (1u << 21)
int  blast_mpeg_hist_dev(blast_ctx* ctx, const uint32_t* d_hdr, uint64_t n, uint32_t* d_hist);                  /* async */
int  blast_mpeg_pick_ref_dev(blast_ctx* ctx, const uint32_t* d_hist, uint32_t* ref_header_out);                 /* mpeg.rs:53-73 */
int  blast_mpeg_first_pos_dev(blast_ctx* ctx, const uint64_t* d_pos, const uint32_t* d_hdr, uint64_t n, uint32_t ref_header,
                              uint64_t* d_first);                                                               /* async */
/* d_first nullable = no duplicate-first quirk; stream_len = length of the WHOLE stream (payload-past-EOF check) */
int  blast_mpeg_classify_dev(blast_ctx* ctx, const uint64_t* d_pos, const uint32_t* d_hdr, uint64_t n, uint32_t ref_header,
                             const uint64_t* d_first, uint64_t stream_len, uint64_t* d_offsets_out, uint64_t cap,
                             uint64_t* n_offsets_out);

/* Sharded scan (multi-GPU, SURVEY.md §8 e): one logical stream cut into contiguous byte ranges, one per GPU.  Every
 * range but the last must be a multiple of 32,768 bytes and is followed in d_bytes by halo_len >= 16 bytes of the next
 * range; the last range has halo_len 0.  Phase 1 returns the range's action on the greedy scan's 4-state machine:
 * exit state and candidate count for each of the 4 possible entry states.  The host folds the ranges in order
 * (entry state of range r = exit state of range r-1 under ITS entry state; range 0 enters in state 0) — an exchange
 * of 48 bytes per GPU — and phase 2 emits the range's candidates (positions + pos_offset) for its true entry state.
 * The two calls must follow each other on the same context (the walk's intermediate lists stay in its scratch). */
typedef struct { uint32_t exit_state[4]; uint64_t count[4]; } blast_mpeg_shard_agg;
int  blast_mpeg_shard_walk_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t own_len, uint64_t halo_len,
                               blast_mpeg_shard_agg* agg_out);
int  blast_mpeg_shard_emit_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t own_len, uint64_t halo_len,
                               uint32_t entry_state, uint64_t pos_offset, uint64_t* d_pos_out, uint32_t* d_hdr_out,
                               uint64_t cap, uint64_t* n_out);

/* Payload gather (mpeg.rs:86-121): concatenated b[pos+skip .. pos+skip+len] of the indexed frames.
 * d_payload_out nullable (size only). */
int  blast_mpeg_gather_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, const uint64_t* d_offsets,
                           uint64_t n_offsets, uint8_t* d_payload_out, uint64_t cap, uint64_t* payload_len_out);
/* mpeg::parse on a host buffer: upload + scan + index + gather; every output nullable except n_offsets_out */
int  blast_mpeg_parse(blast_ctx* ctx, const uint8_t* bytes, uint64_t len, int reference_compat, uint64_t* offsets_out,
                      uint64_t offsets_cap, uint64_t* n_offsets_out, uint32_t* ref_header_out, uint64_t* n_candidates_out,
                      uint8_t* payload_out, uint64_t payload_cap, uint64_t* payload_len_out);

/* ------------------------------------------------------------------ several GPUs, one host process
 * The reference is ONE process (blast/src/main.rs:13-128: decode every asset, then run the Conductor); a host like that
 * drives all GPUs of a box through a blast_group — no launcher, no collective library.  The path shards as
 * SURVEY.md §8(e) lays out:  file i is decoded on member i mod n and its track stays there;  a voice is rendered on the
 * member that holds its track, and the partial buses are reduced tile by tile inside the render kernel over peer
 * memory (blast_peer_bus);  RNG stream s is generated on member s mod n;  one MPEG stream is cut into byte ranges whose
 * 48-byte aggregates are folded on the host, with the header histogram and the first-position table reduced through
 * peer memory.  Results are identical to the single-GPU entry points.  A device id may repeat (at most 3 members per
 * GPU): the multi-member protocol can then be exercised on a single-GPU box — members that share a GPU wait for each
 * other on the device, so their streams need separate hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS=32 in the
 * environment before the first CUDA call). */
typedef struct blast_group blast_group;
int  blast_group_create(blast_group** out, const int* device_ids, uint32_t n_devices);
void blast_group_destroy(blast_group* g);
int  blast_group_set_fused(blast_group* g, int fused);   /* blast_group_render: see blast_peer_bus_set_fused */
uint32_t   blast_group_size(const blast_group* g);
blast_ctx* blast_group_ctx(blast_group* g, uint32_t member);
/* main.rs:18-89 over the group: like blast_pcm_decode_batch; tracks_out[i] (nullable array) is file i's AudioFile.samples
 * in the HBM of member i mod n, owned by the group until blast_group_free_tracks / blast_group_destroy. */
int  blast_group_pcm_decode_batch(blast_group* g, uint32_t n, const uint8_t* const* files, const size_t* lens,
                                  const blast_pcm_desc* descs, int16_t* const* host_out, blast_track* tracks_out);
int  blast_group_free_tracks(blast_group* g);
/* blast_render over the group: tracks[t] must live on member t mod n (as blast_group_pcm_decode_batch leaves them) */
int  blast_group_render(blast_group* g, const blast_track* tracks, uint32_t n_tracks, const blast_voice* voices,
                        uint32_t n_voices, uint32_t out_channels, uint64_t frames, int16_t* host_bus_out);
/* Conductor::{prepare, apply, coordinate} over the group (every member applies every command and tracks every tempo) */
typedef struct blast_group_conductor blast_group_conductor;
int  blast_group_conductor_create(blast_group* g, uint32_t out_channels, uint32_t sample_rate, const blast_track* tracks,
                                  uint32_t n_tracks, blast_group_conductor** out);
void blast_group_conductor_destroy(blast_group_conductor* gc);
int  blast_group_conductor_apply(blast_group_conductor* gc, const blast_command* cmd);
int  blast_group_conductor_coordinate(blast_group_conductor* gc, uint64_t frames, int16_t* host_bus_out);
blast_conductor* blast_group_conductor_member(blast_group_conductor* gc, uint32_t member);   /* state access (get / set_voice) */
/* blast_x128p_fill over the group: same outputs, stream s generated on member s mod n */
int  blast_group_x128p_fill(blast_group* g, uint64_t seed, uint64_t stride, uint64_t n_streams, uint64_t draws_per_stream,
                            int64_t lower, int64_t upper, uint64_t* raw_out, int64_t* ranged_out, uint64_t* checks_out);
/* the frame-offset index of mpeg::parse (mpeg.rs:7-116) for one host buffer cut into byte ranges over the members */
int  blast_group_mpeg_index(blast_group* g, const uint8_t* bytes, uint64_t len, int reference_compat, uint64_t* offsets_out,
                            uint64_t cap, uint64_t* n_offsets_out, uint32_t* ref_header_out, uint64_t* n_candidates_out);

#ifdef __cplusplus
}
#endif
#endif /* BLAST_CUDA_H */
