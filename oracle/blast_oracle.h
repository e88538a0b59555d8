/*
 * blast_oracle.h — CPU restatement of the BLAST hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for audio_decoder_b200.  It restates, in plain C++,
 * the arithmetic of the reference (gitxandert/audio_decoder, Rust) for exactly the
 * hot path named in BASELINE.json: WAV/AIFF PCM decode, the voice render/mix loop,
 * the xoroshiro128+ / Lemire streams and the MPEG frame-sync scan.  Every function
 * cites the reference file:line it follows (paths relative to the reference root,
 * `blast/src/...`).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  Nothing under audio_decoder_b200/ links or imports it.
 *
 * PARITY PINNING: the reference ships no golden vectors, no compiling tests and no
 * fixtures for this path (blast/src/lib.rs:7-35 does not compile; assets git-ignored),
 * and there is no Rust toolchain in the build image, so the reference cannot be run.
 * => "parity unpinned" by the reference's own artefacts.  The oracle is instead pinned
 * against (i) the known-answer vectors hand-derived from the reference source in
 * SURVEY.md §8(c), (ii) an independent second restatement in pure Python/numpy
 * (tests/pyref.py) and (iii) the published SplitMix64 first output.
 */
#ifndef BLAST_ORACLE_H
#define BLAST_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes: 0..4 mirror DecodeError (decode_helpers.rs:1-7); 5 = the reference would panic */
enum {
    ORC_OK = 0,
    ORC_IO = 1,
    ORC_UNSUPPORTED_FORMAT = 2,
    ORC_UNEXPECTED_EOF = 3,
    ORC_INVALID_DATA = 4,
    ORC_REF_PANIC = 5,
    ORC_BAD_ARG = 101
};

const char* orc_last_error(void);

/* ---------------- L0: PCM decode ---------------- */
typedef struct {
    uint32_t sample_rate;
    uint32_t num_channels;
    uint32_t bits_per_sample;
    uint32_t big_endian;   /* 0 = wav (LE pairs), 1 = aiff (BE pairs) */
    uint64_t data_off;     /* byte offset of the first sample byte */
    uint64_t data_len;     /* declared payload length in bytes (data_size / ssnd_size) */
} orc_pcm_desc;

/* header walks: wav.rs:69-138, aiff.rs:99-154 */
int orc_wav_probe(const uint8_t* file, size_t len, orc_pcm_desc* out);
int orc_aiff_probe(const uint8_t* file, size_t len, orc_pcm_desc* out);
/* 80-bit extended -> f64: aiff.rs:51-94 */
double orc_ieee_extended(const uint8_t bytes[10]);
/* f64 -> u32 Rust `as` cast (aiff.rs:182) */
uint32_t orc_f64_as_u32(double x);

/* full parse, faithful loop structure (per-pair bounds-checked reads, growth without
 * reserve): wav.rs:140-154, aiff.rs:156-170.  On success *samples_out is a buffer owned
 * by the oracle (release with orc_free). */
int orc_wav_parse(const uint8_t* file, size_t len, orc_pcm_desc* desc, int16_t** samples_out, size_t* n_out);
int orc_aiff_parse(const uint8_t* file, size_t len, orc_pcm_desc* desc, int16_t** samples_out, size_t* n_out);
/* "good CPU" variant: memcpy / bswap into a caller buffer of orc_pcm_out_len() words */
int orc_pcm_decode_fast(const uint8_t* file, size_t len, const orc_pcm_desc* desc, int16_t* out);
size_t orc_pcm_out_len(const orc_pcm_desc* desc);
/* extension oracle (not in the reference): true 24-bit unpack, 3 bytes -> sign-extended i32 */
void orc_pcm24_unpack(const uint8_t* payload, size_t n_samples, int big_endian, int32_t* out);
void orc_free(void* p);

/* file-name rule: wav.rs:156-164 / aiff.rs:172-180.  Writes a NUL-terminated name. */
int orc_file_name(const char* path, char* out, size_t cap);

/* ---------------- RNG: blast_rand.rs:4-60 ---------------- */
typedef struct { uint64_t s0, s1; } orc_x128p;
void     orc_x128p_new(uint64_t seed, orc_x128p* out);
uint64_t orc_x128p_next_u64(orc_x128p* g);
double   orc_x128p_next_f64(orc_x128p* g);
float    orc_x128p_next_f32(orc_x128p* g);
int64_t  orc_x128p_next_i64_range(orc_x128p* g, int64_t lower, int64_t upper);
void     orc_x128p_fill_u64(orc_x128p* g, uint64_t n, uint64_t* out);
void     orc_x128p_fill_range(orc_x128p* g, int64_t lower, int64_t upper, uint64_t n, int64_t* out);
/* advance by n draws, sequentially (ground truth for jump-ahead) */
void     orc_x128p_discard(orc_x128p* g, uint64_t n);
/* checksums of n draws: xor / wrapping sum of raw u64 and of the ranged i64 values */
void     orc_x128p_checksum(orc_x128p* g, int64_t lower, int64_t upper, uint64_t n,
                            uint64_t* raw_xor, uint64_t* raw_sum, uint64_t* rng_xor, uint64_t* rng_sum);

/* ---------------- tempo: blast_time.rs:58-161 ---------------- */
enum { ORC_TM_PROCESS = 0, ORC_TM_VOICE = 1, ORC_TM_GROUP = 2, ORC_TM_CONTEXT = 3, ORC_TM_TBD = 4 };
enum { ORC_TU_SAMPLES = 0, ORC_TU_MILLIS = 1, ORC_TU_BPM = 2 };
float orc_convert_interval(uint32_t sample_rate, uint32_t unit, float interval);

/* ---------------- L1: the Conductor (engine.rs) ---------------- */
typedef struct {
    const int16_t* samples;  /* interleaved */
    uint64_t n_samples;
    uint32_t num_channels;
    uint32_t sample_rate;
} orc_track;

typedef struct {             /* commands.rs:187-234 TempoRepr */
    uint64_t idx;
    uint32_t owned;
    uint32_t mode;
    uint32_t unit;
    float    interval;
} orc_tempo_repr;

enum { ORC_CMD_LOAD = 0, ORC_CMD_START, ORC_CMD_PAUSE, ORC_CMD_RESUME, ORC_CMD_STOP, ORC_CMD_UNLOAD,
       ORC_CMD_VELOCITY, ORC_CMD_GROUP, ORC_CMD_TC, ORC_CMD_SEQ, ORC_CMD_QUIT };
enum { ORC_IDX_TEMPO = 0, ORC_IDX_VOICE = 1, ORC_IDX_PROCESS = 2, ORC_IDX_GROUP = 3 };

typedef struct {             /* commands.rs:86-161, flattened */
    uint32_t kind;
    uint32_t idx_kind;       /* Idx variant for Start/Pause/Resume/Stop/Seq */
    uint64_t idx;            /* Idx payload; track_idx for Load; voice idx for Unload/Velocity */
    float    val;            /* VelocityArgs.val */
    orc_tempo_repr tempo;    /* Load.tempo_repr / Group.tempo / Tc.tempo / Seq.tempo */
    /* GroupArgs.vs_fs_ps: member i = (voice idx, update_tempo, proc ids) */
    uint32_t n_members;
    const uint64_t* member_voice;
    const uint8_t*  member_update_tempo;
    const uint32_t* member_n_procs;
    const uint64_t* member_proc_ids;   /* concatenated */
    /* SeqArgs */
    uint64_t period;
    uint32_t n_steps;
    const float* steps;
    const float* chance;
    uint64_t rng_s0, rng_s1;
} orc_command;

typedef struct orc_conductor orc_conductor;

/* Conductor::prepare (engine.rs:36-44) + sample_rate::set (runtime.rs:37) */
orc_conductor* orc_conductor_new(uint32_t out_channels, uint32_t sample_rate, const orc_track* tracks, uint32_t n_tracks);
void orc_conductor_free(orc_conductor*);
/* Conductor::apply (engine.rs:83-248).  Returns ORC_REF_PANIC where the reference would panic. */
int  orc_conductor_apply(orc_conductor*, const orc_command*);
/* Conductor::coordinate (engine.rs:46-81) into an interleaved S16 bus [frames x out_channels] */
int  orc_conductor_coordinate(orc_conductor*, uint64_t frames, int16_t* bus);

/* direct field access (VoiceState fields are pub: engine.rs:279-286).  `group` = -1 for
 * Conductor.voices, else index into Conductor.groups. */
typedef struct {
    uint32_t active;
    float    position;
    float    velocity;
    float    gain;
    uint64_t end;
    uint32_t channels;
    uint32_t tempo_current;
    uint32_t tempo_active;
} orc_voice_state;
int orc_conductor_n_voices(orc_conductor*, int group);
int orc_conductor_n_groups(orc_conductor*);
int orc_conductor_get_voice(orc_conductor*, int group, uint32_t idx, orc_voice_state* out);
int orc_conductor_set_voice(orc_conductor*, int group, uint32_t idx, const float* position, const float* velocity, const float* gain, const int* active);
uint64_t orc_clock_current(orc_conductor*);
/* position recurrence alone (engine.rs:407-410,445-447): out[0..n], freeze at trunc(pos) >= end */
void orc_position_walk(float p0, float velocity, uint64_t end, uint64_t n, float* out);

/* ---------------- MPEG: mpeg.rs ---------------- */
typedef struct {
    uint8_t  ok;            /* 1 = parse_header returned Ok */
    uint8_t  err;           /* ORC_* status if !ok */
    uint8_t  version_id;    /* raw 2-bit value (mpeg.rs:377-383) */
    uint8_t  layer_id;      /* raw 2-bit value */
    uint8_t  not_protected;
    uint8_t  padded;
    uint8_t  channel_mode;
    uint8_t  frame_len_ok;  /* compute_frame_len returned Ok */
    uint32_t bitrate;
    double   sr;
    float    version;       /* Header::format */
    int32_t  layer;
    uint64_t payload_len;   /* compute_frame_len (mpeg.rs:207-234) */
    uint32_t skip;          /* 6 if protected else 4 (mpeg.rs:86-89) */
} orc_mpeg_header;

void orc_mpeg_parse_header(uint32_t header, orc_mpeg_header* out);
int  orc_mpeg_match_ref(const orc_mpeg_header* ref, const orc_mpeg_header* other);

/* sync scan, literal (mpeg.rs:17-50).  Outputs candidate (pos, header) pairs in scan order
 * (without the duplicate-first quirk).  Returns ORC_REF_PANIC if the reference would index
 * out of bounds (last byte 0xFF). */
int orc_mpeg_sync_scan(const uint8_t* bytes, uint64_t len, uint64_t* pos_out, uint32_t* hdr_out, uint64_t cap, uint64_t* n_out);

/* the full mpeg::parse (mpeg.rs:7-128).  tie_break: when several headers share the top
 * count the reference follows HashMap order (nondeterministic); the oracle picks the
 * smallest header value.  offsets_out receives frames[*].file_pos after the sort
 * (first position of each header duplicated when reference_compat != 0); payload_out (nullable)
 * receives the concatenated payload. */
int orc_mpeg_parse(const uint8_t* bytes, uint64_t len, int reference_compat,
                   uint64_t* offsets_out, uint64_t offsets_cap, uint64_t* n_offsets,
                   uint32_t* ref_header_out, uint64_t* n_candidates_out,
                   uint8_t* payload_out, uint64_t payload_cap, uint64_t* payload_len_out);

#ifdef __cplusplus
}
#endif
#endif
