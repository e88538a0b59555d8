// blast_oracle.cpp — CPU restatement of the BLAST hot path.  TEST INFRASTRUCTURE ONLY
// (see blast_oracle.h for the rules and the "parity unpinned" statement).
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC  (see oracle/Makefile).
// -ffp-contract=off matters: Rust never fuses a*b+c (engine.rs:435).
//
// Reference citations are relative to /root/reference/blast/src/.

#include "blast_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char* msg) {
    g_err = msg;
    return code;
}

// ---- Rust `as` casts (saturating, NaN -> 0, truncating toward zero) ----
inline int16_t f32_as_i16(float x) {
    if (x != x) return 0;
    if (x >= 32767.0f) return 32767;
    if (x <= -32768.0f) return -32768;
    return (int16_t)(int32_t)x;
}
inline uint64_t f32_as_usize(float x) {
    if (x != x) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)x;
}
inline int64_t f32_as_i64(float x) {
    if (x != x) return 0;
    if (x >= 9223372036854775808.0f) return INT64_MAX;
    if (x <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)x;
}
inline uint32_t f64_as_u32(double x) {
    if (x != x) return 0;
    if (x <= 0.0) return 0;
    if (x >= 4294967295.0) return UINT32_MAX;
    return (uint32_t)x;
}

// ---- byte cursor shared by the two header walks ----
struct Cursor {
    const uint8_t* b;
    size_t len;
    size_t start = 0, end = 0;
    // wav.rs:30-44 / aiff.rs:6-23: advance 4, bounds-check each byte, never compare the id
    bool skip_id() {
        end += 4;
        for (size_t i = start; i < end; ++i)
            if (i >= len) return false;
        start = end;
        return true;
    }
    // wav.rs:46-67 (little-endian accumulate)
    bool le(size_t inc, uint32_t* out) {
        uint32_t value = 0, shift = 0;
        end += inc;
        for (size_t i = start; i < end; ++i) {
            if (i >= len) return false;
            value += (uint32_t)b[i] << shift;
            shift += 8;
        }
        start = end;
        *out = value;
        return true;
    }
    // aiff.rs:25-48 (big-endian accumulate)
    bool be(size_t inc, uint32_t* out) {
        uint32_t value = 0, shift = 8 * (uint32_t)inc - 8;
        end += inc;
        for (size_t i = start; i < end; ++i) {
            if (i >= len) return false;
            value += (uint32_t)b[i] << shift;
            if (shift >= 8) shift -= 8;
        }
        start = end;
        *out = value;
        return true;
    }
};

// compiler-rt __powidf2, which is what f64::powi lowers to (aiff.rs:90)
double powi_f64(double a, int b) {
    const bool recip = b < 0;
    double r = 1.0;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

// payload bounds rule shared by wav.rs:143-151 and aiff.rs:159-167: pairs (i, i+1) for
// i = start, start+2, ... < start+n must all exist, otherwise the whole parse is EOF.
bool payload_in_bounds(size_t len, uint64_t off, uint64_t n) {
    if (n == 0) return true;
    uint64_t need = off + 2 * ((n + 1) / 2);
    return need <= len;
}

}  // namespace

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }
void orc_free(void* p) { std::free(p); }

uint32_t orc_f64_as_u32(double x) { return f64_as_u32(x); }

// wav.rs:69-138
int orc_wav_probe(const uint8_t* file, size_t len, orc_pcm_desc* out) {
    Cursor c{file, len};
    uint32_t riff_size, fmt_size, tag, ch, rate, data_rate, blk, bits, data_size;
    if (!c.skip_id()) return fail(ORC_UNEXPECTED_EOF, "eof");          // "RIFF"
    if (!c.le(4, &riff_size)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.skip_id()) return fail(ORC_UNEXPECTED_EOF, "eof");          // "WAVE"
    if (!c.skip_id()) return fail(ORC_UNEXPECTED_EOF, "eof");          // "fmt "
    if (!c.le(4, &fmt_size)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.le(2, &tag)) return fail(ORC_UNEXPECTED_EOF, "eof");
    // wav.rs:17-28, 95-98
    if (!(tag == 0x0001 || tag == 0x0003 || tag == 0x0006 || tag == 0x0007 || tag == 0xFFFE))
        return fail(ORC_UNSUPPORTED_FORMAT, "Unrecognized format tag");
    if (!c.le(2, &ch)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.le(4, &rate)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.le(4, &data_rate)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.le(2, &blk)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.le(2, &bits)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (fmt_size >= 18) {                                              // wav.rs:112-130
        uint32_t cb, valid_bits, mask, old_fmt;
        if (!c.le(2, &cb)) return fail(ORC_UNEXPECTED_EOF, "eof");
        if (cb > 0) {
            if (!c.le(2, &valid_bits)) return fail(ORC_UNEXPECTED_EOF, "eof");
            if (!c.le(4, &mask)) return fail(ORC_UNEXPECTED_EOF, "eof");
            if (!c.le(2, &old_fmt)) return fail(ORC_UNEXPECTED_EOF, "eof");
            for (size_t i = 0; i < 14; ++i) c.end += i;               // wav.rs:124-127: +91, not +14
            c.start = c.end;
        }
    }
    if (!c.skip_id()) return fail(ORC_UNEXPECTED_EOF, "eof");          // "data"
    if (!c.le(4, &data_size)) return fail(ORC_UNEXPECTED_EOF, "eof");
    out->sample_rate = rate;
    out->num_channels = ch;
    out->bits_per_sample = bits;
    out->big_endian = 0;
    out->data_off = c.start;
    out->data_len = data_size;
    if (!payload_in_bounds(len, out->data_off, out->data_len)) return fail(ORC_UNEXPECTED_EOF, "eof in data");
    return ORC_OK;
}

// aiff.rs:51-94
double orc_ieee_extended(const uint8_t bytes[10]) {
    bool sign = (bytes[0] & 0x80) != 0;
    uint16_t exp = (uint16_t)(((bytes[0] & 0x7F) << 8) | bytes[1]);
    uint64_t mant = 0;
    for (int i = 2; i < 10; ++i) mant = (mant << 8) | bytes[i];
    if (exp == 0 && mant == 0) return 0.0;
    if (exp == 0x7FFF) {
        if (mant == 0) return sign ? -INFINITY : INFINITY;
        return NAN;
    }
    int e = (int)exp - 16383 - 63;
    double val = (double)mant * powi_f64(2.0, e);
    if (sign) val = -val;
    return val;
}

// aiff.rs:99-154
int orc_aiff_probe(const uint8_t* file, size_t len, orc_pcm_desc* out) {
    Cursor c{file, len};
    uint32_t form_size, comm_size, ch, frames, sample_size, ssnd, offset, block;
    if (!c.skip_id()) return fail(ORC_UNEXPECTED_EOF, "eof");          // "FORM"
    if (!c.be(4, &form_size)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.skip_id()) return fail(ORC_UNEXPECTED_EOF, "eof");          // "AIFF"
    if (!c.skip_id()) return fail(ORC_UNEXPECTED_EOF, "eof");          // "COMM"
    if (!c.be(4, &comm_size)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (comm_size != 18) return fail(ORC_INVALID_DATA, "Comm size should be 18");
    if (!c.be(2, &ch)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.be(4, &frames)) return fail(ORC_UNEXPECTED_EOF, "eof");
    if (!c.be(2, &sample_size)) return fail(ORC_UNEXPECTED_EOF, "eof");
    // parse_ieee_extended reads 10 bytes with bounds checks (aiff.rs:52-63)
    uint8_t ext[10];
    c.end += 10;
    for (size_t i = c.start, k = 0; i < c.end; ++i, ++k) {
        if (i >= len) return fail(ORC_UNEXPECTED_EOF, "eof");
        ext[k] = file[i];
    }
    c.start = c.end;
    double rate = orc_ieee_extended(ext);
    if (!c.skip_id()) return fail(ORC_UNEXPECTED_EOF, "eof");          // "SSND"
    if (!c.be(4, &ssnd)) return fail(ORC_UNEXPECTED_EOF, "eof");
    uint32_t ssnd_size = ssnd - 8u;   // aiff.rs:146: u32 subtraction; release builds wrap (debug panics)
    if (!c.be(4, &offset)) return fail(ORC_UNEXPECTED_EOF, "eof");     // read and ignored
    if (!c.be(4, &block)) return fail(ORC_UNEXPECTED_EOF, "eof");      // read and ignored
    out->sample_rate = f64_as_u32(rate);                               // aiff.rs:182
    out->num_channels = ch;
    out->bits_per_sample = sample_size;
    out->big_endian = 1;
    out->data_off = c.start;
    out->data_len = ssnd_size;
    if (!payload_in_bounds(len, out->data_off, out->data_len)) return fail(ORC_UNEXPECTED_EOF, "eof in ssnd");
    return ORC_OK;
}

size_t orc_pcm_out_len(const orc_pcm_desc* d) { return (size_t)((d->data_len + 1) / 2); }

// A growable i16 buffer that grows like Vec::push without with_capacity (amortised doubling,
// realloc + copy) — the reference's allocation pattern (wav.rs:140,153).
namespace {
struct GrowVec {
    int16_t* p = nullptr;
    size_t n = 0, cap = 0;
    inline void push(int16_t v) {
        if (n == cap) {
            size_t nc = cap ? cap * 2 : 4;
            p = (int16_t*)std::realloc(p, nc * sizeof(int16_t));
            cap = nc;
        }
        p[n++] = v;
    }
};

int parse_samples(const uint8_t* file, size_t len, const orc_pcm_desc* d, int16_t** out, size_t* n_out) {
    GrowVec v;
    size_t start = d->data_off, end = d->data_off + d->data_len;
    const bool be = d->big_endian != 0;
    for (size_t i = start; i < end; i += 2) {          // wav.rs:143 / aiff.rs:159
        if (i >= len) { std::free(v.p); return fail(ORC_UNEXPECTED_EOF, "eof"); }
        uint8_t s1 = file[i];
        if (i + 1 >= len) { std::free(v.p); return fail(ORC_UNEXPECTED_EOF, "eof"); }
        uint8_t s2 = file[i + 1];
        uint16_t w = be ? (uint16_t)((s1 << 8) | s2) : (uint16_t)((s2 << 8) | s1);
        v.push((int16_t)w);
    }
    *out = v.p;
    *n_out = v.n;
    return ORC_OK;
}
}  // namespace

int orc_wav_parse(const uint8_t* file, size_t len, orc_pcm_desc* desc, int16_t** samples_out, size_t* n_out) {
    // the probe's early payload bounds check only changes WHEN EOF is reported, not whether
    int rc = orc_wav_probe(file, len, desc);
    if (rc != ORC_OK) return rc;
    return parse_samples(file, len, desc, samples_out, n_out);
}

int orc_aiff_parse(const uint8_t* file, size_t len, orc_pcm_desc* desc, int16_t** samples_out, size_t* n_out) {
    int rc = orc_aiff_probe(file, len, desc);
    if (rc != ORC_OK) return rc;
    return parse_samples(file, len, desc, samples_out, n_out);
}

int orc_pcm_decode_fast(const uint8_t* file, size_t len, const orc_pcm_desc* d, int16_t* out) {
    if (!payload_in_bounds(len, d->data_off, d->data_len)) return fail(ORC_UNEXPECTED_EOF, "eof");
    size_t n = orc_pcm_out_len(d);
    const uint8_t* src = file + d->data_off;
    if (!d->big_endian) {
        std::memcpy(out, src, n * 2);
    } else {
        for (size_t i = 0; i < n; ++i) {
            uint16_t w;
            std::memcpy(&w, src + 2 * i, 2);
            out[i] = (int16_t)__builtin_bswap16(w);
        }
    }
    return ORC_OK;
}

void orc_pcm24_unpack(const uint8_t* p, size_t n, int big_endian, int32_t* out) {
    for (size_t i = 0; i < n; ++i) {
        uint32_t b0 = p[3 * i], b1 = p[3 * i + 1], b2 = p[3 * i + 2];
        uint32_t u = big_endian ? (b0 << 24) | (b1 << 16) | (b2 << 8) : (b2 << 24) | (b1 << 16) | (b0 << 8);
        out[i] = (int32_t)u >> 8;
    }
}

// wav.rs:156-164 / aiff.rs:172-180
int orc_file_name(const char* path, char* out, size_t cap) {
    std::string p(path);
    size_t dot = p.rfind('.');
    if (dot == std::string::npos) return fail(ORC_INVALID_DATA, "File has no name");
    std::string before = p.substr(0, dot), after = p.substr(dot + 1);
    if (before.empty() || after.empty()) return fail(ORC_INVALID_DATA, "File has no name");
    size_t slash = before.rfind('/');
    if (slash == std::string::npos) return fail(ORC_INVALID_DATA, "File is not nested");
    std::string name = before.substr(slash + 1);
    if (name.size() + 1 > cap) return fail(ORC_BAD_ARG, "name buffer too small");
    std::memcpy(out, name.c_str(), name.size() + 1);
    return ORC_OK;
}

// ===================== RNG: blast_rand.rs =====================
namespace {
inline uint64_t splitmix64(uint64_t x) {               // blast_rand.rs:12-18
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t rotl(uint64_t x, unsigned k) { return (x << k) | (x >> (64 - k)); }
inline uint64_t x_next(orc_x128p* g) {                  // blast_rand.rs:31-39
    uint64_t result = g->s0 + g->s1;
    uint64_t s1 = g->s1 ^ g->s0;
    g->s0 = rotl(g->s0, 55) ^ s1 ^ (s1 << 14);
    g->s1 = rotl(s1, 36);
    return result;
}
inline int64_t x_range(orc_x128p* g, int64_t lower, int64_t upper) {   // blast_rand.rs:50-59
    uint64_t r = x_next(g);
    uint64_t range = upper > lower ? (uint64_t)(upper - lower) : (uint64_t)(lower - upper);
    int64_t val = (int64_t)(uint64_t)(((unsigned __int128)r * (unsigned __int128)range) >> 64);
    return (int64_t)((uint64_t)lower + (uint64_t)val);
}
}  // namespace

void orc_x128p_new(uint64_t seed, orc_x128p* out) {     // blast_rand.rs:10-24
    out->s0 = splitmix64(seed);
    out->s1 = splitmix64(seed + 0x9E3779B97F4A7C15ull);
}
uint64_t orc_x128p_next_u64(orc_x128p* g) { return x_next(g); }
double orc_x128p_next_f64(orc_x128p* g) {               // blast_rand.rs:41-44
    return (double)(x_next(g) >> 11) * (1.0 / (double)(1ull << 53));
}
float orc_x128p_next_f32(orc_x128p* g) { return (float)orc_x128p_next_f64(g); }   // blast_rand.rs:46-48
int64_t orc_x128p_next_i64_range(orc_x128p* g, int64_t lo, int64_t hi) { return x_range(g, lo, hi); }
void orc_x128p_fill_u64(orc_x128p* g, uint64_t n, uint64_t* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = x_next(g);
}
void orc_x128p_fill_range(orc_x128p* g, int64_t lo, int64_t hi, uint64_t n, int64_t* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = x_range(g, lo, hi);
}
void orc_x128p_discard(orc_x128p* g, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) (void)x_next(g);
}
void orc_x128p_checksum(orc_x128p* g, int64_t lo, int64_t hi, uint64_t n,
                        uint64_t* raw_xor, uint64_t* raw_sum, uint64_t* rng_xor, uint64_t* rng_sum) {
    uint64_t rx = 0, rs = 0, gx = 0, gs = 0;
    uint64_t range = hi > lo ? (uint64_t)(hi - lo) : (uint64_t)(lo - hi);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t r = x_next(g);
        uint64_t v = (uint64_t)lo + (uint64_t)(((unsigned __int128)r * (unsigned __int128)range) >> 64);
        rx ^= r; rs += r; gx ^= v; gs += v;
    }
    *raw_xor = rx; *raw_sum = rs; *rng_xor = gx; *rng_sum = gs;
}

// ===================== tempo: blast_time.rs =====================
float orc_convert_interval(uint32_t sample_rate, uint32_t unit, float interval) {   // blast_time.rs:151-161
    float frac;
    switch (unit) {
        case ORC_TU_SAMPLES: return interval;
        case ORC_TU_MILLIS: frac = interval / 1000.0f; break;
        default: frac = 60.0f / interval; break;
    }
    return (float)sample_rate * frac;
}

}  // extern "C"

// ===================== L1: Conductor / Voice / Group / Seq =====================
namespace {

struct TempoState {            // blast_time.rs:58-65
    int mode;
    int unit;
    float interval;
    bool active;
    uint32_t current;
    void reset() { current = 0; }
    void start() { reset(); active = true; }              // blast_time.rs:123-126
    void stop() { active = false; reset(); }              // blast_time.rs:136-139
    void update(double delta) { current += f64_as_u32(delta); }   // blast_time.rs:113-115 (wrapping)
    float cur() const { return (float)current / interval; }       // blast_time.rs:118-121
};
using TempoRef = std::shared_ptr<TempoState>;

struct VoiceState {            // engine.rs:279-286
    bool active;
    float position;
    uint64_t end;
    float velocity;
    float gain;
    TempoRef tempo;
};

struct Seq {                   // processes.rs:52-99
    bool active;
    TempoRef tempo;
    uint64_t period;
    std::vector<float> steps, chance;
    orc_x128p rng;
    uint64_t idx;
    // returns false where the reference would panic (index out of bounds)
    bool process(VoiceState& voice) {
        if (!active) return true;
        const TempoState& t = *tempo;
        if (!t.active) return true;
        float current = std::fmod(t.cur(), (float)period);
        if (idx >= steps.size()) return false;
        if (current == steps[idx]) {
            int64_t rand = x_range(&rng, 0, 100);
            if (idx >= chance.size()) return false;
            if (rand < f32_as_i64(chance[idx])) {
                voice.position = voice.velocity >= 0.0f ? 0.0f : (float)voice.end;
            }
            idx += 1;
            idx %= steps.size();
        }
        return true;
    }
    void reset() { idx = 0; }
};

struct Voice {                 // engine.rs:288-295
    const int16_t* samples;
    uint64_t n_samples;
    uint32_t sample_rate;
    uint64_t channels;
    VoiceState state;
    std::vector<Seq> processes;
    std::vector<TempoRef> proc_tempi;

    void start() {             // engine.rs:318-344
        state.active = true;
        for (auto& p : processes) p.reset();
        TempoState& ts = *state.tempo;
        if (ts.mode == ORC_TM_VOICE || ts.mode == ORC_TM_TBD) ts.start();
        for (auto& t : proc_tempi) t->start();
        state.position = state.velocity >= 0.0f ? 0.0f : (float)state.end;
    }
    void pause() { state.active = false; }                 // engine.rs:346-348
    void resume() { state.active = true; }                 // engine.rs:350-359
    void stop() {              // engine.rs:361-384
        state.active = false;
        for (auto& p : processes) p.reset();
        TempoState& ts = *state.tempo;
        if (ts.mode == ORC_TM_VOICE) ts.stop();
        for (auto& t : proc_tempi) { t->active = false; t->reset(); }
        state.position = state.velocity >= 0.0f ? 0.0f : (float)state.end;
    }
    // engine.rs:386-448
    inline bool process(int16_t* acc, uint64_t ch) {
        if (!state.active) return true;
        for (auto& p : processes)
            if (!p.process(state)) return false;
        TempoState& own = *state.tempo;
        if (own.mode == ORC_TM_VOICE || own.mode == ORC_TM_TBD) own.update(1.0);
        for (auto& t : proc_tempi) t->update(1.0);

        uint64_t idx = f32_as_usize(state.position);
        if (idx >= state.end) return true;

        if (channels == 1) {
            if (ch < 2) ch = 0; else return true;
        } else if (ch >= channels) {
            return true;
        }
        float sample;
        float s0 = (float)samples[idx * channels + (ch % channels)];
        if (state.velocity != 1.0f) {
            float frac = state.position - std::trunc(state.position);   // f32::fract
            float s1 = (float)samples[(idx + 1) * channels + (ch % channels)];
            float a = s0 * (1.0f - frac);
            float b = s1 * frac;
            sample = a + b;
        } else {
            sample = s0;
        }
        float scaled = sample * state.gain;
        *acc = (int16_t)((uint16_t)*acc + (uint16_t)f32_as_i16(scaled));   // release-mode wrapping add
        if (ch == channels - 1) state.position += state.velocity;
        return true;
    }
};

struct Group {                 // engine.rs:451-542
    bool active;
    float gain;
    TempoRef tempo;
    std::vector<Voice> voices;
    std::vector<Seq> processes;   // stored, never run (engine.rs:530-542)
    void start() {
        active = true;
        TempoState& ts = *tempo;
        if (ts.mode == ORC_TM_GROUP) { ts.active = true; ts.reset(); }
        for (auto& v : voices) v.start();
    }
    void pause() { active = false; }
    void resume() { active = true; }
    void stop() {
        active = false;
        for (auto& v : voices) v.state.active = false;
        TempoState& ts = *tempo;
        if (ts.mode == ORC_TM_GROUP) { ts.active = false; ts.reset(); }
    }
    inline bool process(int16_t* acc, uint64_t ch) {
        if (!active) return true;
        for (auto& v : voices)
            if (!v.process(acc, ch)) return false;
        TempoState& ts = *tempo;
        if (ts.mode == ORC_TM_GROUP) ts.update(1.0);
        return true;
    }
};

}  // namespace

struct orc_conductor {
    std::vector<Voice> voices;
    std::vector<Group> groups;
    std::vector<TempoRef> tempo_cons;
    uint64_t out_channels;
    std::vector<orc_track> tracks;
    uint32_t sample_rate;
    uint64_t clock = 0;

    TempoRef tempo_new_default() {          // blast_time.rs:84-97 with None
        auto t = std::make_shared<TempoState>();
        t->mode = ORC_TM_TBD;
        t->unit = ORC_TU_SAMPLES;
        t->interval = (float)sample_rate;
        t->active = false;
        t->current = 0;
        return t;
    }
    // engine.rs:252-275; returns nullptr where the reference would panic
    TempoRef tempo_from_repr(const orc_tempo_repr& tr) {
        TempoRef tempo = tempo_new_default();
        if (tr.owned) {
            tempo->interval = orc_convert_interval(sample_rate, tr.unit, tr.interval);   // blast_time.rs:99-104
            tempo->mode = (int)tr.mode;
            tempo->unit = (int)tr.unit;
        } else {
            switch (tr.mode) {
                case ORC_TM_VOICE:
                    if (tr.idx >= voices.size()) return nullptr;
                    tempo = voices[tr.idx].state.tempo;
                    break;
                case ORC_TM_GROUP:
                    if (tr.idx >= groups.size()) return nullptr;
                    tempo = groups[tr.idx].tempo;
                    break;
                case ORC_TM_CONTEXT:
                    if (tr.idx >= tempo_cons.size()) return nullptr;
                    tempo = tempo_cons[tr.idx];
                    break;
                default: break;
            }
        }
        return tempo;
    }
};

extern "C" {

orc_conductor* orc_conductor_new(uint32_t out_channels, uint32_t sample_rate, const orc_track* tracks, uint32_t n_tracks) {
    auto* c = new orc_conductor();
    c->out_channels = out_channels;
    c->sample_rate = sample_rate;
    c->tracks.assign(tracks, tracks + n_tracks);
    return c;
}
void orc_conductor_free(orc_conductor* c) { delete c; }

int orc_conductor_apply(orc_conductor* c, const orc_command* cmd) {
    const char* oob = "index out of bounds (reference panics)";
    switch (cmd->kind) {
        case ORC_CMD_LOAD: {                                 // engine.rs:103-107, 298-316
            if (cmd->idx >= c->tracks.size()) return fail(ORC_REF_PANIC, oob);
            const orc_track& tr = c->tracks[cmd->idx];
            TempoRef tempo = c->tempo_from_repr(cmd->tempo);
            if (!tempo) return fail(ORC_REF_PANIC, oob);
            if (tr.num_channels == 0) return fail(ORC_REF_PANIC, "divide by zero channels");
            uint64_t frames = tr.n_samples / tr.num_channels;
            if (frames == 0) return fail(ORC_REF_PANIC, "usize underflow in Voice::new (empty track)");
            Voice v;
            v.samples = tr.samples;
            v.n_samples = tr.n_samples;
            v.sample_rate = tr.sample_rate;
            v.channels = tr.num_channels;
            v.state = VoiceState{false, 0.0f, frames - 1, 1.0f, 1.0f, tempo};
            c->voices.push_back(std::move(v));
            return ORC_OK;
        }
        case ORC_CMD_START: case ORC_CMD_PAUSE: case ORC_CMD_RESUME: case ORC_CMD_STOP: {   // engine.rs:110-180
            int k = cmd->kind;
            if (cmd->idx_kind == ORC_IDX_VOICE) {
                if (cmd->idx >= c->voices.size()) return fail(ORC_REF_PANIC, oob);
                Voice& v = c->voices[cmd->idx];
                if (k == ORC_CMD_START) v.start(); else if (k == ORC_CMD_PAUSE) v.pause();
                else if (k == ORC_CMD_RESUME) v.resume(); else v.stop();
            } else if (cmd->idx_kind == ORC_IDX_GROUP) {
                if (cmd->idx >= c->groups.size()) return fail(ORC_REF_PANIC, oob);
                Group& g = c->groups[cmd->idx];
                if (k == ORC_CMD_START) g.start(); else if (k == ORC_CMD_PAUSE) g.pause();
                else if (k == ORC_CMD_RESUME) g.resume(); else g.stop();
            } else if (cmd->idx_kind == ORC_IDX_TEMPO) {
                if (cmd->idx >= c->tempo_cons.size()) return fail(ORC_REF_PANIC, oob);
                TempoState& t = *c->tempo_cons[cmd->idx];
                if (k == ORC_CMD_START) t.start(); else if (k == ORC_CMD_PAUSE) t.active = false;
                else if (k == ORC_CMD_RESUME) t.active = true; else t.stop();
            }
            return ORC_OK;
        }
        case ORC_CMD_UNLOAD:                                  // engine.rs:182-184
            if (cmd->idx >= c->voices.size()) return fail(ORC_REF_PANIC, oob);
            c->voices.erase(c->voices.begin() + (ptrdiff_t)cmd->idx);
            return ORC_OK;
        case ORC_CMD_VELOCITY:                                // engine.rs:186-189
            if (cmd->idx >= c->voices.size()) return fail(ORC_REF_PANIC, oob);
            c->voices[cmd->idx].state.velocity = cmd->val;
            return ORC_OK;
        case ORC_CMD_GROUP: {                                 // engine.rs:191-212
            TempoRef tempo = c->tempo_from_repr(cmd->tempo);
            if (!tempo) return fail(ORC_REF_PANIC, oob);
            Group g;
            g.active = false; g.gain = 1.0f; g.tempo = tempo;
            size_t pcur = 0;
            for (uint32_t m = 0; m < cmd->n_members; ++m) {
                uint64_t vi = cmd->member_voice[m];
                if (vi >= c->voices.size()) return fail(ORC_REF_PANIC, oob);
                Voice v = std::move(c->voices[vi]);
                c->voices.erase(c->voices.begin() + (ptrdiff_t)vi);
                uint32_t np = cmd->member_n_procs ? cmd->member_n_procs[m] : 0;
                if (cmd->member_update_tempo && cmd->member_update_tempo[m]) {
                    v.state.tempo = tempo;
                    for (uint32_t p = 0; p < np; ++p) {
                        uint64_t pi = cmd->member_proc_ids[pcur + p];
                        if (pi >= v.processes.size()) return fail(ORC_REF_PANIC, oob);
                        v.processes[pi].tempo = tempo;
                    }
                }
                pcur += np;
                g.voices.push_back(std::move(v));
            }
            c->groups.push_back(std::move(g));
            return ORC_OK;
        }
        case ORC_CMD_TC: {                                    // engine.rs:214-217
            TempoRef tempo = c->tempo_from_repr(cmd->tempo);
            if (!tempo) return fail(ORC_REF_PANIC, oob);
            c->tempo_cons.push_back(tempo);
            return ORC_OK;
        }
        case ORC_CMD_SEQ: {                                   // engine.rs:221-248
            TempoRef tempo = c->tempo_from_repr(cmd->tempo);
            if (!tempo) return fail(ORC_REF_PANIC, oob);
            Seq s;
            s.active = true; s.tempo = tempo; s.period = cmd->period;
            s.steps.assign(cmd->steps, cmd->steps + cmd->n_steps);
            s.chance.assign(cmd->chance, cmd->chance + cmd->n_steps);
            s.rng = orc_x128p{cmd->rng_s0, cmd->rng_s1};
            s.idx = 0;
            if (cmd->idx_kind == ORC_IDX_VOICE) {
                if (cmd->idx >= c->voices.size()) return fail(ORC_REF_PANIC, oob);
                Voice& v = c->voices[cmd->idx];
                v.processes.push_back(std::move(s));
                if (cmd->tempo.mode == ORC_TM_PROCESS) v.proc_tempi.push_back(tempo);
            } else if (cmd->idx_kind == ORC_IDX_GROUP) {
                if (cmd->idx >= c->groups.size()) return fail(ORC_REF_PANIC, oob);
                c->groups[cmd->idx].processes.push_back(std::move(s));
            }
            return ORC_OK;
        }
        case ORC_CMD_QUIT: return ORC_OK;                     // raise(SIGTERM) in the reference
        default: return fail(ORC_BAD_ARG, "unknown command kind");
    }
}

// engine.rs:46-81; bus is interleaved S16 (runtime.rs:272-276)
int orc_conductor_coordinate(orc_conductor* c, uint64_t frames, int16_t* bus) {
    const uint64_t oc = c->out_channels;
    for (uint64_t f = 0; f < frames; ++f) {
        for (uint64_t ch = 0; ch < oc; ++ch) {
            int16_t* sample_ptr = bus + f * oc + ch;
            *sample_ptr = 0;
            for (auto& v : c->voices)
                if (v.state.active)
                    if (!v.process(sample_ptr, ch)) return fail(ORC_REF_PANIC, "Seq index out of bounds");
            for (auto& g : c->groups)
                if (g.active)
                    if (!g.process(sample_ptr, ch)) return fail(ORC_REF_PANIC, "Seq index out of bounds");
        }
        c->clock += 1;
    }
    return ORC_OK;
}

int orc_conductor_n_voices(orc_conductor* c, int group) {
    if (group < 0) return (int)c->voices.size();
    if ((size_t)group >= c->groups.size()) return -1;
    return (int)c->groups[group].voices.size();
}
int orc_conductor_n_groups(orc_conductor* c) { return (int)c->groups.size(); }

static Voice* find_voice(orc_conductor* c, int group, uint32_t idx) {
    std::vector<Voice>* vs = nullptr;
    if (group < 0) vs = &c->voices;
    else if ((size_t)group < c->groups.size()) vs = &c->groups[group].voices;
    if (!vs || idx >= vs->size()) return nullptr;
    return &(*vs)[idx];
}
int orc_conductor_get_voice(orc_conductor* c, int group, uint32_t idx, orc_voice_state* out) {
    Voice* v = find_voice(c, group, idx);
    if (!v) return fail(ORC_BAD_ARG, "no such voice");
    out->active = v->state.active;
    out->position = v->state.position;
    out->velocity = v->state.velocity;
    out->gain = v->state.gain;
    out->end = v->state.end;
    out->channels = (uint32_t)v->channels;
    out->tempo_current = v->state.tempo->current;
    out->tempo_active = v->state.tempo->active;
    return ORC_OK;
}
int orc_conductor_set_voice(orc_conductor* c, int group, uint32_t idx, const float* position, const float* velocity, const float* gain, const int* active) {
    Voice* v = find_voice(c, group, idx);
    if (!v) return fail(ORC_BAD_ARG, "no such voice");
    if (position) v->state.position = *position;
    if (velocity) v->state.velocity = *velocity;
    if (gain) v->state.gain = *gain;
    if (active) v->state.active = *active != 0;
    return ORC_OK;
}
uint64_t orc_clock_current(orc_conductor* c) { return c->clock; }

// The position recurrence of Voice::process in isolation (engine.rs:407-410, 445-447):
// out[s] = position seen by advance-event s (s = 0..n inclusive; out[n] is the state afterwards).
// The advance is skipped, forever, from the first step whose trunc(position) >= end.
void orc_position_walk(float p0, float velocity, uint64_t end, uint64_t n, float* out) {
    float p = p0;
    for (uint64_t s = 0; s <= n; ++s) {
        out[s] = p;
        if (f32_as_usize(p) >= end) continue;
        p += velocity;
    }
}


// ===================== MPEG: mpeg.rs =====================
static const uint32_t BITRATES_COL4[15] = {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0};   // mpeg.rs:255-271

void orc_mpeg_parse_header(uint32_t h, orc_mpeg_header* o) {           // mpeg.rs:367-496
    std::memset(o, 0, sizeof(*o));
    uint8_t b1 = (uint8_t)(h >> 16), b2 = (uint8_t)(h >> 8), b3 = (uint8_t)h;
    uint8_t aaab = b1 >> 4;
    uint8_t version = (uint8_t)((aaab & 0x1) << 1);
    uint8_t bccd = b1 & 0x0F;
    version |= bccd & 0x1;                                             // mpeg.rs:377-383 (takes the protection bit)
    o->version_id = version;
    if (version == 1) { o->err = ORC_UNSUPPORTED_FORMAT; return; }
    uint8_t layer = (bccd >> 1) & 0x3;
    o->layer_id = layer;
    if (layer == 0) { o->err = ORC_UNSUPPORTED_FORMAT; return; }
    o->not_protected = bccd & 0x1;
    uint8_t eeee = b2 >> 4;
    if (eeee == 0 || eeee == 0xF) { o->err = ORC_UNSUPPORTED_FORMAT; return; }
    // match_bitrate (mpeg.rs:273-284): VL = (V << 2) & L is always 0 -> column 4
    uint8_t vl = (uint8_t)((version << 2) & layer);
    int col = vl == 0xF ? 0 : vl == 0xE ? 1 : vl == 0xD ? 2 : vl == 0xB ? 3 : 4;
    (void)col;
    o->bitrate = BITRATES_COL4[eeee - 1];
    uint8_t ffgh = b2 & 0x0F;
    double base = version == 3 ? 32000.0 : version == 2 ? 16000.0 : version == 0 ? 8000.0 : 0.0;   // mpeg.rs:286-303
    uint8_t ff = ffgh >> 2;
    double sr = ff == 0 ? base * 1.378125 : ff == 1 ? base * 1.5 : ff == 2 ? base : 0.0;
    if (sr == 0.0) { o->err = ORC_INVALID_DATA; return; }
    o->sr = sr;
    o->padded = (ffgh >> 1) & 0x1;
    o->channel_mode = (uint8_t)((b3 >> 4) >> 2);
    o->ok = 1;
    // Header::format (mpeg.rs:154-188)
    o->version = version == 0 ? 2.5f : version == 2 ? 2.0f : version == 3 ? 1.0f : 0.0f;
    o->layer = layer == 1 ? 3 : layer == 2 ? 2 : layer == 3 ? 1 : 0;
    bool prot = o->not_protected == 0;
    o->skip = prot ? 6 : 4;                                            // mpeg.rs:86-89
    // compute_frame_len (mpeg.rs:207-234)
    double br = (double)o->bitrate * 1000.0;
    double fl;
    switch (o->layer) {
        case 3: case 2: fl = 144.0 * br / sr; break;
        case 1: fl = (12.0 * br / sr) * 4.0; break;
        default: fl = 20.0; break;
    }
    if (fl < 20.0) { o->frame_len_ok = 0; return; }
    o->frame_len_ok = 1;
    o->payload_len = (uint64_t)fl - (prot ? 20 : 4) + (o->padded == 1 ? 1 : 0);
}

int orc_mpeg_match_ref(const orc_mpeg_header* a, const orc_mpeg_header* b) {   // mpeg.rs:194-204
    bool pa = a->not_protected == 0, pb = b->not_protected == 0;
    return a->version == b->version && a->layer == b->layer && a->sr == b->sr &&
           a->channel_mode == b->channel_mode && pa == pb;
}

int orc_mpeg_sync_scan(const uint8_t* r, uint64_t file_len, uint64_t* pos_out, uint32_t* hdr_out, uint64_t cap, uint64_t* n_out) {
    uint64_t cur = 0, n = 0;                                           // mpeg.rs:17-50
    while (cur < file_len) {
        uint8_t b = r[cur];
        if (b == 0xFF) {
            if (cur + 1 >= file_len) return fail(ORC_REF_PANIC, "index out of bounds: last byte is 0xFF (mpeg.rs:20)");
            if ((r[cur + 1] & 0xE0) == 0xE0) {
                uint64_t fp = cur;
                uint32_t supb = (uint32_t)r[cur] << 24;
                cur += 1; if (cur >= file_len) break;
                supb |= (uint32_t)r[cur] << 16;
                cur += 1; if (cur >= file_len) break;
                supb |= (uint32_t)r[cur] << 8;
                cur += 1; if (cur >= file_len) break;
                supb |= r[cur];
                if (n < cap) { if (pos_out) pos_out[n] = fp; if (hdr_out) hdr_out[n] = supb; }
                n += 1;
                cur += 1;
            } else {
                cur += 1;
            }
        } else {
            cur += 1;
        }
    }
    *n_out = n;
    return n > cap && (pos_out || hdr_out) ? fail(ORC_BAD_ARG, "candidate capacity too small") : ORC_OK;
}

int orc_mpeg_parse(const uint8_t* r, uint64_t file_len, int reference_compat,
                   uint64_t* offsets_out, uint64_t offsets_cap, uint64_t* n_offsets,
                   uint32_t* ref_header_out, uint64_t* n_candidates_out,
                   uint8_t* payload_out, uint64_t payload_cap, uint64_t* payload_len_out) {
    // possibles: HashMap<header, Vec<pos>> with the first position stored twice (mpeg.rs:39)
    std::unordered_map<uint32_t, std::vector<uint64_t>> possibles;
    uint64_t cur = 0, ncand = 0;
    while (cur < file_len) {
        if (r[cur] == 0xFF) {
            if (cur + 1 >= file_len) return fail(ORC_REF_PANIC, "index out of bounds: last byte is 0xFF (mpeg.rs:20)");
            if ((r[cur + 1] & 0xE0) == 0xE0) {
                uint64_t fp = cur;
                if (cur + 3 >= file_len) break;                        // the three `break`s at mpeg.rs:25-37
                uint32_t supb = ((uint32_t)r[cur] << 24) | ((uint32_t)r[cur + 1] << 16) | ((uint32_t)r[cur + 2] << 8) | r[cur + 3];
                auto it = possibles.find(supb);
                if (it == possibles.end()) {
                    auto& v = possibles[supb];
                    if (reference_compat) v.push_back(fp);             // or_insert(vec![fp]) ...
                    v.push_back(fp);                                   // ... .push(fp)
                } else {
                    it->second.push_back(fp);
                }
                ncand += 1;
                cur += 4;
                continue;
            }
        }
        cur += 1;
    }
    if (n_candidates_out) *n_candidates_out = ncand;
    // mpeg.rs:53-58: sort by count descending.  Ties follow HashMap order in the reference
    // (nondeterministic); the oracle breaks them by smallest header value (documented, deterministic).
    std::vector<std::pair<uint32_t, const std::vector<uint64_t>*>> vecs;
    vecs.reserve(possibles.size());
    for (auto& kv : possibles) vecs.push_back({kv.first, &kv.second});
    std::sort(vecs.begin(), vecs.end(), [](auto& a, auto& b) {
        if (a.second->size() != b.second->size()) return a.second->size() > b.second->size();
        return a.first < b.first;
    });
    // mpeg.rs:61-73: first header that parses is the reference; none -> index panic
    orc_mpeg_header ref;
    size_t i = 0;
    for (;; ++i) {
        if (i >= vecs.size()) return fail(ORC_REF_PANIC, "no parsable header (mpeg.rs:64 index out of bounds)");
        orc_mpeg_parse_header(vecs[i].first, &ref);
        if (ref.ok) break;
    }
    if (ref_header_out) *ref_header_out = vecs[i].first;
    // mpeg.rs:77-109
    struct Frame { uint64_t pos; uint64_t start, len; };
    std::vector<Frame> frames;
    for (auto& kv : vecs) {
        orc_mpeg_header h;
        orc_mpeg_parse_header(kv.first, &h);
        if (!h.ok) continue;
        if (!orc_mpeg_match_ref(&ref, &h)) continue;
        if (!h.frame_len_ok) continue;
        for (uint64_t index : *kv.second) {
            uint64_t start = index + h.skip, end = start + h.payload_len;
            if (end > file_len) return fail(ORC_REF_PANIC, "payload past EOF (mpeg.rs:96 index out of bounds)");
            frames.push_back({index, start, h.payload_len});
        }
    }
    std::stable_sort(frames.begin(), frames.end(), [](const Frame& a, const Frame& b) { return a.pos < b.pos; });   // mpeg.rs:112-116
    uint64_t plen = 0;
    for (size_t k = 0; k < frames.size(); ++k) {
        if (offsets_out && k < offsets_cap) offsets_out[k] = frames[k].pos;
        if (payload_out && plen + frames[k].len <= payload_cap) std::memcpy(payload_out + plen, r + frames[k].start, frames[k].len);
        plen += frames[k].len;
    }
    if (n_offsets) *n_offsets = frames.size();
    if (payload_len_out) *payload_len_out = plen;
    if (offsets_out && frames.size() > offsets_cap) return fail(ORC_BAD_ARG, "offset capacity too small");
    if (payload_out && plen > payload_cap) return fail(ORC_BAD_ARG, "payload capacity too small");
    return ORC_OK;
}

}  // extern "C"
