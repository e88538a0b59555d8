// render.cu — the voice render / mix-down path: K3 `voice_position_scan`, K4 `voice_render_mix`,
// K5 `bus_finalize`.
//
// Replaces the reference's per-sample scalar loop
//   blast/src/audio_processing/engine.rs:46-81    Conductor::coordinate  (frame -> channel -> voice)
//   blast/src/audio_processing/engine.rs:386-448  Voice::process         (gain, position stepping, lerp,
//                                                                          channel routing, i16 wrapping mix)
// with a time-parallel formulation that is bit-exact:
//
//  * The only loop-carried state of a voice is `position += velocity` in sequentially rounded f32
//    (engine.rs:446).  Inside one f32 binade the rounded increment is constant after the first
//    in-binade step (round-to-nearest-even settles the mantissa parity), so the whole trajectory
//    is a short list of arithmetic segments  pos(step) = p0 + (step - step0) * d * 2^e  with an
//    integer d — evaluated EXACTLY by one int multiply + one FFMA.  K3 builds that list per voice
//    with ordinary f32 adds (one thread per voice, O(#binades) work) and also emits one 16-byte
//    record per (tile, voice) so that K4 starts every tile with a single load.
//  * "steps" are advance events, not frames (engine.rs:419-427,445-447): a mono voice on >= 2
//    outputs advances twice per frame (L reads step 2f, R reads step 2f+1); a C-channel voice on
//    >= C outputs once per frame; on fewer outputs never.  A voice freezes at the first step whose
//    trunc(position) >= end (engine.rs:407-410 returns before the advance).
//  * The mix is integer: (sample * gain) as i16 (saturating, truncating, NaN -> 0) wrapping-added
//    into an i16 slot (engine.rs:441).  Wrapping i16 addition is addition mod 2^16, so K4 sums
//    int32 partials (any order, any number of GPUs) and K5 keeps the low 16 bits.
//
// K4 layout: grid = (frame tiles of 2048) x (voice groups); 256 threads; lanes own consecutive
// frames, so a warp reads 128 contiguous source bytes per stereo voice per load; every thread keeps
// 8 frames x out_channels int32 accumulators in registers over all voices of its group and issues one
// RED.ADD.S32 per bus slot at the end.  Roofline: HBM (2 B per voice-frame-channel of source, read once).
#include <algorithm>
#include <cstring>
#include <vector>

#include "blast_internal.h"

namespace {

constexpr int kMaxSeg = 160;
constexpr int kFT = 2048;                      // frames per tile
constexpr int kThreads = 256;
constexpr int kFPT = kFT / kThreads;           // frames per thread
constexpr int kVoiceBatch = 64;                // voices staged in shared memory at a time
constexpr int kMaxOut = 8;

struct VoiceDev {
    const int16_t* smp;
    uint32_t end;          // Voice::new: samples.len()/channels - 1 (engine.rs:302)
    uint32_t C;            // track channels
    float pos, vel, gain;
    uint32_t active;
    uint32_t S;            // advance events per frame for this voice on this bus: 0, 1 or 2
    uint32_t nch;          // bus channels this voice feeds
    uint32_t pad0, pad1;
};
static_assert(sizeof(VoiceDev) == 48, "VoiceDev layout");

struct Seg {               // positions for steps [step0, next.step0): p0 + (step-step0)*d*scale
    uint32_t step0;
    float p0;
    int32_t d;
    float scale;
};

struct TileRec {           // state of one voice at the first step of one tile
    float p0;
    int32_t d;
    float scale;
    uint32_t meta;         // [23:0] steps this segment still covers (saturating), [31:24] segment index
};

__device__ __forceinline__ uint32_t f2u_sat(float x) {       // Rust `as usize` on f32, clamped to u32
    uint32_t r;
    asm("cvt.rzi.u32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ int32_t f2i16_sat(float x) {      // Rust `as i16` on f32 (saturating, NaN -> 0)
    int32_t r;
    asm("{\n\t.reg .s16 t;\n\tcvt.rzi.s16.f32 t, %1;\n\tcvt.s32.s16 %0, t;\n\t}" : "=r"(r) : "f"(x));
    return r;
}

// sign / biased exponent / integer significand (in units of the binade's ulp); false for inf / NaN
__device__ __forceinline__ bool decompose(float x, uint32_t& sign, uint32_t& E, int32_t& q) {
    uint32_t b = __float_as_uint(x);
    sign = b >> 31;
    E = (b >> 23) & 0xFF;
    if (E == 255) return false;
    uint32_t M = b & 0x7FFFFF;
    q = (int32_t)(E ? (M | 0x800000u) : M);
    return true;
}
__device__ __forceinline__ float ulp_of_binade(uint32_t E) {  // 2^(max(E,1)-150)
    int e = (int)(E ? E : 1) - 150;
    return e >= -126 ? __uint_as_float((uint32_t)(e + 127) << 23) : __uint_as_float(1u << (e + 149));
}
__device__ __forceinline__ float seg_eval(float p0, int32_t d, float scale, uint32_t k) {
    // exact: |k*d| < 2^24 inside a run, scale is a power of two, the result is representable
    return __fmaf_rn((float)(int32_t)(k * (uint32_t)d), scale, p0);
}

// ---------------------------------------------------------------- K3
// One thread per voice.  Builds the segment list for `total = frames * S` advance events, the
// per-tile records, and writes the position after the render back into the voice.
__global__ void voice_position_scan(VoiceDev* __restrict__ voices, uint32_t n_voices, uint32_t frames,
                                    Seg* __restrict__ segs, uint32_t* __restrict__ nsegs,
                                    TileRec* __restrict__ recs, uint32_t n_tiles, uint32_t* __restrict__ err) {
    uint32_t vi = blockIdx.x * blockDim.x + threadIdx.x;
    if (vi >= n_voices) return;
    VoiceDev v = voices[vi];
    Seg* sg = segs + (size_t)vi * kMaxSeg;
    uint32_t n = 0;
    auto emit = [&](uint32_t step0, float p0, int32_t d, float scale) {
        if (n < (uint32_t)kMaxSeg) sg[n] = Seg{step0, p0, d, scale};
        n += 1;
    };
    float p = v.pos;
    const float vel = v.vel;
    const uint32_t total = v.active ? frames * v.S : 0;
    uint32_t s = 0;
    if (total == 0) emit(0, p, 0, 0.0f);
    while (s < total) {
        if (f2u_sat(p) >= v.end) { emit(s, p, 0, 0.0f); break; }                 // frozen (engine.rs:407-410)
        const float p1 = __fadd_rn(p, vel);
        if (__float_as_uint(p1) == __float_as_uint(p)) { emit(s, p, 0, 0.0f); break; }   // fixed point
        const float p2 = __fadd_rn(p1, vel);
        uint32_t s0, s1, s2, E0, E1, E2;
        int32_t q0, q1, q2;
        bool run = decompose(p, s0, E0, q0) && decompose(p1, s1, E1, q1) && decompose(p2, s2, E2, q2) &&
                   s0 == s1 && s1 == s2 && E0 == E1 && E1 == E2;
        if (run) {
            const int32_t d = q2 - q1;                      // magnitude increment in ulps, settled parity
            const bool from_p = (q1 - q0) == d;
            const int32_t qs = from_p ? q0 : q1;
            const uint32_t s_run = from_p ? s : s + 1;
            const float p_run = from_p ? p : p1;
            uint32_t kmax;                                  // further in-binade steps from qs
            if (d == 0) {
                kmax = 0xFFFFFFFFu;
            } else if (d > 0) {
                const int32_t qhi = (E0 ? 0xFFFFFF : 0x7FFFFF) - 1;
                kmax = qs <= qhi ? (uint32_t)((qhi - qs) / d) : 0u;
            } else {
                const int32_t qlo = E0 ? 0x800001 : 0;
                kmax = qs >= qlo ? (uint32_t)((qs - qlo) / (-d)) : 0u;
            }
            if (s0 == 0 && d > 0) {
                // growing positive position: stop the run at the first frozen step
                const int e = (int)(E0 ? E0 : 1) - 150;
                uint64_t thr;                               // smallest q with q * 2^e >= end
                if (e >= 0) thr = e >= 32 ? 1ull : (((uint64_t)v.end + (1ull << e) - 1) >> e);
                else thr = (-e) >= 40 ? ~0ull : ((uint64_t)v.end << (-e));
                if (thr > (uint64_t)qs) {
                    uint64_t kf = (thr - (uint64_t)qs + (uint64_t)d - 1) / (uint64_t)d;
                    if (kf < (uint64_t)kmax) kmax = (uint32_t)kf;
                } else {
                    kmax = 0;
                }
            }
            const uint32_t room = total - s_run;            // s_run <= total because s < total
            if (kmax > room) kmax = room;
            if (kmax >= 1) {
                if (!from_p) emit(s, p, 0, 0.0f);
                const int32_t ds = s0 ? -d : d;
                const float scale = ulp_of_binade(E0);
                emit(s_run, p_run, ds, scale);
                p = seg_eval(p_run, ds, scale, kmax);
                s = s_run + kmax;
                continue;
            }
        }
        emit(s, p, 0, 0.0f);
        p = p1;
        s += 1;
    }
    if (n > (uint32_t)kMaxSeg) {
        atomicExch(err, 1u);
        n = kMaxSeg;
    }
    nsegs[vi] = n;
    voices[vi].pos = p;

    // per-tile records, layout [tile][voice] so that K4's staging loads are coalesced
    uint32_t j = 0;
    for (uint32_t t = 0; t < n_tiles; ++t) {
        const uint32_t st = t * (uint32_t)kFT * v.S;
        while (j + 1 < n && sg[j + 1].step0 <= st) ++j;
        const Seg g = sg[j];
        const uint32_t k0 = st - g.step0;
        const uint32_t next = (j + 1 < n) ? sg[j + 1].step0 : 0xFFFFFFFFu;
        uint32_t left = next - st;
        if (left > 0xFFFFFFu) left = 0xFFFFFFu;
        TileRec r;
        r.p0 = seg_eval(g.p0, g.d, g.scale, k0);
        r.d = g.d;
        r.scale = g.scale;
        r.meta = left | (j << 24);
        recs[(size_t)t * n_voices + vi] = r;
    }
}

// ---------------------------------------------------------------- K4
struct VoiceS {            // what K4 needs per voice, staged in shared memory
    const int16_t* smp;
    uint32_t end, C;
    float vel, gain;
    uint32_t S, nch;
    uint32_t active, nseg;
    TileRec rec;
};

__device__ __forceinline__ float position_at(const VoiceS& v, const Seg* __restrict__ sg, bool fast,
                                             uint32_t tile_step0, uint32_t step_local) {
    if (fast) return seg_eval(v.rec.p0, v.rec.d, v.rec.scale, step_local);
    // the tile straddles a segment boundary: walk the (short) segment list from the tile's segment
    const uint32_t abs_step = tile_step0 + step_local;
    uint32_t j = v.rec.meta >> 24;
    while (j + 1 < v.nseg && sg[j + 1].step0 <= abs_step) ++j;
    const Seg g = sg[j];
    return seg_eval(g.p0, g.d, g.scale, abs_step - g.step0);
}

// (sample * gain) as i16 for one source channel at position p (engine.rs:429-442)
__device__ __forceinline__ int32_t voice_sample(const int16_t* __restrict__ sp, uint32_t C, float p, float vel,
                                                float gain) {
    const float s0 = (float)sp[0];
    float smp = s0;
    if (vel != 1.0f) {
        const float frac = __fsub_rn(p, truncf(p));                   // f32::fract
        const float s1 = (float)sp[C];
        smp = __fadd_rn(__fmul_rn(s0, __fsub_rn(1.0f, frac)), __fmul_rn(s1, frac));
    }
    return f2i16_sat(__fmul_rn(smp, gain));
}

template <int OC>
__global__ void __launch_bounds__(kThreads)
voice_render_mix(const VoiceDev* __restrict__ voices, uint32_t n_voices, uint32_t voices_per_group,
                 const Seg* __restrict__ segs, const uint32_t* __restrict__ nsegs,
                 const TileRec* __restrict__ recs, uint32_t frames, int32_t* __restrict__ bus, int use_atomic) {
    __shared__ VoiceS sv[kVoiceBatch];
    const uint32_t tile = blockIdx.x;
    const uint32_t f0 = tile * (uint32_t)kFT;
    const uint32_t nf = min((uint32_t)kFT, frames - f0);
    const uint32_t vbeg = blockIdx.y * voices_per_group;
    const uint32_t vend = min(n_voices, vbeg + voices_per_group);

    int32_t acc[kFPT][OC];
#pragma unroll
    for (int j = 0; j < kFPT; ++j)
#pragma unroll
        for (int c = 0; c < OC; ++c) acc[j][c] = 0;

    for (uint32_t vb = vbeg; vb < vend; vb += kVoiceBatch) {
        const uint32_t nb = min((uint32_t)kVoiceBatch, vend - vb);
        __syncthreads();
        if (threadIdx.x < nb) {
            const VoiceDev v = voices[vb + threadIdx.x];
            VoiceS s;
            s.smp = v.smp; s.end = v.end; s.C = v.C; s.vel = v.vel; s.gain = v.gain; s.S = v.S; s.nch = v.nch;
            s.active = v.active; s.nseg = nsegs[vb + threadIdx.x];
            s.rec = recs[(size_t)tile * n_voices + vb + threadIdx.x];
            sv[threadIdx.x] = s;
        }
        __syncthreads();
        for (uint32_t i = 0; i < nb; ++i) {
            const VoiceS& v = sv[i];
            if (!v.active) continue;
            const Seg* __restrict__ sg = segs + (size_t)(vb + i) * kMaxSeg;
            const uint32_t tile_step0 = f0 * v.S;
            const bool fast = (v.rec.meta & 0xFFFFFFu) >= (uint32_t)kFT * v.S + 2u || v.S == 0;
            if (v.C == 2 && v.nch == 2) {
                // stereo voice on a >= 2-channel bus: one 32-bit load fetches L and R of a frame
                const uint32_t* __restrict__ pairs = reinterpret_cast<const uint32_t*>(v.smp);
#pragma unroll
                for (int j = 0; j < kFPT; ++j) {
                    const uint32_t fl = threadIdx.x + j * kThreads;
                    if (fl < nf) {
                        const float p = position_at(v, sg, fast, tile_step0, fl);
                        const uint32_t idx = f2u_sat(p);
                        if (idx < v.end) {
                            const uint32_t w0 = __ldg(pairs + idx);
                            float l = (float)(int16_t)(w0 & 0xFFFF), r = (float)(int16_t)(w0 >> 16);
                            if (v.vel != 1.0f) {
                                const uint32_t w1 = __ldg(pairs + idx + 1);
                                const float frac = __fsub_rn(p, truncf(p));
                                const float om = __fsub_rn(1.0f, frac);
                                const float l1 = (float)(int16_t)(w1 & 0xFFFF), r1 = (float)(int16_t)(w1 >> 16);
                                l = __fadd_rn(__fmul_rn(l, om), __fmul_rn(l1, frac));
                                r = __fadd_rn(__fmul_rn(r, om), __fmul_rn(r1, frac));
                            }
                            acc[j][0] += f2i16_sat(__fmul_rn(l, v.gain));
                            if (OC > 1) acc[j][OC > 1 ? 1 : 0] += f2i16_sat(__fmul_rn(r, v.gain));
                        }
                    }
                }
            } else if (v.C == 1) {
                // mono voice: bus channels 0 and 1 read consecutive steps (engine.rs:419-422)
#pragma unroll
                for (int j = 0; j < kFPT; ++j) {
                    const uint32_t fl = threadIdx.x + j * kThreads;
                    if (fl < nf) {
#pragma unroll
                        for (int c = 0; c < (OC < 2 ? OC : 2); ++c) {
                            if ((uint32_t)c < v.nch) {
                                const float p = position_at(v, sg, fast, tile_step0, fl * v.S + c);
                                const uint32_t idx = f2u_sat(p);
                                if (idx < v.end) acc[j][c] += voice_sample(v.smp + idx, 1, p, v.vel, v.gain);
                            }
                        }
                    }
                }
            } else {
                // generic C-channel voice: one position per frame, channel ch reads source channel ch
#pragma unroll
                for (int j = 0; j < kFPT; ++j) {
                    const uint32_t fl = threadIdx.x + j * kThreads;
                    if (fl < nf) {
                        const float p = position_at(v, sg, fast, tile_step0, fl * v.S);
                        const uint32_t idx = f2u_sat(p);
                        if (idx < v.end) {
                            const int16_t* sp = v.smp + (size_t)idx * v.C;
#pragma unroll
                            for (int c = 0; c < OC; ++c)
                                if ((uint32_t)c < v.nch) acc[j][c] += voice_sample(sp + c, v.C, p, v.vel, v.gain);
                        }
                    }
                }
            }
        }
    }

#pragma unroll
    for (int j = 0; j < kFPT; ++j) {
        const uint32_t fl = threadIdx.x + j * kThreads;
        if (fl < nf) {
            int32_t* out = bus + (size_t)(f0 + fl) * OC;
#pragma unroll
            for (int c = 0; c < OC; ++c) {
                if (use_atomic) atomicAdd(out + c, acc[j][c]);
                else out[c] = acc[j][c];
            }
        }
    }
}

// ---------------------------------------------------------------- K5
// i16 wrapping accumulate == int32 sum mod 2^16 (engine.rs:441, release semantics)
__global__ void bus_finalize(const int32_t* __restrict__ partial, int16_t* __restrict__ bus, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n4 = n / 4;
    const bool aligned = (((uintptr_t)partial & 15) == 0) && (((uintptr_t)bus & 7) == 0);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (aligned) {
        for (uint64_t k = i; k < n4; k += stride) {
            const int4 v = reinterpret_cast<const int4*>(partial)[k];
            uint2 o;
            o.x = ((uint32_t)v.x & 0xFFFF) | ((uint32_t)v.y << 16);
            o.y = ((uint32_t)v.z & 0xFFFF) | ((uint32_t)v.w << 16);
            reinterpret_cast<uint2*>(bus)[k] = o;
        }
        for (uint64_t k = n4 * 4 + i; k < n; k += stride) bus[k] = (int16_t)partial[k];
    } else {
        for (uint64_t k = i; k < n; k += stride) bus[k] = (int16_t)partial[k];
    }
}

}  // namespace

struct blast_scene {
    uint32_t n_voices = 0;
    uint32_t out_channels = 0;
    std::vector<blast_track> tracks;
    std::vector<blast_voice> voices;     // host mirror of the ABI voices (positions refreshed on get)
    VoiceDev* d_voices = nullptr;
    Seg* d_segs = nullptr;
    uint32_t* d_nsegs = nullptr;
    uint32_t* d_err = nullptr;
    TileRec* d_recs = nullptr;
    size_t recs_cap = 0;                 // in records
};

namespace {

int make_voice_dev(const blast_scene* sc, const blast_voice& in, uint32_t index, VoiceDev* out) {
    if (in.track >= sc->tracks.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "voice %u: track index %u out of bounds (reference panics)", index, in.track);
    const blast_track& tr = sc->tracks[in.track];
    if (tr.num_channels == 0) return blast::set_error(BLAST_ERR_REF_PANIC, "voice %u: track has 0 channels (reference divides by zero)", index);
    const uint64_t frames = tr.n_samples / tr.num_channels;
    if (frames == 0) return blast::set_error(BLAST_ERR_REF_PANIC, "voice %u: empty track (usize underflow in Voice::new, engine.rs:302)", index);
    if (frames - 1 > 0x7FFFFFFFull) return blast::set_error(BLAST_ERR_CAPACITY, "voice %u: track longer than 2^31 frames", index);
    if (((uintptr_t)tr.d_samples & 3) != 0) return blast::set_error(BLAST_ERR_ARG, "voice %u: track samples must be 4-byte aligned", index);
    VoiceDev v{};
    v.smp = tr.d_samples;
    v.end = (uint32_t)(frames - 1);
    v.C = tr.num_channels;
    v.pos = in.position;
    v.vel = in.velocity;
    v.gain = in.gain;
    v.active = in.active ? 1u : 0u;
    const uint32_t oc = sc->out_channels;
    if (v.C == 1) {                 // engine.rs:419-422
        v.nch = oc < 2 ? oc : 2;
        v.S = v.nch;
    } else if (oc >= v.C) {         // engine.rs:425-427, 445-447
        v.nch = v.C;
        v.S = 1;
    } else {                        // ch == C-1 never happens: the voice never advances
        v.nch = oc;
        v.S = 0;
    }
    *out = v;
    return BLAST_OK;
}

int upload_voices(blast_ctx* ctx, blast_scene* sc) {
    std::vector<VoiceDev> hv(sc->n_voices);
    for (uint32_t i = 0; i < sc->n_voices; ++i)
        if (int rc = make_voice_dev(sc, sc->voices[i], i, &hv[i])) return rc;
    if (sc->n_voices) {
        BLAST_CUDA_TRY(cudaMemcpyAsync(sc->d_voices, hv.data(), hv.size() * sizeof(VoiceDev), cudaMemcpyHostToDevice, ctx->stream));
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return BLAST_OK;
}

}  // namespace

extern "C" {

int blast_scene_create(blast_ctx* ctx, const blast_track* tracks, uint32_t n_tracks, const blast_voice* voices,
                       uint32_t n_voices, uint32_t out_channels, blast_scene** out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_scene_create: out is null");
    BLAST_REQUIRE((tracks || n_tracks == 0) && (voices || n_voices == 0), BLAST_ERR_ARG, "blast_scene_create: null argument");
    if (out_channels < 1 || out_channels > (uint32_t)kMaxOut)
        return blast::set_error(BLAST_ERR_UNSUPPORTED, "out_channels must be 1..%d (got %u)", kMaxOut, out_channels);
    blast_scene* sc = new blast_scene();
    sc->n_voices = n_voices;
    sc->out_channels = out_channels;
    sc->tracks.assign(tracks, tracks + n_tracks);
    sc->voices.assign(voices, voices + n_voices);
    auto fail = [&](int rc) { blast_scene_destroy(ctx, sc); return rc; };
    const size_t nv = n_voices ? n_voices : 1;
    if (cudaMalloc(&sc->d_voices, nv * sizeof(VoiceDev)) != cudaSuccess ||
        cudaMalloc(&sc->d_segs, nv * kMaxSeg * sizeof(Seg)) != cudaSuccess ||
        cudaMalloc(&sc->d_nsegs, nv * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&sc->d_err, sizeof(uint32_t)) != cudaSuccess)
        return fail(blast::set_error(BLAST_ERR_CUDA, "blast_scene_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (int rc = upload_voices(ctx, sc)) return fail(rc);
    *out = sc;
    return BLAST_OK;
}

void blast_scene_destroy(blast_ctx* ctx, blast_scene* sc) {
    if (!sc) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (sc->d_voices) cudaFree(sc->d_voices);
    if (sc->d_segs) cudaFree(sc->d_segs);
    if (sc->d_nsegs) cudaFree(sc->d_nsegs);
    if (sc->d_err) cudaFree(sc->d_err);
    if (sc->d_recs) cudaFree(sc->d_recs);
    delete sc;
}

int blast_scene_set_voices(blast_ctx* ctx, blast_scene* sc, const blast_voice* voices, uint32_t n_voices) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc && voices, BLAST_ERR_ARG, "blast_scene_set_voices: null argument");
    BLAST_REQUIRE(n_voices == sc->n_voices, BLAST_ERR_ARG, "blast_scene_set_voices: voice count differs from the scene's");
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::vector<blast_voice> keep = sc->voices;
    sc->voices.assign(voices, voices + n_voices);
    int rc = upload_voices(ctx, sc);
    if (rc != BLAST_OK) sc->voices = keep;
    return rc;
}

int blast_scene_get_voices(blast_ctx* ctx, blast_scene* sc, blast_voice* out, uint32_t n_voices) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc && out, BLAST_ERR_ARG, "blast_scene_get_voices: null argument");
    BLAST_REQUIRE(n_voices == sc->n_voices, BLAST_ERR_ARG, "blast_scene_get_voices: voice count differs from the scene's");
    std::vector<VoiceDev> hv(sc->n_voices);
    if (sc->n_voices) {
        BLAST_CUDA_TRY(cudaMemcpyAsync(hv.data(), sc->d_voices, hv.size() * sizeof(VoiceDev), cudaMemcpyDeviceToHost, ctx->stream));
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    for (uint32_t i = 0; i < sc->n_voices; ++i) {
        sc->voices[i].position = hv[i].pos;
        out[i] = sc->voices[i];
    }
    return BLAST_OK;
}

int blast_scene_render_dev(blast_ctx* ctx, blast_scene* sc, uint64_t frames, int32_t* d_partial_bus) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc && (d_partial_bus || frames == 0), BLAST_ERR_ARG, "blast_scene_render_dev: null argument");
    if (frames == 0) return BLAST_OK;
    if (frames > 0x7FFFFFFFull) return blast::set_error(BLAST_ERR_CAPACITY, "at most 2^31-1 frames per render call");
    const uint32_t oc = sc->out_channels;
    const uint32_t n_tiles = (uint32_t)((frames + kFT - 1) / kFT);
    const size_t slots = (size_t)frames * oc;
    if (sc->n_voices == 0) {
        BLAST_CUDA_TRY(cudaMemsetAsync(d_partial_bus, 0, slots * sizeof(int32_t), ctx->stream));
        return BLAST_OK;
    }
    const size_t need = (size_t)n_tiles * sc->n_voices;
    if (need > sc->recs_cap) {
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (sc->d_recs) BLAST_CUDA_TRY(cudaFree(sc->d_recs));
        sc->d_recs = nullptr;
        sc->recs_cap = 0;
        BLAST_CUDA_TRY(cudaMalloc(&sc->d_recs, need * sizeof(TileRec)));
        sc->recs_cap = need;
    }
    BLAST_CUDA_TRY(cudaMemsetAsync(sc->d_err, 0, sizeof(uint32_t), ctx->stream));
    voice_position_scan<<<(sc->n_voices + 31) / 32, 32, 0, ctx->stream>>>(
        sc->d_voices, sc->n_voices, (uint32_t)frames, sc->d_segs, sc->d_nsegs, sc->d_recs, n_tiles, sc->d_err);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;

    // voice groups: enough CTAs to fill the GPU a few times over, groups of >= 64 voices
    uint32_t groups = 1;
    const uint32_t want_ctas = (uint32_t)ctx->sm_count * 16;
    while (n_tiles * groups < want_ctas && sc->n_voices / (groups * 2) >= 64) groups *= 2;
    const uint32_t per_group = (sc->n_voices + groups - 1) / groups;
    groups = (sc->n_voices + per_group - 1) / per_group;
    const int use_atomic = groups > 1;
    if (use_atomic) BLAST_CUDA_TRY(cudaMemsetAsync(d_partial_bus, 0, slots * sizeof(int32_t), ctx->stream));
    dim3 grid(n_tiles, groups);
#define BLAST_LAUNCH_MIX(OCV)                                                                              \
    voice_render_mix<OCV><<<grid, kThreads, 0, ctx->stream>>>(sc->d_voices, sc->n_voices, per_group, sc->d_segs, \
                                                               sc->d_nsegs, sc->d_recs, (uint32_t)frames,       \
                                                               d_partial_bus, use_atomic)
    switch (oc) {
        case 1: BLAST_LAUNCH_MIX(1); break;
        case 2: BLAST_LAUNCH_MIX(2); break;
        case 3: BLAST_LAUNCH_MIX(3); break;
        case 4: BLAST_LAUNCH_MIX(4); break;
        case 5: BLAST_LAUNCH_MIX(5); break;
        case 6: BLAST_LAUNCH_MIX(6); break;
        case 7: BLAST_LAUNCH_MIX(7); break;
        default: BLAST_LAUNCH_MIX(8); break;
    }
#undef BLAST_LAUNCH_MIX
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

int blast_scene_check(blast_ctx* ctx, blast_scene* sc) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc != nullptr, BLAST_ERR_ARG, "blast_scene_check: null scene");
    uint32_t e = 0;
    BLAST_CUDA_TRY(cudaMemcpyAsync(&e, sc->d_err, sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (e) return blast::set_error(BLAST_ERR_CAPACITY, "a voice trajectory needed more than %d position segments", kMaxSeg);
    return BLAST_OK;
}

int blast_bus_finalize_dev(blast_ctx* ctx, const int32_t* d_partial, int16_t* d_bus, uint64_t n_slots) {
    if (int rc = blast::bind(ctx)) return rc;
    if (n_slots == 0) return BLAST_OK;
    BLAST_REQUIRE(d_partial && d_bus, BLAST_ERR_ARG, "blast_bus_finalize_dev: null argument");
    uint64_t blocks = (n_slots / 4 + 255) / 256;
    int grid = (int)std::min<uint64_t>(std::max<uint64_t>(blocks, 1), (uint64_t)ctx->sm_count * 8);
    bus_finalize<<<grid, 256, 0, ctx->stream>>>(d_partial, d_bus, n_slots);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

int blast_render(blast_ctx* ctx, const blast_track* tracks, uint32_t n_tracks, const blast_voice* voices,
                 uint32_t n_voices, uint32_t out_channels, uint64_t frames, int16_t* host_bus_out,
                 blast_voice* voices_after) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(host_bus_out || frames == 0, BLAST_ERR_ARG, "blast_render: host_bus_out is null");
    blast_scene* sc = nullptr;
    int rc = blast_scene_create(ctx, tracks, n_tracks, voices, n_voices, out_channels, &sc);
    if (rc != BLAST_OK) return rc;
    const size_t slots = (size_t)frames * out_channels;
    int32_t* d_partial = nullptr;
    int16_t* d_bus = nullptr;
    auto done = [&](int code) {
        cudaStreamSynchronize(ctx->stream);
        if (d_partial) cudaFree(d_partial);
        if (d_bus) cudaFree(d_bus);
        blast_scene_destroy(ctx, sc);
        return code;
    };
    if (slots) {
        if (cudaMalloc(&d_partial, slots * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&d_bus, slots * sizeof(int16_t) + 16) != cudaSuccess)
            return done(blast::set_error(BLAST_ERR_CUDA, "blast_render: cudaMalloc failed"));
        if ((rc = blast_scene_render_dev(ctx, sc, frames, d_partial)) != BLAST_OK) return done(rc);
        if ((rc = blast_bus_finalize_dev(ctx, d_partial, d_bus, slots)) != BLAST_OK) return done(rc);
        if (cudaMemcpyAsync(host_bus_out, d_bus, slots * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            return done(blast::set_error(BLAST_ERR_CUDA, "blast_render: D2H failed"));
        if ((rc = blast_scene_check(ctx, sc)) != BLAST_OK) return done(rc);
    }
    if (voices_after && n_voices)
        if ((rc = blast_scene_get_voices(ctx, sc, voices_after, n_voices)) != BLAST_OK) return done(rc);
    return done(BLAST_OK);
}

}  // extern "C"
