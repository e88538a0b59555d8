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
//    with ordinary f32 adds (one warp per voice: lane 0 walks, O(#binades) work — one segment for an
//    integer velocity from an integer position) and the warp then emits one 16-byte record per
//    (tile, voice) so that K4 starts every tile with a single load.
//  * "steps" are advance events, not frames (engine.rs:419-427,445-447): a mono voice on >= 2
//    outputs advances twice per frame (L reads step 2f, R reads step 2f+1); a C-channel voice on
//    >= C outputs once per frame; on fewer outputs never.  A voice freezes at the first step whose
//    trunc(position) >= end (engine.rs:407-410 returns before the advance).
//  * The mix is integer: (sample * gain) as i16 (saturating, truncating, NaN -> 0) wrapping-added
//    into an i16 slot (engine.rs:441).  Wrapping i16 addition is addition mod 2^16, so K4 sums
//    int32 partials (any order, any number of GPUs) and K5 keeps the low 16 bits.
//
// K4 layout: work items = (frame tiles of 2048) x (voice groups), taken from a counter by persistent
// CTAs (3 per SM); 256 consumer threads + one producer warp; lanes own consecutive frames, so a warp reads
// 128 contiguous staged bytes per stereo voice per load; every thread keeps 8 frames x out_channels int32
// accumulators in registers over all voices of the item and issues one RED.ADD.S32 per bus slot at its
// end.  Roofline: HBM (2 B per voice-frame-channel of source, read once).  (> 2 bus channels: the plain
// one-CTA-per-item kernel `voice_render_mix`.)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "render_internal.h"

using namespace blast_rdr;

namespace {

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

// ---- advance map of voices that carry Seq processes.  Their step unit is the CALL (frame * oc + channel)
// because a retrigger can land between the channels of one frame; A(c) = advance events before call c.
__device__ __forceinline__ uint32_t adv_count(uint32_t c, uint32_t adv) {
    if (adv == 0) return c;
    const uint32_t oc = adv & 0xFF, lo = (adv >> 8) & 0xFF, na = (adv >> 16) & 0xFF;
    const uint32_t f = c / oc, r = c - f * oc;
    const uint32_t in = r > lo ? min(r - lo, na) : 0u;
    return f * na + in;
}
__device__ __forceinline__ uint32_t adv_first_call(uint32_t a, uint32_t adv) {   // smallest c with A(c) >= a
    if (adv == 0 || a == 0) return a;
    const uint32_t oc = adv & 0xFF, lo = (adv >> 8) & 0xFF, na = (adv >> 16) & 0xFF;
    if (na == 0) return 0xFFFFFFFFu;
    const uint32_t f = (a - 1) / na, r = (a - 1) - f * na;
    const unsigned long long c = (unsigned long long)f * oc + lo + r + 1;
    return c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c;
}
__device__ __forceinline__ float seg_pos(const Seg& g, uint32_t abs_step, uint32_t adv) {
    return seg_eval(g.p0, g.d, g.scale, adv_count(abs_step, adv) - adv_count(g.step0, adv));
}


// ---------------------------------------------------------------- bus sink: tile hand-off between ranks (peer memory)
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Step counters compared with wrap-around.  The spin is bounded: a rank that never publishes (it failed, or its process
// died) costs its peers timeout_ms, then bit 2 of *err is set and the caller carries on — blast_peer_bus_check reports it.
__device__ __noinline__ void wait_flag(const uint32_t* flag, uint32_t value, uint32_t timeout_ms, uint32_t* err) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int32_t)(v - value) >= 0) return;
    const uint64_t t0 = global_ns(), budget = (uint64_t)timeout_ms * 1000000ull;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - value) >= 0) return;
        if (timeout_ms && global_ns() - t0 > budget) {
            if (err) atomicOr(err, 4u);
            return;
        }
        __nanosleep(64);
    }
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// One thread, after a CTA-level barrier that follows the CTA's bus writes of one (tile, voice group) work item:
// the CTA that flushes the last group of a tile publishes the tile to the rank that will reduce it.
// The count is a RELEASE atomic (MEMBAR.ALL.GPU + ATOMG): unlike __threadfence() it does not invalidate the SM's L1
// (CCTL.IVALL), which the render's producer warps keep warm with voice rows and tile records.  Only the CTA that
// completes a tile pays the acquire side (the fence after the atomic makes the other CTAs' bus writes part of what its
// system-scope flag store releases).
// `n_parts` = counts a tile needs (its voice groups).  A single rank (world == 1) never leaves gpu scope.
__device__ __forceinline__ void sink_fence(const BusSink& s) {
    if (s.world > 1) __threadfence_system(); else __threadfence();
}
__device__ __forceinline__ void sink_tile_flushed(const BusSink& s, uint32_t tile, uint32_t n_parts) {
    uint32_t prev;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(s.tile_count + tile) : "memory");
    if (prev + 1u == n_parts) {
        sink_fence(s);
        st_release_sys(s.ready_at[tile % s.world] + tile, s.step);
    }
}
// The reduction of one tile by `nthreads` threads of one CTA (sync = their barrier): wait until every rank has published
// it, sum the int32 partial tiles of all ranks (peer memory: system-scope loads, never the non-coherent path), keep the
// low 16 bits (== i16 wrapping accumulate, engine.rs:441) and store them into the root's S16 bus.  The rank's last tile
// raises the done / ack flags.
template <typename Sync>
__device__ __forceinline__ void sink_reduce_tile(const BusSink& s, uint32_t tile, uint32_t tid, uint32_t nthreads, Sync sync) {
    if (tid < s.world) wait_flag(s.ready_mine + (size_t)tid * s.max_tiles + tile, s.step, s.timeout_ms, s.err);
    sync();
    const uint64_t base = (uint64_t)tile * s.tile_slots;
    const uint64_t left = s.n_slots - base;
    const uint32_t n = left < (uint64_t)s.tile_slots ? (uint32_t)left : s.tile_slots;
    const uint32_t n4 = n / 4;                                   // base is a multiple of 4 (tile_slots is)
    for (uint32_t k = tid; k < n4; k += nthreads) {
        int4 acc = make_int4(0, 0, 0, 0);
        for (uint32_t r0 = 0; r0 < s.world; r0 += 4) {           // four NVLink round trips in flight per thread
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = make_uint4(0, 0, 0, 0);
                if (r0 + j < s.world)
                    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                                 : "l"(reinterpret_cast<const uint4*>(s.part[r0 + j] + base) + k));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc.x += (int32_t)v[j].x; acc.y += (int32_t)v[j].y; acc.z += (int32_t)v[j].z; acc.w += (int32_t)v[j].w;
            }
        }
        uint2 o;
        o.x = ((uint32_t)acc.x & 0xFFFF) | ((uint32_t)acc.y << 16);
        o.y = ((uint32_t)acc.z & 0xFFFF) | ((uint32_t)acc.w << 16);
        reinterpret_cast<uint2*>(s.out + base)[k] = o;           // local on the root, an NVLink store elsewhere
    }
    for (uint32_t k = n4 * 4 + tid; k < n; k += nthreads) {
        int32_t acc = 0;
        for (uint32_t r = 0; r < s.world; ++r) {
            uint32_t x;
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(x) : "l"(s.part[r] + base + k));
            acc += (int32_t)x;
        }
        s.out[base + k] = (int16_t)acc;
    }
    sync();
    if (tid == 0) {
        sink_fence(s);
        const uint32_t prev = atomicAdd(s.red_count, 1u);
        if (prev + 1u == s.n_my_tiles) {
            *s.red_count = 0u;                                   // every other reduction of this step has finished
            sink_fence(s);
            for (uint32_t i = 0; i < s.n_done; ++i) st_release_sys(s.done[i], s.step);
        }
    }
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// out of line: the rare tile hand-off must not take registers from K4's consumer loop (72 at 3 CTAs per SM)
__device__ __noinline__ void k4_reduce_tile(const BusSink& s, uint32_t tile) {
    sink_reduce_tile(s, tile, threadIdx.x, 256u, [] { consumer_bar(); });
}
__device__ __noinline__ void k4_tile_flushed(const BusSink& s, uint32_t tile, uint32_t n_parts) {
    sink_tile_flushed(s, tile, n_parts);
}

// ---------------------------------------------------------------- K3a: Seq event scan (processes.rs:69-90)
// One thread per voice that carries Seq processes.  A Seq fires at call c when
//   fmodf((tempo.current as f32) / interval, period as f32) == steps[idx]        (exact f32 equality)
// with tempo.current = base + rate * c (u32, wrapping), then draws next_i64_range(0, 100) and, if the draw is
// below chance[idx], retriggers the voice.  x(c) = f32(current) / interval is monotone in c (rate >= 1, no
// wrap), and fmodf is exact, so a hit needs x(c) == steps[idx] + k * period for an integer k: for each k the
// first call reaching that value is found by bisection.  Other cases (rate 0, wrap, odd intervals) walk
// call by call.  Output: the sorted, de-duplicated retrigger calls of the voice.
__device__ __forceinline__ uint64_t xo_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
__device__ __forceinline__ uint64_t xo_next(unsigned long long& s0, unsigned long long& s1) {   // blast_rand.rs:31-39
    const uint64_t r = s0 + s1;
    const uint64_t t = s1 ^ s0;
    s0 = xo_rotl(s0, 55) ^ t ^ (t << 14);
    s1 = xo_rotl(t, 36);
    return r;
}
__device__ __forceinline__ long long f32_as_i64(float x) {      // Rust `as i64`
    if (x != x) return 0;
    if (x >= 9223372036854775808.0f) return 0x7FFFFFFFFFFFFFFFll;
    if (x <= -9223372036854775808.0f) return (long long)0x8000000000000000ull;
    return (long long)x;
}
__device__ __forceinline__ float seq_cur(uint32_t base, uint32_t rate, uint32_t c, float interval, float period_f) {
    const uint32_t cur = base + rate * c;                                        // u32 wrapping (blast_time.rs:113-115)
    return fmodf(__fdiv_rn((float)cur, interval), period_f);                      // blast_time.rs:118-121, processes.rs:77
}

// One WARP per voice: every lane carries the same Seq state (the walk over the hits is sequential: each hit draws from
// the Seq's generator and moves its step index), and the search for a hit — the first call whose tick quotient reaches
// the wanted value — is a 32-way search over the lanes instead of a bisection (5 dependent rounds of fdiv instead of 21
// for 2^21 calls: the kernel is nothing but that latency chain).  Lane 0 writes.
constexpr int kSeqScanThreads = 128;
__global__ void __launch_bounds__(kSeqScanThreads)
seq_event_scan(const VoiceDev* __restrict__ voices, uint32_t n_voices, SeqDev* __restrict__ seqs,
               uint32_t n_calls, uint32_t* __restrict__ events, uint32_t* __restrict__ nevents,
               uint32_t* __restrict__ err) {
    const uint32_t vi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (vi >= n_voices) return;                                 // warp-uniform
    const VoiceDev v = voices[vi];
    const uint32_t n_seq = v.first_seq >> 24, s_first = v.first_seq & 0xFFFFFFu;
    uint32_t* ev = events + (size_t)vi * kMaxEvents;
    uint32_t n_ev = 0;
    if (!v.active || n_seq == 0) { if (lane == 0) nevents[vi] = 0; return; }
    for (uint32_t si = 0; si < n_seq && n_ev <= (uint32_t)kMaxEvents; ++si) {
        SeqDev q = seqs[s_first + si];
        // n_steps == 0: the host parks Seqs whose tempo is inactive this way (processes.rs:74-75).
        // period 0: x % 0.0 is NaN, which equals nothing.
        if (q.n_steps == 0 || !(q.period_f > 0.0f)) continue;
        const float P = q.period_f;
        const unsigned long long last_tick = (unsigned long long)q.base + (unsigned long long)q.rate * (n_calls - 1);
        // x(c) = f32(current) / interval is monotone in c, and a period spans >= 16 calls (bounds the k walk)
        const bool monotone = q.rate >= 1 && q.interval > 0.0f && q.interval < 3.0e38f && last_tick <= 0xFFFFFFFFull &&
                              __fmul_rn(q.interval, P) >= __fmul_rn(16.0f, (float)q.rate);
        uint32_t c = 0;
        while (c < n_calls && n_ev <= (uint32_t)kMaxEvents) {
            const float t = q.steps[q.idx];
            uint32_t hit = 0xFFFFFFFFu;
            if (q.rate == 0) {
                if (seq_cur(q.base, 0, 0, q.interval, P) == t) hit = c;         // the tempo stands still
            } else if (monotone) {
                if (!(t >= 0.0f && t < P)) break;                               // fmodf never returns it: stuck forever
                const double Pd = (double)P;
                const float x_last = __fdiv_rn((float)(uint32_t)last_tick, q.interval);
                double k = floor((double)__fdiv_rn((float)(q.base + q.rate * c), q.interval) / Pd) - 1.0;
                if (k < 0.0) k = 0.0;
                for (;; k += 1.0) {
                    const double X = (double)t + k * Pd;                        // fmodf is exact: a hit needs x(c) == X
                    if (X > (double)x_last) break;
                    const float Xf = (float)X;
                    if ((double)Xf != X) continue;                              // not an f32: x(c) can never equal it
                    uint32_t lo = c, hi = n_calls;                              // first call with x >= Xf (x is monotone in the call)
                    while (hi - lo > 32u) {
                        const uint32_t step = (hi - lo) / 32u;                      // lane l probes lo + l * step  (< hi)
                        const uint32_t probe = lo + lane * step;
                        const uint32_t b = __ballot_sync(0xFFFFFFFFu, __fdiv_rn((float)(q.base + q.rate * probe), q.interval) >= Xf);
                        if (b == 0u) { lo = lo + 31u * step + 1u; }
                        else {
                            const uint32_t f = (uint32_t)__ffs((int)b) - 1u;       // first lane whose probe is at or past the hit
                            hi = lo + f * step;
                            if (f == 0u) break;                                     // lo itself: hi == lo ends the search
                            lo = lo + (f - 1u) * step + 1u;
                        }
                    }
                    if (hi > lo) {
                        const uint32_t probe = lo + lane;
                        const uint32_t b = __ballot_sync(0xFFFFFFFFu, probe < hi && __fdiv_rn((float)(q.base + q.rate * probe), q.interval) >= Xf);
                        lo = b ? lo + (uint32_t)__ffs((int)b) - 1u : hi;
                    }
                    if (lo < n_calls && seq_cur(q.base, q.rate, lo, q.interval, P) == t) { hit = lo; break; }
                }
            } else {
                for (uint32_t cc = c; cc < n_calls; ++cc)
                    if (seq_cur(q.base, q.rate, cc, q.interval, P) == t) { hit = cc; break; }
            }
            if (hit == 0xFFFFFFFFu) break;
            const uint64_t r = xo_next(q.s0, q.s1);                                // next_i64_range(0, 100): blast_rand.rs:50-59
            const long long draw = (long long)__umul64hi(r, 100ull);
            if (draw < f32_as_i64(q.chance[q.idx])) {
                if (n_ev < (uint32_t)kMaxEvents && lane == 0) ev[n_ev] = hit;
                n_ev += 1;
            }
            q.idx = (q.idx + 1) % q.n_steps;
            c = hit + 1;
        }
        if (lane == 0) {
            seqs[s_first + si].idx = q.idx;
            seqs[s_first + si].s0 = q.s0;
            seqs[s_first + si].s1 = q.s1;
        }
    }
    if (lane != 0) return;
    if (n_ev > (uint32_t)kMaxEvents) { atomicOr(err, 2u); n_ev = kMaxEvents; }
    // sort + de-duplicate (several Seqs of one voice may fire at the same call; the effect is the same reset)
    for (uint32_t i = 1; i < n_ev; ++i) {
        const uint32_t x = ev[i];
        uint32_t j = i;
        while (j > 0 && ev[j - 1] > x) { ev[j] = ev[j - 1]; --j; }
        ev[j] = x;
    }
    uint32_t m = 0;
    for (uint32_t i = 0; i < n_ev; ++i)
        if (m == 0 || ev[m - 1] != ev[i]) ev[m++] = ev[i];
    nevents[vi] = m;
}

// ---------------------------------------------------------------- K3
// One thread per voice.  Builds the segment list for `total = frames * S` advance events, the
// per-tile records, and writes the position after the render back into the voice.
// Positions of `n_adv` advance events starting from position p: emits arithmetic segments through
// emit(advance index relative to the start, p0, d, scale) and returns the position after the last advance
// (the walk stops at a frozen position or a fixed point, which then holds forever).
template <typename Emit>
__device__ __forceinline__ float build_epoch(float p, const float vel, const uint32_t end, const uint32_t total, Emit emit) {
    uint32_t s = 0;
    if (total == 0) emit(0u, p, 0, 0.0f);
    while (s < total) {
        if (f2u_sat(p) >= end) { emit(s, p, 0, 0.0f); break; }                   // frozen (engine.rs:407-410)
        if (vel >= 1.0f && vel < 16777216.0f && p >= 0.0f && p < 16777216.0f) {
            // integer position, integer velocity (the default velocity 1.0 from position 0): every `position += velocity`
            // below 2^24 is exact whatever binades it crosses, so the run up to the first frozen step (or 2^24, or the
            // end of the epoch) is ONE segment in units of 1.0 instead of three or four pieces per binade
            const uint32_t pi = (uint32_t)p, vv = (uint32_t)vel;
            if ((float)pi == p && (float)vv == vel) {
                uint32_t kmax = (16777215u - pi) / vv;
                const uint64_t kf = ((uint64_t)(end - pi) + vv - 1u) / vv;         // first frozen step (end > pi here)
                if (kf < (uint64_t)kmax) kmax = (uint32_t)kf;
                if (kmax > total - s) kmax = total - s;
                if (kmax >= 1) {
                    emit(s, p, (int32_t)vv, 1.0f);
                    p = (float)(pi + kmax * vv);
                    s += kmax;
                    continue;
                }
            }
        }
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
                if (e >= 0) thr = e >= 32 ? 1ull : (((uint64_t)end + (1ull << e) - 1) >> e);
                else thr = (-e) >= 40 ? ~0ull : ((uint64_t)end << (-e));
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
    return p;
}

// The serial part of K3, run by lane 0 of the voice's warp.  Builds the segment list for `frames * S` steps (advance
// events, or calls for voices with Seq processes, whose retrigger events start new epochs: processes.rs:82-85) and
// writes the position after the render back into the voice.  Returns the number of segments.
__device__ __noinline__ uint32_t scan_voice(VoiceDev* __restrict__ voices, const uint32_t vi, uint32_t frames,
                                            Seg* __restrict__ segs, uint32_t* __restrict__ nsegs, uint32_t* __restrict__ err,
                                            const uint32_t* __restrict__ events, const uint32_t* __restrict__ nevents,
                                            const uint32_t seg_cap, const uint32_t oc, Split* __restrict__ splits,
                                            uint32_t* __restrict__ nsplits, uint3* cmds) {
    // cmds (shared, one list per warp): cmds[0] = {number of commands, cap_eff, 0}; cmds[1 + i] = {first segment index,
    // segment count, step offset} — epochs that are copies of the voice's template, left to the whole warp
    cmds[0] = make_uint3(0u, 0u, 0u);
    VoiceDev v = voices[vi];
    Seg* sg = segs + (size_t)vi * seg_cap;
    uint32_t n = 0;
    float p = v.pos;
    const uint32_t total = v.active ? frames * v.S : 0;
    const uint32_t n_seq = v.first_seq >> 24;
    if (nsplits) nsplits[vi] = 0;
    if (n_seq == 0 || total == 0) {
        p = build_epoch(p, v.vel, v.end, total, [&](uint32_t a, float p0, int32_t d, float scale) {
            if (n < seg_cap) sg[n] = Seg{a, p0, d, scale};
            n += 1;
        });
    } else if (v.adv == 0) {
        // Voice with Seq processes that advances (S >= 1): steps stay advance events, exactly as for a plain voice.
        // A retrigger before the read of call c = f * oc + k (processes.rs:82-85) starts a new epoch at home position
        // at step A(c) = advance events before call c.  For a C >= 2 voice hit at 0 < k < C the channels < k of frame
        // f have already read the OLD position of step f: that is recorded as a split and patched by K4b.
        const uint32_t* ev = events + (size_t)vi * kMaxEvents;
        const uint32_t n_ev = nevents[vi];
        const float home = v.vel >= 0.0f ? 0.0f : (float)v.end;
        Split* sp = splits + (size_t)vi * kMaxEvents;
        uint32_t n_sp = 0, cur = 0, last0 = 0xFFFFFFFFu;
        // Every epoch after a retrigger starts at `home` with the same velocity, so its segment sequence is the same
        // every time up to where it is cut off: it is built once, as a template at the tail of the voice's segment
        // area, and instantiated per epoch by shifting step0 (the closed form per segment is exact, so the position
        // at the cut is too).  Only the first epoch (carried-in position) is built directly.
        constexpr uint32_t kTpl = 64;
        const bool use_tpl = n_ev > 0 && seg_cap > 4 * kTpl;
        const uint32_t cap_eff = use_tpl ? seg_cap - kTpl : seg_cap;            // segments the voice itself may use
        Seg* tpl = sg + cap_eff;
        uint32_t n_tpl = 0;
        bool tpl_ok = false;
        if (use_tpl) {
            // total + 1 steps: the position AFTER the last advance of an epoch (step n_adv <= total) must be covered too
            build_epoch(home, v.vel, v.end, total + 1u, [&](uint32_t rel, float p0, int32_t d, float scale) {
                if (n_tpl < kTpl) tpl[n_tpl] = Seg{rel, p0, d, scale};
                n_tpl += 1;
            });
            tpl_ok = n_tpl <= kTpl;
        }
        auto put_capped = [&](uint32_t step0, float p0, int32_t d, float scale) {
            if (step0 == last0 && n > 0) {                                        // an epoch of zero steps: replaced by its successor
                if (n <= cap_eff) sg[n - 1] = Seg{step0, p0, d, scale};
                return;
            }
            last0 = step0;
            if (n < cap_eff) sg[n] = Seg{step0, p0, d, scale};
            n += 1;
        };
        bool at_home = false;                                                    // the current epoch starts at `home`
        for (uint32_t e = 0; e <= n_ev; ++e) {
            uint32_t a = total, f = 0, k = 0;
            if (e < n_ev) {
                const uint32_t c = ev[e];
                f = c / oc;
                k = c - f * oc;
                a = v.C == 1 ? f * v.S + min(k, v.nch) : f + (k >= v.C ? 1u : 0u);
                if (a > total) a = total;
            }
            const uint32_t base_step = cur, n_adv = a - cur;
            if (at_home && tpl_ok) {
                // the epoch uses template segments 0 .. t-1 (segment 0 always, then those that start before its last
                // step; step0 ascends: bisection).  Equivalent to put_capped() on each of them, but only the bookkeeping
                // is done here: the copy itself is the warp's (voice_position_scan), ~45 segments x ~87 epochs per voice
                // on C3 + Seq were this thread's serial load-store chain.
                uint32_t t = 1, t_hi = n_tpl;
                while (t < t_hi) {
                    const uint32_t mid = (t + t_hi) >> 1;
                    if (tpl[mid].step0 < n_adv) t = mid + 1; else t_hi = mid;
                }
                const uint32_t start = (base_step == last0 && n > 0) ? n - 1 : n;   // an epoch of zero steps is replaced
                const uint32_t n_cmd = cmds[0].x;
                cmds[1 + n_cmd] = make_uint3(start, t, base_step);                   // at most n_ev + 1 <= kMaxEvents + 1 epochs
                cmds[0] = make_uint3(n_cmd + 1, cap_eff, 0u);
                n = start + t;
                last0 = base_step + tpl[t - 1].step0;
                // position after n_adv advances = position of step n_adv: in the last template segment that starts at or
                // before it (a segment's closed form is only valid up to its own last step)
                const Seg g = (t < n_tpl && tpl[t].step0 <= n_adv) ? tpl[t] : tpl[t - 1];
                p = seg_eval(g.p0, g.d, g.scale, n_adv - g.step0);
            } else {
                p = build_epoch(p, v.vel, v.end, n_adv, [&](uint32_t rel, float p0, int32_t d, float scale) {
                    put_capped(base_step + rel, p0, d, scale);
                });
            }
            if (e == n_ev) break;
            if (v.C >= 2 && k > 0 && k < v.C && k < oc) {                         // channels < k of frame f read the old position
                if (n_sp < (uint32_t)kMaxEvents) sp[n_sp] = Split{f, k, p, 0u};
                n_sp += 1;
            }
            p = home;
            cur = a;
            at_home = true;
        }
        if (n > cap_eff) n = seg_cap + 1;                                        // reported as an overflow below
        nsplits[vi] = min(n_sp, (uint32_t)kMaxEvents);
    } else {
        // steps are calls; epochs are delimited by the retrigger events found by seq_event_scan
        const uint32_t* ev = events + (size_t)vi * kMaxEvents;
        const uint32_t n_ev = nevents[vi];
        uint32_t cc = 0, e = 0, last0 = 0xFFFFFFFFu;
        auto put = [&](uint32_t step0, float p0, int32_t d, float scale) {
            if (step0 == last0 && n > 0) {                                        // same call as the previous segment: replace it
                if (n <= seg_cap) sg[n - 1] = Seg{step0, p0, d, scale};
                return;
            }
            last0 = step0;
            if (n < seg_cap) sg[n] = Seg{step0, p0, d, scale};
            n += 1;
        };
        for (;;) {
            if (e < n_ev && ev[e] == cc) {                                        // retrigger before the read at call cc
                p = v.vel >= 0.0f ? 0.0f : (float)v.end;                           // processes.rs:82-85
                e += 1;
            }
            const uint32_t stop = (e < n_ev && ev[e] < total) ? ev[e] : total;   // next retrigger (or the end); stop > cc
            const uint32_t a_cc = adv_count(cc, v.adv);
            const uint32_t n_adv = adv_count(stop, v.adv) - a_cc;                // advancing calls in [cc, stop)
            p = build_epoch(p, v.vel, v.end, n_adv, [&](uint32_t a, float p0, int32_t d, float scale) {
                put(a == 0 ? cc : adv_first_call(a_cc + a, v.adv), p0, d, scale);
            });
            // calls between the last advancing call and `stop` read the position after all n_adv advances
            if (n_adv > 0 && adv_count(stop - 1, v.adv) - a_cc == n_adv) put(adv_first_call(a_cc + n_adv, v.adv), p, 0, 0.0f);
            cc = stop;
            if (stop >= total) break;
        }
    }
    if (n > seg_cap) {
        atomicOr(err, 1u);
        n = seg_cap;
    }
    nsegs[vi] = n;
    voices[vi].pos = p;
    return n;
}

// The piece table of one tile of a stereo voice on a stereo bus whose trajectory has several segments inside the tile
// (binade crossings after a retrigger, a freeze, ...): every segment's part of the tile is a PIECE — frame range, integer
// significand at tile frame 0, increment, fraction bits — and consecutive pieces whose source spans fit one stage together
// form an ITEM that K4 stages with one bulk copy.  Layout in the voice's pool: per item one header row
// {first source frame, last source frame, piece count, 0} followed by its piece rows {fa | fe << 16, q0, d, sh | silent << 8}
// (what consume_stereo_multi reads).  Returns the rows used, 0 when the tile cannot be tabulated (a piece with a NaN or
// negative position, more than kMaxPieces segments, a piece that does not fit a stage): K4 then cuts it itself.
// rows == nullptr: count only.  *n_items_out = items written.
constexpr int kMaxPieces = 32;                // pieces of one voice in one tile handled by a single staged item
__device__ __forceinline__ bool tab_span_fits(const int16_t* smp, uint32_t lo, uint32_t hi, uint32_t stage_bytes) {
    const unsigned long long b0 = (unsigned long long)smp + (unsigned long long)lo * 4ull;
    const unsigned long long b1 = (unsigned long long)smp + ((unsigned long long)hi + 2ull) * 4ull;
    return ((b1 + 15ull) & ~15ull) - (b0 & ~15ull) <= (unsigned long long)stage_bytes;
}
__device__ __noinline__ uint32_t build_tile_table(const Seg* __restrict__ sg, uint32_t nseg, uint32_t j0, uint32_t f0, uint32_t nf,
                                                  uint32_t end, const int16_t* smp, uint32_t stage_bytes, uint4* rows,
                                                  uint32_t* n_items_out) {
    const uint32_t last_abs = f0 + nf - 1;
    uint32_t used = 0, n_items = 0, n_in_item = 0, hdr_at = 0;
    uint32_t lo_run = 0xFFFFFFFFu, hi_run = 0u;
    auto close_item = [&]() {
        if (n_in_item == 0) return;
        if (lo_run != 0xFFFFFFFFu) {                               // something audible: keep the item
            if (rows) rows[hdr_at] = make_uint4(lo_run, hi_run, n_in_item, 0u);
            n_items += 1;
        } else {
            used = hdr_at;                                         // a run of silent pieces: dropped (frames outside every piece read as zero)
        }
        n_in_item = 0;
        lo_run = 0xFFFFFFFFu;
        hi_run = 0u;
    };
    uint32_t n_pieces = 0;
    for (uint32_t j = j0; j < nseg; ++j) {
        const Seg g = sg[j];
        if (j > j0 && g.step0 > last_abs) break;
        if (++n_pieces > (uint32_t)kMaxPieces) return 0u;
        const uint32_t nxt = (j + 1 < nseg) ? sg[j + 1].step0 : 0xFFFFFFFFu;
        const uint32_t ls = max(g.step0, f0), le = min(nxt, f0 + nf);
        const uint32_t fa = ls - f0, fe = le - f0;
        const float p_a = seg_eval(g.p0, g.d, g.scale, ls - g.step0);
        const float p_l = seg_eval(g.p0, g.d, g.scale, le - 1 - g.step0);
        const bool weird = (p_a != p_a) || (p_l != p_l);
        const uint32_t lo = f2u_sat(fminf(p_a, p_l)), hi = f2u_sat(fmaxf(p_a, p_l));
        const bool silent = !weird && lo >= end;
        uint32_t shv = 0;
        int32_t q0v = 0;
        if (!silent) {
            if (weird || !(p_a >= 0.0f) || !(p_l >= 0.0f) || hi >= end) return 0u;
            if (g.d != 0) {
                const uint32_t eb = (__float_as_uint(g.scale) >> 23) & 0xFF;
                if (eb < 96 || eb > 127) return 0u;
                shv = 127 - eb;
            } else {
                const int E = (int)((__float_as_uint(p_a) >> 23) & 0xFF);
                const int s_ = p_a == 0.0f ? 0 : max(0, 150 - E);
                if (s_ > 31) return 0u;
                shv = (uint32_t)s_;
            }
            q0v = __float2int_rz(__fmul_rn(p_a, __uint_as_float((127u + shv) << 23))) - (int32_t)fa * g.d;
            if (!tab_span_fits(smp, lo, hi, stage_bytes)) return 0u;
            // does the piece still fit the stage of the current item?
            const uint32_t nlo = min(lo_run, lo), nhi = max(hi_run, hi);
            if (n_in_item > 0 && lo_run != 0xFFFFFFFFu && !tab_span_fits(smp, nlo, nhi, stage_bytes)) close_item();
        }
        if (n_in_item == 0) { hdr_at = used; used += 1; }
        if (!silent) { lo_run = min(lo_run, lo); hi_run = max(hi_run, hi); }
        if (rows) rows[used] = make_uint4(fa | (fe << 16), (uint32_t)q0v, (uint32_t)g.d, shv | (silent ? 0x100u : 0u));
        used += 1;
        n_in_item += 1;
    }
    close_item();
    *n_items_out = n_items;
    return n_items ? used : 0u;
}

// K3: one WARP per voice.  Lane 0 walks the trajectory (scan_voice); then the 32 lanes write the voice's per-tile
// records — the state of the voice at the first step of every tile, found by bisection in the segment list the warp has
// just written (still in L1).  Layout [tile][voice] so that K4's staging loads are coalesced.  (As a serial loop in the
// walking thread the records were 131 us for C2's 1,024 voices x 352 tiles; as a kernel of their own, one thread per
// (tile, voice), 24 us plus a launch boundary.)  A voice whose list overflowed leaves records K4 never reads: K4 exits
// on *err.
constexpr int kScanThreads = 128;
constexpr size_t kZeroPerBlock = 8192;           // int32 slots cleared per housekeeping block
__global__ void __launch_bounds__(kScanThreads)
voice_position_scan(VoiceDev* __restrict__ voices, uint32_t n_voices, uint32_t frames,
                    Seg* __restrict__ segs, uint32_t* __restrict__ nsegs, uint32_t* __restrict__ err,
                    const uint32_t* __restrict__ events, const uint32_t* __restrict__ nevents,
                    const uint32_t seg_cap, const uint32_t oc, Split* __restrict__ splits,
                    uint32_t* __restrict__ nsplits, TileRec* __restrict__ recs, uint32_t n_tiles,
                    uint32_t n_voice_blocks, uint32_t* __restrict__ err_next, uint32_t* __restrict__ work,
                    uint32_t* __restrict__ zero, size_t n_zero, const BusSink sink, uint4* __restrict__ pool,
                    const uint32_t pool_rows, const uint32_t stage_bytes, const VoiceDev* __restrict__ rewind) {
    if (blockIdx.x >= n_voice_blocks) {
        // housekeeping blocks, concurrent with the walks: the next render's error word and K4's work counter, and the
        // int32 partial bus when K4 will accumulate with atomics (as stream memsets these were two more operations —
        // and engine switches — between the kernels of every render)
        const uint32_t b = blockIdx.x - n_voice_blocks;
        if (sink.world > 1) {
            // the partial bus is read by the peers that reduce its tiles: nobody clears it (and K4, which follows this
            // kernel, does not write it) before every rank has finished with the previous step
            if (threadIdx.x < sink.world) wait_flag(sink.ack_mine + threadIdx.x, sink.step - 1u, sink.timeout_ms, sink.err);
            __syncthreads();
        }
        if (b == 0) {
            if (threadIdx.x == 0) {
                *err_next = 0u;
                *work = 0u;
            }
            if (sink.world)
                for (uint32_t i = threadIdx.x; i < sink.n_tiles; i += kScanThreads) sink.tile_count[i] = 0u;
        }
        const size_t lo = (size_t)b * kZeroPerBlock, hi = lo + kZeroPerBlock < n_zero ? lo + kZeroPerBlock : n_zero;
        for (size_t i = lo + threadIdx.x; i < hi; i += kScanThreads) zero[i] = 0u;
        return;
    }
    const uint32_t vi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (vi >= n_voices) return;                                 // warp-uniform
    if (rewind) {                                               // blast_scene_restore_dev: start from the uploaded voice, no copy kernel
        if (lane == 0) voices[vi] = rewind[vi];
        __syncwarp();
    }
    const uint32_t active = voices[vi].active, S = voices[vi].S, adv = voices[vi].adv;      // scan_voice only writes .pos
    __shared__ uint3 s_cmds[kScanThreads / 32][kMaxEvents + 2];
    uint3* cmds = s_cmds[threadIdx.x >> 5];
    uint32_t n = 0;
    if (lane == 0) n = scan_voice(voices, vi, frames, segs, nsegs, err, events, nevents, seg_cap, oc, splits, nsplits, cmds);
    n = __shfl_sync(0xFFFFFFFFu, n, 0);
    __syncwarp();                                               // orders lane 0's segment stores before the reads below
    Seg* sgw = segs + (size_t)vi * seg_cap;
    {
        // template epochs of a voice with Seq processes: copy segments 0 .. count-1 of the template (stored behind the
        // voice's own cap_eff segments) to their place, step0 shifted; in order, a later epoch may replace the last
        // segment of an earlier one
        const uint32_t n_cmd = cmds[0].x, cap_eff = cmds[0].y;
        const Seg* tpl = sgw + cap_eff;
        for (uint32_t c = 0; c < n_cmd; ++c) {
            const uint3 cm = cmds[1 + c];
            for (uint32_t tt = lane; tt < cm.y; tt += 32) {
                const uint32_t idx = cm.x + tt;
                if (idx < cap_eff) {
                    Seg g = tpl[tt];
                    g.step0 += cm.z;
                    sgw[idx] = g;
                }
            }
            __syncwarp();
        }
    }
    if (!active) return;                                        // K4 never reads the records of an inactive voice
    const Seg* sg = sgw;
    // piece tables (stereo voice on a stereo bus, steps = frames): built here, where a warp per voice has the segments
    // at hand, so that K4's single producer warp only bulk-copies them
    const VoiceDev vd = voices[vi];
    const bool tabulate = pool != nullptr && oc == 2 && vd.C == 2 && vd.nch == 2 && adv == 0 && S == 1;
    uint4* vpool = pool ? pool + (size_t)vi * pool_rows : nullptr;
    uint32_t pool_used = 0;                                     // warp-uniform
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += 32) {
        const uint32_t t = t0 + lane;
        const bool in = t < n_tiles;
        TileRec r{};
        uint32_t j = 0, need = 0, n_it = 0, f0 = 0, nf = 0;
        if (in) {
            const uint32_t st = t * (uint32_t)kFT * S;
            uint32_t lo = 0, hi = n;                            // last j with sg[j].step0 <= st (sg[0].step0 == 0)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (sg[mid].step0 <= st) lo = mid; else hi = mid;
            }
            j = lo;
            const Seg g = sg[j];
            const uint32_t next = (j + 1 < n) ? sg[j + 1].step0 : 0xFFFFFFFFu;
            uint32_t left = next - st;
            if (left > 0xFFFFu) left = 0xFFFFu;              // only compared with kFT * S + 2 <= 16,386
            r.p0 = seg_pos(g, st, adv);
            r.d = g.d;
            r.scale = g.scale;
            r.meta = left | (j << 16);
            f0 = t * (uint32_t)kFT;
            nf = min((uint32_t)kFT, frames - f0);
            if (tabulate && left < nf) need = build_tile_table(sg, n, j, f0, nf, vd.end, vd.smp, stage_bytes, nullptr, &n_it);
        }
        if (tabulate && __any_sync(0xFFFFFFFFu, need != 0u)) {
            uint32_t inc = need;                                // exclusive scan over the lanes: where each tile's rows go
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, inc, dlt);
                if ((int)lane >= dlt) inc += a;
            }
            const uint32_t at = pool_used + inc - need;
            if (need != 0u && at + need <= pool_rows) {
                build_tile_table(sg, n, j, f0, nf, vd.end, vd.smp, stage_bytes, vpool + at, &n_it);
                r.d = (int32_t)at;                              // (neither field is read for a tile with several segments)
                r.scale = __uint_as_float(kTabTag | n_it);
            }
            pool_used += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (in) recs[(size_t)t * n_voices + vi] = r;
    }
}

// ---------------------------------------------------------------- K4
struct VoiceS {            // what K4 needs per voice, staged in shared memory
    const int16_t* smp;
    uint32_t end, C;
    float vel, gain;
    uint32_t S, nch;
    uint32_t active, nseg;
    uint32_t adv;
    TileRec rec;
};

__device__ __forceinline__ float position_eval(bool fast, float p0, int32_t d, float scale, uint32_t recmeta,
                                               const Seg* __restrict__ sg, uint32_t nseg, uint32_t tile_step0,
                                               uint32_t step_local, uint32_t adv = 0) {
    if (fast) return seg_eval(p0, d, scale, adv_count(tile_step0 + step_local, adv) - adv_count(tile_step0, adv));
    // the tile straddles a segment boundary: walk the (short) segment list from the tile's segment
    const uint32_t abs_step = tile_step0 + step_local;
    uint32_t j = recmeta >> 16;
    while (j + 1 < nseg && sg[j + 1].step0 <= abs_step) ++j;
    return seg_pos(sg[j], abs_step, adv);
}

__device__ __forceinline__ float position_at(const VoiceS& v, const Seg* __restrict__ sg, bool fast,
                                             uint32_t tile_step0, uint32_t step_local) {
    return position_eval(fast, v.rec.p0, v.rec.d, v.rec.scale, v.rec.meta, sg, v.nseg, tile_step0, step_local, v.adv);
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
                 const TileRec* __restrict__ recs, uint32_t frames, int32_t* __restrict__ bus, int use_atomic,
                 const uint32_t* __restrict__ err, const uint32_t seg_cap) {
    __shared__ VoiceS sv[kVoiceBatch];
    if (*err) return;                                           // truncated trajectories must not be rendered
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
            s.active = v.active; s.nseg = nsegs[vb + threadIdx.x]; s.adv = v.adv;
            s.rec = recs[(size_t)tile * n_voices + vb + threadIdx.x];
            sv[threadIdx.x] = s;
        }
        __syncthreads();
        for (uint32_t i = 0; i < nb; ++i) {
            const VoiceS& v = sv[i];
            if (!v.active) continue;
            const Seg* __restrict__ sg = segs + (size_t)(vb + i) * seg_cap;
            const uint32_t tile_step0 = f0 * v.S;
            const bool fast = (v.rec.meta & 0xFFFFu) >= (uint32_t)kFT * v.S + 2u || v.S == 0;
            if (v.C == 2 && v.nch == 2 && v.adv == 0) {
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
                        float p = position_at(v, sg, fast, tile_step0, fl * v.S);
#pragma unroll
                        for (int c = 0; c < OC; ++c) {
                            if ((uint32_t)c < v.nch) {
                                // a retrigger can land between the channels of one frame (processes.rs:82-85)
                                if (v.adv && c > 0) p = position_at(v, sg, fast, tile_step0, fl * v.S + c);
                                const uint32_t idx = f2u_sat(p);
                                if (idx < v.end) acc[j][c] += voice_sample(v.smp + (size_t)idx * v.C + c, v.C, p, v.vel, v.gain);
                            }
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


// ---------------------------------------------------------------- K4 (TMA pipeline, out_channels <= 2)
// Warp-specialised render/mix tile: one producer warp walks the voices of the group and cuts each
// voice's part of the tile into PIECES — maximal frame ranges that lie inside one arithmetic
// position segment (normally one piece = the whole tile; more only where the f32 trajectory crosses
// a binade, retriggers or freezes inside the tile).  For every audible piece it issues ONE bulk async
// copy (cp.async.bulk -> UBLKCP) of the contiguous source span into a 4-deep shared-memory ring,
// signalled through mbarriers.  Eight consumer warps (lanes = consecutive frames) read the staged
// samples with conflict-free LDS, interpolate, scale, cast and accumulate int32 in registers.
// Memory latency is decoupled from arithmetic, arbitrary velocities gather from shared memory instead
// of HBM, and every source byte crosses HBM once per tile.  Spans that do not fit a stage
// (|velocity| > 2 on stereo), NaN trajectories and frames that straddle two segments fall back to
// direct global gathers inside the same kernel.
//
// Measured on B200 (tools/micro/pipe_rates.cu): F2I / I2F.S16 / FRND run on the XU pipe at 16
// lanes/clk/SM, PRMT / LOP3 at 64, FADD / FMUL / FMNMX / IADD at 128.  The hot paths therefore keep
// XU work to the two final saturating casts per frame: i16 -> f32 is PRMT / SHF + I2FP.F32.S32 (32 lanes/clk/SM,
// not XU; see unpack_pair), the per-channel products are FMUL2 pairs, and positions are carried as the integer
// significand of their segment.
constexpr int kStages = 4;
constexpr int kStageBytes = 16 * 1024 + 256;
constexpr int kConsumers = 256;
constexpr int kTmaThreads = kConsumers + 32;
enum : uint32_t { kModeStaged = 1, kModeDirect = 2, kModeEnd = 3, kModeFlush = 4, kModeReduce = 5 };
enum : uint32_t { kPathGeneric = 0, kPathStereoUnit = 1, kPathStereoLerp = 2, kPathStereoMulti = 3, kPathStereoUnit2 = 4 };
constexpr uint32_t kPairHalf = 8192 + 128;    // stage offset of the second voice of a kPathStereoUnit2 item (a full unit tile is <= 8,208 B)
static_assert(2 * kPairHalf <= kStageBytes, "two unit tiles per stage");

struct StageMeta {            // written by the producer before it arrives on the stage's full barrier
    // hot header (one LDS.128)
    uint32_t mode;            // kMode* | path << 8 | full-range flag << 16 | slow flag << 17 | three-segment flag << 18
    float gain;
    uint32_t a0_off;          // byte offset in the stage such that frame fl of the tile maps to a0_off + fl*4 (unit path)
    uint32_t frange;          // fa | fb << 16: the piece covers tile frames [fa, fb)
    // integer trajectory of the fast paths (second LDS.128)
    int32_t q0;               // significand at tile frame 0 (extrapolated), q(fl) = q0 + fl * d
    int32_t d;
    uint32_t sh;              // position = q * 2^-sh
    float scale;              // 2^-sh as float (segment ulp)
    // generic path
    const int16_t* smp;
    const Seg* sg;
    float p0;                 // position at the first step of the piece
    float vel;
    uint32_t end;
    uint32_t base_idx;        // frame index staged at byte_off
    uint32_t byte_off;
    uint32_t shape;           // C | S << 8 | nch << 16
    // offset 72.  A three-segment kPathStereoLerp piece (mode bit 18) keeps (f1, f2, q1, d1, q2, d2) in these six words:
    // q(fl) = q1 + fl * d1 from tile frame f1 on, q2 + fl * d2 from f2 on (consume_stereo_lerp reads them by offset)
    uint32_t nseg;
    uint32_t seg_hint;        // segment index at the piece start (slow pieces walk from here)
    uint32_t f0;              // first frame of the work item's tile (generic path; kModeFlush items carry it in a0_off)
    uint32_t pad_[3];
};
static_assert(sizeof(StageMeta) == 96 && offsetof(StageMeta, nseg) == 72, "StageMeta layout");
constexpr size_t kMetaStride = 96;
constexpr size_t kTmaSmem = (size_t)kStages * kStageBytes + kStages * kMetaStride + 2 * kStages * sizeof(uint64_t) +
                            (size_t)kStages * kMaxPieces * sizeof(uint4) + 32;     // + the producer's uncounted-flush words

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// the same on 32-bit shared addresses computed once (the consumer loop keeps them in registers)
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    // suspend-time hint: a waiting warp sleeps in hardware until the phase completes instead of re-issuing the test
    // (without it the spin was 12.7 % of all issued instructions on C3 + Seq, taken from the warps that had work)
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@!p bra WAIT_%=;\n\t}"
                 ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int32_t lds_s16(uint32_t addr) {
    int32_t v;
    asm volatile("{\n\t.reg .s16 t;\n\tld.shared.s16 t, [%1];\n\tcvt.s32.s16 %0, t;\n\t}" : "=r"(v) : "r"(addr));
    return v;
}

// packed stereo frame (L | R << 16) -> two exact floats without the XU pipe, and with as little of the half-rate ALU pipe
// as possible (on the interpolated scene ALU is the busiest pipe, 66 %, the FMA pipe idles at 21 %:
// profiles/r02_render_c3mixed_full.txt).  L: ((w & 0xFFFF) ^ 0x4B008000) is the float 2^23 + (L + 32768), so one LOP3 and
// one FADD (FMA pipe) give L exactly.  R: arithmetic shift + I2FP.F32.S32 (written so that ptxas cannot pick I2F.S16, which
// runs on the 16-lane XU pipe).  3 ALU + 1 FMA instructions per frame instead of 4 ALU (PRMT, SHF, 2 x I2FP).
__device__ __forceinline__ void unpack_pair(uint32_t w, float& l, float& r) {
    uint32_t lo;
    asm("lop3.b32 %0, %1, 0x0000FFFF, 0x4B008000, 0x6A;" : "=r"(lo) : "r"(w));      // (a & b) ^ c
    // R: 0x4B400000 + sext(R) is the float 2^23 + 2^22 + R (the two's-complement add borrows from mantissa bit 22 for a
    // negative R): ONE LEA.HI.SX32 instead of SHF + I2FP (quarter rate on the ALU pipe).  The two magic constants are
    // subtracted by ONE packed add: -(2^23 + 32768) | -(2^23 + 2^22) << 32, both exact.
    const uint32_t hi = (uint32_t)(((int32_t)w >> 16) + 0x4B400000);
    asm("{\n\t.reg .b64 t, c;\n\tmov.b64 t, {%2, %3};\n\tmov.b64 c, 0xCB400000CB008000;\n\tadd.rn.f32x2 t, t, c;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=f"(l), "=f"(r) : "r"(lo), "r"(hi));
}

// (a * s, b * s), each product rounded to nearest like the scalar FMUL: ONE issue slot (Blackwell FMUL2).  Never
// followed by a packed add: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false, which
// would change the rounding the reference's `s0 * (1 - frac) + s1 * frac` has (scalar adds are used there).
__device__ __forceinline__ void mul2(float a, float b, float s, float& x, float& y) {
    asm("{\n\t.reg .b64 t, u;\n\tmov.b64 t, {%2, %3};\n\tmov.b64 u, {%4, %4};\n\tmul.rn.f32x2 t, t, u;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=f"(x), "=f"(y) : "f"(a), "f"(b), "f"(s));
}
// (frame * gain) as i16 on both channels, added to the accumulators
__device__ __forceinline__ void gain_cast_add(float l, float r, float gain, int32_t& al, int32_t& ar) {
    float x, y;
    mul2(l, r, gain, x, y);
    al += f2i16_sat(x);
    ar += f2i16_sat(y);
}

// velocity == 1.0 inside a unit-step run: frame index advances by one per frame, no interpolation.
// kGainOne: gain == 1.0 — the reference's only value unless the host sets the field (no Command changes it,
// engine.rs:304): (sample as f32 * 1.0) as i16 == sample, so the mix is two sign extensions and two integer adds.
template <bool kFull, bool kGainOne>
__device__ __forceinline__ void consume_stereo_unit(uint32_t stage_addr, uint32_t a0_off, float gain, uint32_t frange,
                                                    int32_t (&acc)[kFPT][2]) {
    const uint32_t a0 = stage_addr + a0_off + threadIdx.x * 4u;
    auto add = [&](uint32_t w, int32_t& al, int32_t& ar) {
        if (kGainOne) {
            al += (int32_t)(int16_t)(w & 0xFFFFu);
            ar += (int32_t)w >> 16;
        } else {
            float l, r;
            unpack_pair(w, l, r);
            gain_cast_add(l, r, gain, al, ar);
        }
    };
    if (kFull) {
        uint32_t w[kFPT];
#pragma unroll
        for (int j = 0; j < kFPT; ++j) w[j] = lds_u32(a0 + (uint32_t)j * kConsumers * 4u);
#pragma unroll
        for (int j = 0; j < kFPT; ++j) add(w[j], acc[j][0], acc[j][1]);
    } else {
        // partial piece: only the 256-frame slabs it touches are visited (uniform skip), lanes predicated
        const uint32_t fa = frange & 0xFFFF, fb = frange >> 16, span = fb - fa;
        const uint32_t j_lo = fa / kConsumers, j_hi = (fb - 1) / kConsumers;
#pragma unroll
        for (int j = 0; j < kFPT; ++j) {
            if ((uint32_t)j < j_lo || (uint32_t)j > j_hi) continue;
            if ((threadIdx.x + j * kConsumers - fa) < span) add(lds_u32(a0 + (uint32_t)j * kConsumers * 4u), acc[j][0], acc[j][1]);
        }
    }
}

// two full unit tiles (two voices) staged in one item: one barrier round trip and one meta read for 16 KB of samples
template <bool kGainOne>
__device__ __forceinline__ void consume_stereo_unit2(uint32_t stage_addr, uint32_t a0_off, float gain, uint32_t a1_off, float gain1,
                                                     int32_t (&acc)[kFPT][2]) {
    const uint32_t a0 = stage_addr + a0_off + threadIdx.x * 4u, a1 = stage_addr + a1_off + threadIdx.x * 4u;
    auto add = [&](uint32_t w, float g, int32_t& al, int32_t& ar) {
        if (kGainOne) {
            al += (int32_t)(int16_t)(w & 0xFFFFu);
            ar += (int32_t)w >> 16;
        } else {
            float l, r;
            unpack_pair(w, l, r);
            gain_cast_add(l, r, g, al, ar);
        }
    };
    uint32_t w[kFPT], x[kFPT];
#pragma unroll
    for (int j = 0; j < kFPT; ++j) w[j] = lds_u32(a0 + (uint32_t)j * kConsumers * 4u);
#pragma unroll
    for (int j = 0; j < kFPT; ++j) x[j] = lds_u32(a1 + (uint32_t)j * kConsumers * 4u);
#pragma unroll
    for (int j = 0; j < kFPT; ++j) add(w[j], gain, acc[j][0], acc[j][1]);
#pragma unroll
    for (int j = 0; j < kFPT; ++j) add(x[j], gain1, acc[j][0], acc[j][1]);
}

// (a * s, b * s), each rounded once, issued as an FMA with a -0.0 addend that only exists at run time (a kernel
// parameter): x * s + (-0.0) == rn(x * s) for every x and s (signed zeros, infinities, NaNs and denormals included).
// As FMAs the products cannot be contracted with the packed add that follows them — ptxas turns mul.rn.f32x2 +
// add.rn.f32x2 into FFMA2 even under -fmad=false, which would round `s0 * (1 - frac) + s1 * frac` once instead of three
// times (engine.rs:430-438) — so the interpolation of a stereo frame is FFMA2, FFMA2, FADD2 and the gain a fourth FFMA2.
__device__ __forceinline__ void mulz2(float a, float b, float s, float nz, float& x, float& y) {
    asm("{\n\t.reg .b64 t, u, v;\n\tmov.b64 t, {%2, %3};\n\tmov.b64 u, {%4, %4};\n\tmov.b64 v, {%5, %5};\n\tfma.rn.f32x2 t, t, u, v;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=f"(x), "=f"(y) : "f"(a), "f"(b), "f"(s), "f"(nz));
}
__device__ __forceinline__ void add2(float a, float b, float c, float d, float& x, float& y) {
    asm("{\n\t.reg .b64 t, u;\n\tmov.b64 t, {%2, %3};\n\tmov.b64 u, {%4, %5};\n\tadd.rn.f32x2 t, t, u;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=f"(x), "=f"(y) : "f"(a), "f"(b), "f"(c), "f"(d));
}
// one interpolated stereo frame: s0 * (1 - frac) + s1 * frac per channel, four separately rounded operations, then
// (x * gain) as i16 (engine.rs:430-442)
__device__ __forceinline__ void lerp_math(uint32_t w0, uint32_t w1, float frac, float om, float gain, float nz,
                                          int32_t& al, int32_t& ar) {
    float l0, r0, l1, r1, a0, b0, a1, b1, x, y;
    unpack_pair(w0, l0, r0);
    unpack_pair(w1, l1, r1);
    mulz2(l0, r0, om, nz, a0, b0);
    mulz2(l1, r1, frac, nz, a1, b1);
    add2(a0, b0, a1, b1, x, y);
    mulz2(x, y, gain, nz, a0, b0);
    al += f2i16_sat(a0);
    ar += f2i16_sat(b0);
}
// any velocity inside one arithmetic segment with positions in [0, 2^24)
__device__ __forceinline__ void lerp_frame(uint32_t w0, uint32_t w1, uint32_t fbits, float scale, float gain, float nz,
                                           int32_t& al, int32_t& ar) {
    // fract = position - trunc(position), exactly (engine.rs:433); fbits < 2^24 so I2FP is exact
    const float frac = __fmul_rn(__int2float_rn((int)fbits), scale);
    lerp_math(w0, w1, frac, __fsub_rn(1.0f, frac), gain, nz, al, ar);
}

// The fraction without a conversion: with sh <= 23 fraction bits, (q & mask) | (150 - sh) << 23 is the float
// 2^(23-sh) + fract, so fract = that - 2^(23-sh) and 1 - fract = (2^(23-sh) + 1) - that, both exact: one LOP3 and two
// FADDs instead of LOP + I2FP (quarter rate) + FMUL + FSUB.
// kThree: the piece is up to THREE consecutive segments of the voice brought to a common unit (the finest ulp 2^-sh among
// them: a position of a coarser binade is a multiple of it as well), i.e. one trajectory whose (q0, d) changes at tile
// frames f1 and f2.  That is the tile in which an interpolated voice crosses a power of two: the run up to the last
// position of the old binade, mostly one odd step, the run in the new binade (build_epoch) — four of the thirteen tiles
// between two retriggers of the C3 + Seq scene.  Two compares and four selects per frame; the table path
// (consume_stereo_multi) is left with the tiles right after a retrigger.
template <bool kFull, bool kThree>
__device__ __forceinline__ void consume_stereo_lerp(uint32_t stage_addr, const StageMeta& m, uint32_t meta_addr, float nz,
                                                    int32_t (&acc)[kFPT][2]) {
    const uint32_t mask = (1u << m.sh) - 1u;
    const uint32_t sbase = stage_addr + m.byte_off - m.base_idx * 4u;
    const int32_t q0 = m.q0, d = m.d;
    const uint32_t sh = m.sh;
    const float gain = m.gain;
    const uint32_t magic = (150u - sh) << 23;
    const float neg_c = -__uint_as_float(magic), one_c = __fadd_rn(__uint_as_float(magic), 1.0f);
    uint32_t f1 = 0, f2 = 0, q1 = 0, d1 = 0, q2 = 0, d2 = 0;
    if (kThree) {
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+72];" : "=r"(f1), "=r"(f2) : "r"(meta_addr));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+80];" : "=r"(q1), "=r"(d1), "=r"(q2), "=r"(d2) : "r"(meta_addr));
    }
    auto q_of = [&](uint32_t fl) -> uint32_t {
        if (kThree) {
            const bool s1 = fl >= f1, s2 = fl >= f2;
            const uint32_t qq = s2 ? q2 : (s1 ? q1 : (uint32_t)q0), dd = s2 ? d2 : (s1 ? d1 : (uint32_t)d);
            return qq + fl * dd;
        }
        return (uint32_t)(q0 + (int32_t)fl * d);
    };
    if (kFull) {
        constexpr int kB = kThree ? 4 : kFPT;          // frames in flight (the three-segment variant has six more live registers)
#pragma unroll
        for (int h = 0; h < kFPT; h += kB) {
            uint32_t w0[kB], w1[kB], fb[kB];
#pragma unroll
            for (int j = 0; j < kB; ++j) {
                const uint32_t q = q_of(threadIdx.x + (uint32_t)(h + j) * kConsumers);
                const uint32_t a = sbase + (q >> sh) * 4u;
                fb[j] = (q & mask) | magic;
                w0[j] = lds_u32(a);
                w1[j] = lds_u32(a + 4u);
            }
#pragma unroll
            for (int j = 0; j < kB; ++j) {
                const float v = __uint_as_float(fb[j]);
                lerp_math(w0[j], w1[j], __fadd_rn(v, neg_c), __fsub_rn(one_c, v), gain, nz, acc[h + j][0], acc[h + j][1]);
            }
        }
    } else {
        const uint32_t fa = m.frange & 0xFFFF, fe = m.frange >> 16, span = fe - fa;
        const uint32_t j_lo = fa / kConsumers, j_hi = (fe - 1) / kConsumers;
#pragma unroll
        for (int j = 0; j < kFPT; ++j) {
            if ((uint32_t)j < j_lo || (uint32_t)j > j_hi) continue;
            const uint32_t fl = threadIdx.x + j * kConsumers;
            if ((fl - fa) < span) {
                const uint32_t q = q_of(fl);
                const uint32_t a = sbase + (q >> sh) * 4u;
                const float v = __uint_as_float((q & mask) | magic);
                lerp_math(lds_u32(a), lds_u32(a + 4u), __fadd_rn(v, neg_c), __fsub_rn(one_c, v), gain, nz, acc[j][0], acc[j][1]);
            }
        }
    }
}

// several pieces of one stereo voice inside one tile, all staged by ONE bulk copy: the producer leaves a table of
// (frame range, q0, d, sh) per piece.  The pieces are consecutive ascending frame ranges and a thread's frames ascend
// with j, so ONE forward walk over the table finds the piece of each of its frames: every frame is evaluated once
// (a loop over the pieces with predicated slabs cost ~90 instructions per piece and warp — 28 % of all instructions of
// C3 + Seq, where the tile after a retrigger crosses ~20 binades), the loads of four frames are issued back to back,
// and a frame outside every audible piece reads as the zero frame ((0 * gain) as i16 == 0, NaN / inf gains included).
template <bool kLerp>
__device__ __forceinline__ void consume_stereo_multi(uint32_t stage_addr, const StageMeta& m, uint32_t ptab, uint32_t n_p, float nz,
                                                     int32_t (&acc)[kFPT][2]) {
    const uint32_t sbase = stage_addr + m.byte_off - m.base_idx * 4u;
    const float gain = m.gain;
    uint32_t k = 0, frange, q0u, du, shf;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(frange), "=r"(q0u), "=r"(du), "=r"(shf) : "r"(ptab));
    constexpr int kHalf = 4;
#pragma unroll
    for (int h = 0; h < kFPT; h += kHalf) {
        uint32_t addr[kHalf], fb[kHalf];
        float sc[kHalf];
#pragma unroll
        for (int jj = 0; jj < kHalf; ++jj) {
            const uint32_t fl = threadIdx.x + (uint32_t)(h + jj) * kConsumers;
            while (k + 1 < n_p && fl >= (frange >> 16)) {
                ++k;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(frange), "=r"(q0u), "=r"(du), "=r"(shf) : "r"(ptab + k * 16u));
            }
            const uint32_t fa = frange & 0xFFFF, fe = frange >> 16, sh = shf & 0xFFu;
            const bool in = (fl - fa) < (fe - fa) && !(shf & 0x100u);            // 0x100: silent piece (frozen / past the end)
            const uint32_t q = q0u + fl * du;
            addr[jj] = in ? sbase + (q >> sh) * 4u : 0xFFFFFFFFu;
            if (kLerp) {
                fb[jj] = q & ((1u << sh) - 1u);
                sc[jj] = __uint_as_float((127u - sh) << 23);
            }
        }
        uint32_t w0[kHalf], w1[kHalf];
#pragma unroll
        for (int jj = 0; jj < kHalf; ++jj) {
            w0[jj] = addr[jj] != 0xFFFFFFFFu ? lds_u32(addr[jj]) : 0u;
            if (kLerp) w1[jj] = addr[jj] != 0xFFFFFFFFu ? lds_u32(addr[jj] + 4u) : 0u;
        }
#pragma unroll
        for (int jj = 0; jj < kHalf; ++jj) {
            if (kLerp) {
                lerp_frame(w0[jj], w1[jj], fb[jj], sc[jj], gain, nz, acc[h + jj][0], acc[h + jj][1]);
            } else {
                float l, r;
                unpack_pair(w0[jj], l, r);
                gain_cast_add(l, r, gain, acc[h + jj][0], acc[h + jj][1]);
            }
        }
    }
}

// generic piece: any channel layout; samples from the stage (kStaged) or straight from global memory.
// Slow pieces (a frame whose advance events straddle two segments) walk the segment list per step.
template <int OC, bool kStaged>
__device__ __forceinline__ void consume_generic(const StageMeta& m, uint32_t stage_addr, uint32_t f0,
                                                int32_t (&acc)[kFPT][OC]) {
    const uint32_t C = m.shape & 0xFF, S = (m.shape >> 8) & 0xFF, nch = (m.shape >> 16) & 0xFF;
    const uint32_t fa = m.frange & 0xFFFF, span = (m.frange >> 16) - fa;
    const bool slow = (m.mode >> 17) & 1;
    const bool lerp = m.vel != 1.0f;
    const uint32_t adv = m.a0_off;                  // generic pieces carry the voice's advance map here
    const uint32_t sbase = stage_addr + m.byte_off;
    const uint32_t j_lo = fa / kConsumers, j_hi = (fa + span - 1) / kConsumers;
#pragma unroll
    for (int j = 0; j < kFPT; ++j) {
        if ((uint32_t)j < j_lo || (uint32_t)j > j_hi) continue;
        const uint32_t fl = threadIdx.x + j * kConsumers;
        if ((fl - fa) < span) {
#pragma unroll
            for (int c = 0; c < OC; ++c) {
                if ((uint32_t)c < nch) {
                    // steps since the piece start; voices with Seq processes step per call (adv != 0)
                    const uint32_t step = (fl - fa) * S + ((C == 1 || adv) ? (uint32_t)c : 0u);
                    const uint32_t abs0 = (f0 + fa) * S;
                    float p;
                    if (!slow) {
                        p = seg_eval(m.p0, m.d, m.scale, adv ? adv_count(abs0 + step, adv) - adv_count(abs0, adv) : step);
                    } else {
                        const uint32_t abs_step = abs0 + step;
                        uint32_t k = m.seg_hint;
                        while (k + 1 < m.nseg && m.sg[k + 1].step0 <= abs_step) ++k;
                        p = seg_pos(m.sg[k], abs_step, adv);
                    }
                    const uint32_t idx = f2u_sat(p);
                    if (idx < m.end) {
                        const uint32_t sc = (C == 1) ? 0u : (uint32_t)c;
                        float s0, s1 = 0.0f;
                        if (kStaged) {
                            const uint32_t a = sbase + ((idx - m.base_idx) * C + sc) * 2u;
                            s0 = (float)lds_s16(a);
                            if (lerp) s1 = (float)lds_s16(a + C * 2u);
                        } else {
                            const int16_t* sp = m.smp + (size_t)idx * C + sc;
                            s0 = (float)sp[0];
                            if (lerp) s1 = (float)sp[C];
                        }
                        float smp = s0;
                        if (lerp) {
                            const float frac = __fsub_rn(p, truncf(p));                  // f32::fract
                            smp = __fadd_rn(__fmul_rn(s0, __fsub_rn(1.0f, frac)), __fmul_rn(s1, frac));
                        }
                        acc[j][c] += f2i16_sat(__fmul_rn(smp, m.gain));
                    }
                }
            }
        }
    }
}

// Persistent: gridDim.x = min(work items, 3 per SM).  A work item is (tile, voice group); the producer warp takes the
// next one from a global counter and keeps the stage ring full ACROSS items, so a CTA has no prologue / epilogue
// bubble per item (as one CTA per item this cost ~11 us per CTA round: C2's mix ran at 5.3 TB/s, C3 at 6.7).  A
// kModeFlush item at the end of each work item makes the consumers add their accumulators into the bus.
// kSink: the bus-sink code (tile counts, reduce items) is compiled in; kTables: so is the producer's piece-table path.
// Both are rare configurations and both cost the producer warp registers (it spills) — the plain kernel carries neither.
template <int OC, bool kSink, bool kTables>
__global__ void __launch_bounds__(kTmaThreads, 3)
voice_render_mix_tma(const VoiceDev* __restrict__ voices, uint32_t n_voices, uint32_t voices_per_group,
                     uint32_t n_groups, const Seg* __restrict__ segs, const uint32_t* __restrict__ nsegs,
                     const TileRec* __restrict__ recs, uint32_t frames, int32_t* __restrict__ bus, int use_atomic,
                     const uint32_t* __restrict__ err, const uint32_t seg_cap, uint32_t* __restrict__ work,
                     const __grid_constant__ BusSink sink, const float nz, const uint4* __restrict__ pool,
                     const uint32_t pool_rows, const uint32_t opts) {
    extern __shared__ __align__(128) uint8_t smem[];
    if (*err) return;                                           // truncated trajectories must not be rendered (uniform exit)
    uint8_t* stages = smem;
    uint8_t* meta_base = smem + (size_t)kStages * kStageBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(meta_base + kStages * kMetaStride);
    uint64_t* empty = full + kStages;
    uint4* ptabs = reinterpret_cast<uint4*>(empty + kStages);          // [kStages][kMaxPieces]
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // With a sink the queue runs `lag` virtual tiles past the last one: the item that renders group 0 of tile T also
    // reduces tile T - lag when this rank owns it.  Every work item a reduction waits for (here or on a peer) has a smaller
    // index than its own, i.e. is already held by a running CTA: no CTA ever waits for work nobody has taken.
    const uint32_t n_tiles = (frames + (uint32_t)kFT - 1u) / (uint32_t)kFT;
    const uint32_t n_render_items = n_tiles * n_groups;
    const uint32_t n_items = (kSink && sink.world) ? (n_tiles + sink.lag) * n_groups : n_render_items;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumers / 32) {
        // ------------------------------------------------ producer warp
        uint32_t o = 0;                                   // stage items enqueued so far (uniform across the warp)
        uint32_t grabbed = 0;                             // lane 0: the next work item, taken one item ahead
        if (lane == 0) grabbed = atomicAdd(work, 1u);
        // Tile hand-off (bus sink).  A flush stage makes the consumers write the work item into the bus; once they have
        // released that stage (its empty barrier) all of their bus writes are issued, and THIS warp — which mostly waits —
        // counts the item with a release atomic and publishes the tile when it was the last one.  The consumers pay
        // nothing.  s_ft[stage] = tile of a flush stage that has not been counted yet.
        volatile uint32_t* s_ft = reinterpret_cast<volatile uint32_t*>(ptabs + kStages * kMaxPieces);
        if (lane == 0)
            for (int k = 0; k < kStages; ++k) s_ft[k] = 0xFFFFFFFFu;
        __syncwarp();
        // one lane: wait until fill number o_idx may overwrite its stage; count the flush it replaces
        auto acquire = [&](uint32_t o_idx) {
            const uint32_t st = o_idx % kStages, round = o_idx / kStages;
            if (round > 0) mbar_wait(empty + st, (round - 1) & 1);
            if (kSink && sink.world) {
                const uint32_t t = s_ft[st];
                if (t != 0xFFFFFFFFu) { k4_tile_flushed(sink, t, n_groups); s_ft[st] = 0xFFFFFFFFu; }
            }
        };
        // one lane: every flush stage enqueued so far is consumed and counted (before this CTA's consumers are sent to
        // wait for a peer, and at the end): a CTA never waits for a tile while it holds a count back
        auto drain = [&](uint32_t o_now) {
            if (!kSink || !sink.world) return;
            for (uint32_t k = 1; k <= (uint32_t)kStages && k <= o_now; ++k) {
                const uint32_t oi = o_now - k, st = oi % kStages;
                const uint32_t t = s_ft[st];
                if (t != 0xFFFFFFFFu) {
                    mbar_wait(empty + st, (oi / kStages) & 1);
                    k4_tile_flushed(sink, t, n_groups);
                    s_ft[st] = 0xFFFFFFFFu;
                }
            }
        };
        auto prefetch_batch = [&](uint32_t t, uint32_t vi_, uint32_t vend_) {
            // the voice row, its tile record and its segment count of a batch the producer will cut later: without this
            // the two dependent misses (voice, then record: ~2 us under load) at every batch start outlast the ring's slack
            if (vi_ < vend_) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(voices + vi_));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(recs + (size_t)t * n_voices + vi_));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nsegs + vi_));
            }
        };
        for (;;) {
        // tile-major item order: all voice groups of tile 0 first.  Early tiles are the expensive ones (a voice that
        // starts at position 0 crosses ~20 binades inside its first tile), so they must not be taken last.
        const uint32_t item = __shfl_sync(0xFFFFFFFFu, grabbed, 0);
        if (item >= n_items) break;
        if (lane == 0) grabbed = atomicAdd(work, 1u);     // the item after this one: in flight while this one is staged
        const uint32_t tile = item / n_groups;
        const bool render_item = !kSink || tile < n_tiles;     // (virtual items exist only behind a sink)
        const uint32_t f0 = tile * (uint32_t)kFT;
        const uint32_t nf = render_item ? min((uint32_t)kFT, frames - f0) : 0u;
        const uint32_t vbeg = (item % n_groups) * voices_per_group;
        const uint32_t vend = render_item ? min(n_voices, vbeg + voices_per_group) : vbeg;
        for (uint32_t vb = vbeg; vb < vend; vb += 32) {
            const uint32_t vi = vb + lane;
            if (vb + 32 < vend) {
                prefetch_batch(tile, vi + 32, vend);
            } else {                                      // last batch of this item: the first batch of the next one
                const uint32_t ni = __shfl_sync(0xFFFFFFFFu, grabbed, 0);
                if (ni < n_render_items) {
                    const uint32_t vb_n = (ni % n_groups) * voices_per_group;
                    prefetch_batch(ni / n_groups, vb_n + lane, min(n_voices, vb_n + voices_per_group));
                }
            }
            // per-lane voice state for the piece walk
            VoiceDev v{};
            const Seg* sg = nullptr;
            uint32_t nseg = 0, seg_j = 0, cur = nf;        // cur = next tile frame not yet covered
            TileRec r{};
            if (vi < vend) {
                v = voices[vi];
                if (v.active) {
                    r = recs[(size_t)tile * n_voices + vi];
                    sg = segs + (size_t)vi * seg_cap;
                    nseg = nsegs[vi];
                    seg_j = r.meta >> 16;
                    cur = 0;
                }
            }
            // ---- issue the pieces of this round: items in lane order, two consecutive full unit tiles share one stage (adds
            // commute, the lanes need not be adjacent).  Every lane that owns an item does all of it itself — wait for
            // its stage, write the meta row, arm the barrier, start the copies — and kStages lanes do so at a time: the
            // stages of one such wave are distinct and depend only on earlier waves.  (One item at a time, with the
            // lanes taking turns between warp barriers, cost ~110 issue slots per item in a warp that gets one slot in
            // thirteen cycles: a third of the time of the single producer warp K4 waits for on scenes with retriggers.)
            auto issue_round = [&](StageMeta& m, const uint32_t bytes, const unsigned long long src) {
                const uint32_t have = __ballot_sync(0xFFFFFFFFu, (m.mode & 0xFF) != 0);
                constexpr uint32_t kUnitFull = kModeStaged | (kPathStereoUnit << 8) | (1u << 16);
                const uint32_t pairable = OC == 2 ? __ballot_sync(0xFFFFFFFFu, m.mode == kUnitFull && bytes <= kPairHalf) : 0u;
                const uint32_t lt = (1u << lane) - 1u;
                const bool is_pair = (pairable >> lane) & 1u;
                const bool second = is_pair && (__popc(pairable & lt) & 1);
                const uint32_t above = pairable & ~lt & ~(1u << lane);
                const int partner = (is_pair && !second && above) ? __ffs(above) - 1 : -1;
                const uint32_t items = have & ~__ballot_sync(0xFFFFFFFFu, second);
                const uint32_t irank = __popc(items & lt), n_it = __popc(items);
                const bool mine = (items >> lane) & 1u;
                const int from = partner >= 0 ? partner : (int)lane;
                const unsigned long long src2 = __shfl_sync(0xFFFFFFFFu, src, from);
                const uint32_t bytes2 = __shfl_sync(0xFFFFFFFFu, bytes, from), a0_2 = __shfl_sync(0xFFFFFFFFu, m.a0_off, from);
                const float gain2 = __shfl_sync(0xFFFFFFFFu, m.gain, from);
                if (mine && partner >= 0) {           // the second voice's stage offset and gain ride in the trajectory slots
                    m.mode = kModeStaged | (kPathStereoUnit2 << 8);
                    m.q0 = (int32_t)(kPairHalf + a0_2);
                    m.scale = gain2;
                }
                for (uint32_t w0 = 0; w0 < n_it; w0 += (uint32_t)kStages) {
                    if (mine && irank - w0 < (uint32_t)kStages) {
                        const uint32_t oi = o + irank, st = oi % kStages;
                        acquire(oi);
                        *reinterpret_cast<StageMeta*>(meta_base + st * kMetaStride) = m;
                        if ((m.mode & 0xFF) == kModeStaged) {
                            mbar_arrive_expect_tx(full + st, bytes + (partner >= 0 ? bytes2 : 0u));
                            bulk_g2s(stages + (size_t)st * kStageBytes, reinterpret_cast<const void*>(src), bytes, full + st);
                            if (partner >= 0)
                                bulk_g2s(stages + (size_t)st * kStageBytes + kPairHalf, reinterpret_cast<const void*>(src2), bytes2, full + st);
                        } else {
                            mbar_arrive(full + st);
                        }
                    }
                    __syncwarp();
                }
                o += n_it;
            };
            bool first_round = true;
            bool hold_multi = false;      // stereo voice whose tile spans several segments: try ONE staged item first
            while (__any_sync(0xFFFFFFFFu, cur < nf)) {
                // ---- every lane cuts its next piece
                StageMeta m{};
                uint32_t bytes = 0;
                unsigned long long src = 0;
                bool unit_tile = false;
                // (voices with Seq processes step per call and their segments also end at retriggers: for them the
                // tile must lie inside ONE segment; a stereo voice on a stereo bus advances once per frame either way)
                if (first_round && cur < nf && OC == 2 && v.C == 2 && v.nch == 2 && v.vel == 1.0f && r.p0 >= 0.0f &&
                    ((v.first_seq >> 24) == 0 || (r.meta & 0xFFFFu) >= nf * v.S)) {
                    // velocity 1.0: while position + frames stays below 2^24 and the start is a multiple of the
                    // coarsest ulp it will meet, every `position += 1.0` is exact, whatever binades it crosses:
                    // the whole tile is one unit-step piece (frame index = floor(p0) + frame).
                    const float top = __fadd_rn(r.p0, (float)nf);
                    if (top < 16777216.0f) {
                        const float u = ulp_of_binade((__float_as_uint(top) >> 23) & 0xFF);
                        const float qf = __fdiv_rn(r.p0, u);
                        const uint32_t idx_a = f2u_sat(r.p0);
                        // a plain voice that runs into its end inside the tile (a clip exactly as long as the render: the
                        // LAST tile of every voice) plays unit steps up to the freeze and is silent after it: still one
                        // piece.  (Left to the multi-piece path, whose table the producer builds one voice at a time, those
                        // tiles — the last work items of the kernel — were a ~40 us tail on C2's mix.)
                        uint32_t n_aud = nf;
                        if (idx_a + nf > v.end) n_aud = ((v.first_seq >> 24) == 0 && idx_a < v.end) ? v.end - idx_a : 0u;
                        if (truncf(qf) == qf && n_aud > 0) {
                            const unsigned long long base = (unsigned long long)v.smp;
                            const unsigned long long b0 = base + (unsigned long long)idx_a * 4ull;
                            const unsigned long long b1 = base + ((unsigned long long)idx_a + n_aud) * 4ull;
                            const unsigned long long a0 = b0 & ~15ull, a1 = (b1 + 15ull) & ~15ull;
                            unit_tile = true;
                            src = a0;
                            bytes = (uint32_t)(a1 - a0);
                            m.a0_off = (uint32_t)(b0 - a0);
                            m.mode = kModeStaged | (kPathStereoUnit << 8) | ((n_aud == (uint32_t)kFT ? 1u : 0u) << 16);
                            m.gain = v.gain;
                            m.frange = 0u | (n_aud << 16);
                            cur = nf;
                        }
                    }
                }
                if (first_round && !unit_tile && cur < nf && OC == 2 && v.C == 2 && v.nch == 2 && v.adv == 0 &&
                    (r.meta & 0xFFFFu) < nf * v.S)
                    hold_multi = true;
                if (!unit_tile && cur < nf && !hold_multi) {
                    uint32_t fa = cur, fb;
                    float p_a;
                    int32_t d;
                    float scale;
                    bool slow = false;
                    const uint32_t tile_step0 = f0 * v.S;
                    if (v.S == 0) {
                        fb = nf; p_a = r.p0; d = 0; scale = 0.0f;
                    } else if (cur == 0 && (r.meta & 0xFFFFu) >= nf * v.S) {
                        fb = nf; p_a = r.p0; d = r.d; scale = r.scale;           // the common case: one piece
                    } else {
                        const uint32_t abs0 = tile_step0 + cur * v.S;
                        while (seg_j + 1 < nseg && sg[seg_j + 1].step0 <= abs0) ++seg_j;
                        const Seg g = sg[seg_j];
                        const uint32_t seg_end = (seg_j + 1 < nseg) ? sg[seg_j + 1].step0 : 0xFFFFFFFFu;
                        const uint32_t whole = (seg_end - abs0) / v.S;           // frames entirely inside the segment
                        p_a = seg_pos(g, abs0, v.adv);
                        d = g.d; scale = g.scale;
                        if (whole >= 1) {
                            fb = min(nf, cur + whole);
                        } else {
                            fb = cur + 1;                                        // frame straddles two segments
                            slow = true;
                        }
                    }
                    cur = fb;
                    const uint32_t last_step = v.S ? (fb - fa) * v.S - 1 : 0;
                    float p_last;
                    {
                        const uint32_t abs_first = tile_step0 + fa * v.S, abs_last = abs_first + last_step;
                        if (!slow) {
                            p_last = seg_eval(p_a, d, scale, adv_count(abs_last, v.adv) - adv_count(abs_first, v.adv));
                        } else {
                            uint32_t k = seg_j;
                            while (k + 1 < nseg && sg[k + 1].step0 <= abs_last) ++k;
                            p_last = seg_pos(sg[k], abs_last, v.adv);
                        }
                    }
                    const bool weird = (p_a != p_a) || (p_last != p_last) || (v.vel != v.vel);
                    const uint32_t idx_lo = f2u_sat(fminf(p_a, p_last));
                    const uint32_t idx_hi = f2u_sat(fmaxf(p_a, p_last));
                    uint32_t mode = 0, path = kPathGeneric;
                    if (weird) {
                        mode = kModeDirect;
                    } else if (idx_lo < v.end) {
                        const uint32_t hi_c = min(idx_hi, v.end - 1);
                        const unsigned long long base = (unsigned long long)v.smp;
                        const unsigned long long b0 = base + (unsigned long long)idx_lo * v.C * 2ull;
                        const unsigned long long b1 = base + ((unsigned long long)hi_c + 2ull) * v.C * 2ull;
                        const unsigned long long a0 = b0 & ~15ull, a1 = (b1 + 15ull) & ~15ull;
                        if (a1 - a0 <= (unsigned long long)kStageBytes) {
                            mode = kModeStaged;
                            src = a0;
                            bytes = (uint32_t)(a1 - a0);
                            m.base_idx = idx_lo;
                            m.byte_off = (uint32_t)(b0 - a0);
                            // fast consumer paths: one segment, every frame audible, stereo voice on a
                            // stereo bus, positions in [0, 2^24)
                            if (OC == 2 && v.C == 2 && v.nch == 2 && !slow && idx_hi < v.end && p_a >= 0.0f && p_last >= 0.0f) {
                                if (v.vel == 1.0f && __fmul_rn((float)d, scale) == 1.0f) {
                                    path = kPathStereoUnit;
                                    m.a0_off = m.byte_off + (f2u_sat(p_a) - idx_lo - fa) * 4u;
                                } else if (v.vel != 1.0f && d != 0) {
                                    const uint32_t eb = (__float_as_uint(scale) >> 23) & 0xFF;     // scale = 2^(eb-127)
                                    if (eb >= 104 && eb <= 127) {             // <= 23 fraction bits: consume_stereo_lerp's fraction trick
                                        path = kPathStereoLerp;
                                        m.sh = 127 - eb;
                                        const int32_t qa = __float2int_rz(__fmul_rn(p_a, __uint_as_float((127u + m.sh) << 23)));
                                        m.q0 = qa - (int32_t)fa * d;
                                    }
                                }
                            } else if (OC == 2 && v.C == 1 && v.nch == 2 && v.adv == 0 && !slow && idx_hi < v.end && p_a >= 0.0f &&
                                       v.vel == 1.0f && __fmul_rn((float)d, scale) == 1.0f && (idx_lo & 1u) == 0) {
                                // a mono voice at velocity 1.0 on a stereo bus advances on both channels (engine.rs:419-422,
                                // 445-447): L reads sample i0 + 2f, R reads i0 + 2f + 1 — with i0 even that is exactly a packed
                                // (L, R) pair per frame, i.e. the stereo unit path on the same bytes
                                path = kPathStereoUnit;
                                m.a0_off = m.byte_off - fa * 4u;
                            }
                        } else {
                            mode = kModeDirect;
                        }
                    }                                     // else: silent piece, nothing to enqueue
                    const bool fullr = (fa == 0 && fb == (uint32_t)kFT);
                    m.mode = mode | (path << 8) | ((fullr ? 1u : 0u) << 16) | ((slow ? 1u : 0u) << 17);
                    if (path == kPathGeneric) m.a0_off = v.adv;
                    m.gain = v.gain;
                    m.frange = fa | (fb << 16);
                    m.d = d;
                    m.scale = scale;
                    m.smp = v.smp;
                    m.sg = sg;
                    m.p0 = p_a;
                    m.vel = v.vel;
                    m.end = v.end;
                    m.shape = v.C | (v.S << 8) | (v.nch << 16);
                    m.nseg = nseg;
                    m.seg_hint = seg_j;
                    m.f0 = f0;
                }
                // ---- after the first round: voices held back for the multi-piece path, one at a time, with
                // the whole warp cooperating (lane l looks at segment j0 + l of that voice)
                issue_round(m, bytes, src);
                if (first_round && (opts & 4u) && __any_sync(0xFFFFFFFFu, hold_multi)) {
                    // Lane-local: a held tile that two or three segments cover — an interpolated voice crossing a power of
                    // two: four of five multi-segment tiles on scenes with retriggers — is cut by its own lane into ONE
                    // three-segment piece (consume_stereo_lerp<., true>) and issued like the single-segment pieces above,
                    // all lanes at once.  (The cooperative path below takes ~375 issue slots of this warp per voice.)
                    StageMeta m3{};
                    uint32_t bytes3 = 0;
                    unsigned long long src3 = 0;
                    if (hold_multi && v.vel != 1.0f) {
                        const uint32_t last_abs = f0 + nf - 1;                      // S == 1 for held voices
                        uint32_t cnt = 1;
                        Seg g[3];
                        uint32_t nxt[3];
                        g[0] = sg[seg_j];
                        g[1] = g[0]; g[2] = g[0];
                        nxt[0] = nxt[1] = nxt[2] = 0xFFFFFFFFu;
#pragma unroll
                        for (int t = 1; t <= 3; ++t) {
                            if (cnt == (uint32_t)t && seg_j + t < nseg) {
                                const Seg gn = sg[seg_j + t];
                                nxt[t - 1] = gn.step0;
                                if (t < 3 && gn.step0 <= last_abs) { g[t] = gn; cnt = t + 1; }
                            }
                        }
                        const uint32_t nxt_last = cnt == 3u ? nxt[2] : (cnt == 2u ? nxt[1] : nxt[0]);
                        const bool covered = nxt_last > last_abs;                    // no fourth segment inside the tile
                        bool ok3 = covered && cnt >= 2u;
                        uint32_t lo_g = 0xFFFFFFFFu, hi_g = 0u, sh_c = 0u;
                        uint32_t fa_[3] = {0, 0, 0}, fe_[3] = {0, 0, 0}, sh_[3] = {0, 0, 0};
                        int32_t q_[3] = {0, 0, 0};
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            if (ok3 && (uint32_t)t < cnt) {
                                const uint32_t ls = max(g[t].step0, f0), le = min(nxt[t], f0 + nf);
                                fa_[t] = ls - f0; fe_[t] = le - f0;
                                const float p_a = seg_eval(g[t].p0, g[t].d, g[t].scale, ls - g[t].step0);
                                const float p_l = seg_eval(g[t].p0, g[t].d, g[t].scale, le - 1 - g[t].step0);
                                const uint32_t lo = f2u_sat(fminf(p_a, p_l)), hi = f2u_sat(fmaxf(p_a, p_l));
                                // audible, inside the clip, non-negative (NaN fails the comparisons)
                                bool good = p_a >= 0.0f && p_l >= 0.0f && hi < v.end;
                                if (good) {
                                    if (g[t].d != 0) {
                                        const uint32_t eb = (__float_as_uint(g[t].scale) >> 23) & 0xFF;
                                        good = eb >= 104 && eb <= 127;
                                        sh_[t] = 127 - eb;
                                    } else {
                                        const int E = (int)((__float_as_uint(p_a) >> 23) & 0xFF);
                                        const int s_ = p_a == 0.0f ? 0 : max(0, 150 - E);
                                        good = s_ <= 23;
                                        sh_[t] = (uint32_t)s_;
                                    }
                                }
                                if (good) {
                                    const int32_t qa = __float2int_rz(__fmul_rn(p_a, __uint_as_float((127u + sh_[t]) << 23)));
                                    q_[t] = qa - (int32_t)fa_[t] * g[t].d;
                                    lo_g = min(lo_g, lo); hi_g = max(hi_g, hi); sh_c = max(sh_c, sh_[t]);
                                } else {
                                    ok3 = false;
                                }
                            }
                        }
                        if (ok3 && ((hi_g + 2u) >> (31u - sh_c)) == 0u) {
                            const unsigned long long base = (unsigned long long)v.smp;
                            const unsigned long long b0 = base + (unsigned long long)lo_g * 4ull;
                            const unsigned long long b1 = base + ((unsigned long long)hi_g + 2ull) * 4ull;
                            const unsigned long long a0 = b0 & ~15ull, a1 = (b1 + 15ull) & ~15ull;
                            if (a1 - a0 <= (unsigned long long)kStageBytes) {
                                const uint32_t fe_all = cnt == 3u ? fe_[2] : fe_[1];
                                const uint32_t fullr = (fa_[0] == 0u && fe_all == (uint32_t)kFT) ? 1u : 0u;
                                m3.mode = kModeStaged | (kPathStereoLerp << 8) | (fullr << 16) | (1u << 18);
                                m3.gain = v.gain;
                                m3.frange = fa_[0] | (fe_all << 16);
                                m3.base_idx = lo_g;
                                m3.byte_off = (uint32_t)(b0 - a0);
                                m3.sh = sh_c;
                                m3.q0 = (int32_t)((uint32_t)q_[0] << (sh_c - sh_[0]));
                                m3.d = (int32_t)((uint32_t)g[0].d << (sh_c - sh_[0]));
                                m3.nseg = fa_[1];                                                  // f1
                                m3.seg_hint = cnt == 3u ? fa_[2] : fe_all;                          // f2
                                m3.f0 = (uint32_t)q_[1] << (sh_c - sh_[1]);
                                m3.pad_[0] = (uint32_t)g[1].d << (sh_c - sh_[1]);
                                m3.pad_[1] = (uint32_t)q_[2] << ((sh_c - sh_[2]) & 31u);
                                m3.pad_[2] = (uint32_t)g[2].d << ((sh_c - sh_[2]) & 31u);
                                src3 = a0;
                                bytes3 = (uint32_t)(a1 - a0);
                                cur = nf;
                                hold_multi = false;
                            }
                        }
                    }
                    issue_round(m3, bytes3, src3);
                }
                if (first_round) {
                    // tiles whose piece table K3 has already built (TileRec.scale carries the tag): the lane reads its
                    // item headers and issues two bulk copies per item — the source span and the piece rows — nothing
                    // is computed here.  (Cutting these tiles in this warp, one voice at a time, was what K4 waited for on
                    // scenes with retriggers: a quarter of all voice-tiles, ~6 segments each.)
                    {
                        const bool has_tab = kTables && hold_multi && pool != nullptr && (__float_as_uint(r.scale) & 0xFFC00000u) == kTabTag;
                        const uint4* vrows = has_tab ? pool + (size_t)vi * pool_rows + (uint32_t)r.d : nullptr;
                        uint4 hdr = make_uint4(0, 0, 0, 0);
                        if (has_tab) hdr = __ldg(vrows);                            // all lanes' first headers in one round trip
                        for (uint32_t rest = __ballot_sync(0xFFFFFFFFu, has_tab); rest; rest &= rest - 1) {
                            const int i = __ffs(rest) - 1;
                            uint32_t o_mine = o;
                            if ((int)lane == i) {
                                const uint32_t n_it = __float_as_uint(r.scale) & 0x003FFFFFu;
                                const uint4* row = vrows;
                                for (uint32_t it = 0; it < n_it; ++it) {
                                    if (it) hdr = __ldg(row);
                                    const uint32_t lo_g = hdr.x, hi_g = hdr.y, cnt = hdr.z;
                                    const unsigned long long base = (unsigned long long)v.smp;
                                    const unsigned long long b0 = base + (unsigned long long)lo_g * 4ull;
                                    const unsigned long long b1 = base + ((unsigned long long)hi_g + 2ull) * 4ull;
                                    const unsigned long long a0 = b0 & ~15ull, a1 = (b1 + 15ull) & ~15ull;
                                    const uint32_t st = o_mine % kStages;
                                    acquire(o_mine);
                                    StageMeta mm{};
                                    mm.mode = kModeStaged | (kPathStereoMulti << 8);
                                    mm.gain = v.gain;
                                    mm.a0_off = cnt;
                                    mm.frange = (v.vel != 1.0f) ? 1u : 0u;
                                    mm.base_idx = lo_g;
                                    mm.byte_off = (uint32_t)(b0 - a0);
                                    *reinterpret_cast<StageMeta*>(meta_base + st * kMetaStride) = mm;
                                    mbar_arrive_expect_tx(full + st, (uint32_t)(a1 - a0) + cnt * 16u);
                                    bulk_g2s(stages + (size_t)st * kStageBytes, reinterpret_cast<const void*>(a0), (uint32_t)(a1 - a0), full + st);
                                    bulk_g2s(ptabs + st * kMaxPieces, row + 1, cnt * 16u, full + st);
                                    row += 1u + cnt;
                                    o_mine += 1;
                                }
                                cur = nf;
                                hold_multi = false;
                            }
                            o = __shfl_sync(0xFFFFFFFFu, o_mine, i);
                        }
                    }
                    for (uint32_t rest = __ballot_sync(0xFFFFFFFFu, hold_multi); rest; rest &= rest - 1) {
                        const int i = __ffs(rest) - 1;
                        const Seg* sgi = reinterpret_cast<const Seg*>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)sg, i));
                        const uint32_t nseg_i = __shfl_sync(0xFFFFFFFFu, nseg, i), j0 = __shfl_sync(0xFFFFFFFFu, seg_j, i);
                        const uint32_t end_i = __shfl_sync(0xFFFFFFFFu, v.end, i);
                        const float gain_i = __shfl_sync(0xFFFFFFFFu, v.gain, i), vel_i = __shfl_sync(0xFFFFFFFFu, v.vel, i);
                        const unsigned long long smp_i = __shfl_sync(0xFFFFFFFFu, (unsigned long long)v.smp, i);
                        const uint32_t last_abs = f0 + nf - 1;                      // S == 1 for these voices
                        const uint32_t jl = j0 + lane;
                        bool ex = jl < nseg_i;
                        Seg g{};
                        uint32_t nxt = 0xFFFFFFFFu;
                        if (ex) {
                            g = sgi[jl];
                            if (jl + 1 < nseg_i) nxt = sgi[jl + 1].step0;
                        }
                        ex = ex && (lane == 0 || g.step0 <= last_abs);
                        const uint32_t exm = __ballot_sync(0xFFFFFFFFu, ex);
                        const bool overflow = exm == 0xFFFFFFFFu && __shfl_sync(0xFFFFFFFFu, nxt, 31) <= last_abs;
                        bool ok = true, silent = true;
                        uint32_t idx_lo = 0xFFFFFFFFu, idx_hi = 0, shv = 0, fa = 0, fe = 0;
                        int32_t q0v = 0;
                        float p_a = 0.0f;
                        if (ex) {
                            const uint32_t ls = max(g.step0, f0), le = min(nxt, f0 + nf);
                            fa = ls - f0; fe = le - f0;
                            p_a = seg_eval(g.p0, g.d, g.scale, ls - g.step0);
                            const float p_l = seg_eval(g.p0, g.d, g.scale, le - 1 - g.step0);
                            const bool weird = (p_a != p_a) || (p_l != p_l);
                            const uint32_t lo = f2u_sat(fminf(p_a, p_l)), hi = f2u_sat(fmaxf(p_a, p_l));
                            silent = !weird && lo >= end_i;
                            ok = silent;
                            if (!silent && !weird && p_a >= 0.0f && p_l >= 0.0f && hi < end_i) {
                                if (g.d != 0) {
                                    const uint32_t eb = (__float_as_uint(g.scale) >> 23) & 0xFF;
                                    if (eb >= 96 && eb <= 127) { shv = 127 - eb; ok = true; }
                                } else {
                                    const int E = (int)((__float_as_uint(p_a) >> 23) & 0xFF);
                                    const int s_ = p_a == 0.0f ? 0 : max(0, 150 - E);
                                    if (s_ <= 31) { shv = (uint32_t)s_; ok = true; }
                                }
                                if (ok) {
                                    const int32_t qa = __float2int_rz(__fmul_rn(p_a, __uint_as_float((127u + shv) << 23)));
                                    q0v = qa - (int32_t)fa * g.d;
                                    idx_lo = lo; idx_hi = hi;
                                }
                            }
                        }
                        const bool all_ok = __all_sync(0xFFFFFFFFu, ok) && !overflow;
                        // every audible piece must fit a stage on its own; then the pieces are packed, in order, into as few
                        // staged items as possible (normally one; two when a retrigger jumps to a far-away source region)
                        auto span_fits = [&](uint32_t lo, uint32_t hi) -> bool {
                            const unsigned long long b0 = smp_i + (unsigned long long)lo * 4ull;
                            const unsigned long long b1 = smp_i + ((unsigned long long)hi + 2ull) * 4ull;
                            return ((b1 + 15ull) & ~15ull) - (b0 & ~15ull) <= (unsigned long long)kStageBytes;
                        };
                        const bool own_fits = !ex || silent || span_fits(idx_lo, idx_hi);
                        if (all_ok && __all_sync(0xFFFFFFFFu, own_fits)) {
                            const uint32_t n_ex = (uint32_t)__popc(exm);            // segments j0 .. j0 + n_ex - 1 (a contiguous run of lanes)
                            uint32_t g0 = 0;
                            while (g0 < n_ex) {
                                // running min / max of the source range over lanes g0 .. lane
                                uint32_t lo_run = (lane >= g0 && ex) ? idx_lo : 0xFFFFFFFFu;
                                uint32_t hi_run = (lane >= g0 && ex && !silent) ? idx_hi : 0u;
#pragma unroll
                                for (int d = 1; d < 32; d <<= 1) {
                                    const uint32_t ul = __shfl_up_sync(0xFFFFFFFFu, lo_run, d), uh = __shfl_up_sync(0xFFFFFFFFu, hi_run, d);
                                    if (lane >= g0 + (uint32_t)d) { lo_run = min(lo_run, ul); hi_run = max(hi_run, uh); }
                                }
                                const bool fit = lane >= g0 && lane < n_ex && (lo_run == 0xFFFFFFFFu || span_fits(lo_run, hi_run));
                                const uint32_t fitm = __ballot_sync(0xFFFFFFFFu, fit) >> g0;
                                const uint32_t cnt = fitm == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~fitm) - 1u;   // >= 1: every piece fits on its own
                                const uint32_t L = g0 + cnt - 1u;
                                const uint32_t lo_g = __shfl_sync(0xFFFFFFFFu, lo_run, L), hi_g = __shfl_sync(0xFFFFFFFFu, hi_run, L);
                                if (lo_g != 0xFFFFFFFFu) {                                          // something audible in this group
                                    const unsigned long long b0 = smp_i + (unsigned long long)lo_g * 4ull;
                                    const unsigned long long b1 = smp_i + ((unsigned long long)hi_g + 2ull) * 4ull;
                                    const unsigned long long a0 = b0 & ~15ull, a1 = (b1 + 15ull) & ~15ull;
                                    const uint32_t st = o % kStages;
                                    if (lane == 0) acquire(o);
                                    __syncwarp();
                                    // A group of ONE audible segment needs no table: it is an ordinary partial piece, unit or
                                    // interpolated, and the consumers skip the 256-frame slabs it does not touch (the part of a
                                    // tile before a retrigger, whose source lies far from the home position's).  Two or three
                                    // audible segments of an interpolated voice are ONE piece in their finest common unit.
                                    uint32_t kind = 0;                    // 0: table, 1: unit piece, 2: interpolated piece, 3: three segments
                                    const bool in_grp = lane >= g0 && lane <= L;
                                    const bool lerp_seg = in_grp && ex && !silent && vel_i != 1.0f;       // (ok holds for every lane here)
                                    uint32_t sh_c = lerp_seg ? shv : 0u;
#pragma unroll
                                    for (int dd = 16; dd >= 1; dd >>= 1) sh_c = max(sh_c, __shfl_xor_sync(0xFFFFFFFFu, sh_c, dd));
                                    const bool all_lerp = __all_sync(0xFFFFFFFFu, !in_grp || lerp_seg);
                                    // my segment in the common unit (wrap-around arithmetic: the true q of an in-range frame is below 2^31)
                                    const uint32_t q_c = (uint32_t)q0v << ((sh_c - shv) & 31u), d_c = (uint32_t)g.d << ((sh_c - shv) & 31u);
                                    const uint32_t l1 = min(g0 + 1u, L), l2 = min(g0 + 2u, L);
                                    const uint32_t fa_1 = __shfl_sync(0xFFFFFFFFu, fa, l1), q_1 = __shfl_sync(0xFFFFFFFFu, q_c, l1),
                                                   d_1 = __shfl_sync(0xFFFFFFFFu, d_c, l1);
                                    const uint32_t fa_2 = __shfl_sync(0xFFFFFFFFu, fa, l2), q_2 = __shfl_sync(0xFFFFFFFFu, q_c, l2),
                                                   d_2 = __shfl_sync(0xFFFFFFFFu, d_c, l2);
                                    const uint32_t fe_L = __shfl_sync(0xFFFFFFFFu, fe, L);
                                    if (lane == g0) {
                                        if (cnt == 1u && (opts & 1u)) {
                                            if (vel_i == 1.0f && __fmul_rn((float)g.d, g.scale) == 1.0f) kind = 1;
                                            else if (lerp_seg && g.d != 0 && shv <= 23u) kind = 2;
                                        } else if ((cnt == 2u || cnt == 3u) && (opts & 2u) && all_lerp && sh_c <= 23u &&
                                                   ((hi_g + 2u) >> (31u - sh_c)) == 0u) {
                                            kind = 3;
                                        }
                                        if (kind != 0u) {
                                            StageMeta mm{};
                                            const uint32_t fe_all = kind == 3u ? fe_L : fe;
                                            const uint32_t fullr = (fa == 0u && fe_all == (uint32_t)kFT) ? 1u : 0u;
                                            mm.gain = gain_i;
                                            mm.frange = fa | (fe_all << 16);
                                            mm.base_idx = lo_g;
                                            mm.byte_off = (uint32_t)(b0 - a0);
                                            if (kind == 1u) {
                                                mm.mode = kModeStaged | (kPathStereoUnit << 8) | (fullr << 16);
                                                mm.a0_off = mm.byte_off + (f2u_sat(p_a) - lo_g - fa) * 4u;
                                            } else if (kind == 2u) {
                                                mm.mode = kModeStaged | (kPathStereoLerp << 8) | (fullr << 16);
                                                mm.sh = shv;
                                                mm.q0 = q0v;
                                                mm.d = g.d;
                                            } else {
                                                mm.mode = kModeStaged | (kPathStereoLerp << 8) | (fullr << 16) | (1u << 18);
                                                mm.sh = sh_c;
                                                mm.q0 = (int32_t)q_c;
                                                mm.d = (int32_t)d_c;
                                                mm.nseg = fa_1;                          // f1
                                                mm.seg_hint = cnt == 3u ? fa_2 : fe_L;   // f2 (nothing switches at the end of the piece)
                                                mm.f0 = q_1;
                                                mm.pad_[0] = d_1;
                                                mm.pad_[1] = q_2;
                                                mm.pad_[2] = d_2;
                                            }
                                            *reinterpret_cast<StageMeta*>(meta_base + st * kMetaStride) = mm;
                                        }
                                    }
                                    kind = __shfl_sync(0xFFFFFFFFu, kind, g0);
                                    if (kind == 0u && ex && lane >= g0 && lane <= L)
                                        ptabs[st * kMaxPieces + lane - g0] = make_uint4(fa | (fe << 16), (uint32_t)q0v, (uint32_t)g.d, shv | (silent ? 0x100u : 0u));
                                    if (kind == 0u && lane == 0) {
                                        StageMeta mm{};
                                        mm.mode = kModeStaged | (kPathStereoMulti << 8);
                                        mm.gain = gain_i;
                                        mm.a0_off = cnt;
                                        mm.frange = (vel_i != 1.0f) ? 1u : 0u;
                                        mm.base_idx = lo_g;
                                        mm.byte_off = (uint32_t)(b0 - a0);
                                        *reinterpret_cast<StageMeta*>(meta_base + st * kMetaStride) = mm;
                                    }
                                    __syncwarp();
                                    if (lane == 0) {
                                        mbar_arrive_expect_tx(full + st, (uint32_t)(a1 - a0));
                                        bulk_g2s(stages + (size_t)st * kStageBytes, reinterpret_cast<const void*>(a0), (uint32_t)(a1 - a0), full + st);
                                    }
                                    o += 1;
                                }
                                g0 = L + 1u;
                            }
                            if (lane == i) cur = nf;
                        }
                    }
                    hold_multi = false;            // whoever is left walks piece by piece below
                }
                first_round = false;
            }
        }
        if (render_item) {   // end of the work item: the consumers add their accumulators into the bus
            const uint32_t st = o % kStages;
            if (lane == 0) {
                acquire(o);
                StageMeta* ms = reinterpret_cast<StageMeta*>(meta_base + st * kMetaStride);
                ms->mode = kModeFlush;
                ms->a0_off = f0;
                ms->frange = nf;
                if (kSink && sink.world) s_ft[st] = tile;
                mbar_arrive(full + st);
            }
            __syncwarp();
            o += 1;
        }
        if (kSink && sink.world && item % n_groups == 0 && tile >= sink.lag && (tile - sink.lag) % sink.world == sink.rank) {
            const uint32_t st = o % kStages;
            if (lane == 0) {
                drain(o);
                acquire(o);
                StageMeta* ms = reinterpret_cast<StageMeta*>(meta_base + st * kMetaStride);
                ms->mode = kModeReduce;
                ms->a0_off = tile - sink.lag;
                mbar_arrive(full + st);
            }
            __syncwarp();
            o += 1;
        }
        }   // next work item
        if (lane == 0) {
            const uint32_t st = o % kStages;
            drain(o);
            acquire(o);
            reinterpret_cast<StageMeta*>(meta_base + st * kMetaStride)->mode = kModeEnd;
            mbar_arrive(full + st);
        }
        return;
    }

    // ---------------------------------------------------- consumer warps
    int32_t acc[kFPT][OC];
#pragma unroll
    for (int j = 0; j < kFPT; ++j)
#pragma unroll
        for (int c = 0; c < OC; ++c) acc[j][c] = 0;

    // one opaque shared-window base (asm volatile: ptxas would otherwise rebuild it from SR_CgaCtaId every iteration)
    uint32_t sm0;
    asm volatile("mov.u32 %0, %1;" : "=r"(sm0) : "r"(smem_u32(smem)));
    constexpr uint32_t kMetaOff = (uint32_t)kStages * kStageBytes, kFullOff = kMetaOff + (uint32_t)(kStages * kMetaStride),
                       kEmptyOff = kFullOff + (uint32_t)kStages * 8u;
    uint32_t st = 0, phase = 0;
    for (;;) {
        mbar_wait_a(sm0 + kFullOff + st * 8u, phase);
        const uint32_t stage_addr = sm0 + st * (uint32_t)kStageBytes;
        const uint32_t meta_addr = sm0 + kMetaOff + st * (uint32_t)kMetaStride;
        const StageMeta* mp = reinterpret_cast<const StageMeta*>(meta_base + st * kMetaStride);
        uint32_t mode; float gain; uint32_t a0_off, frange;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(mode), "=f"(gain), "=r"(a0_off), "=r"(frange) : "r"(meta_addr));
        if (OC == 2 && mode == (kModeStaged | (kPathStereoUnit2 << 8))) {            // the common item first: one compare
            auto& a2 = reinterpret_cast<int32_t (&)[kFPT][2]>(acc);
            uint32_t a1_off, d_, sh_; float gain1;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(a1_off), "=r"(d_), "=r"(sh_), "=f"(gain1) : "r"(meta_addr));
            if (gain == 1.0f && gain1 == 1.0f) consume_stereo_unit2<true>(stage_addr, a0_off, gain, a1_off, gain1, a2);
            else consume_stereo_unit2<false>(stage_addr, a0_off, gain, a1_off, gain1, a2);
            __syncwarp();
            if (lane == 0) mbar_arrive_a(sm0 + kEmptyOff + st * 8u);
            st = (st + 1) % kStages;
            phase ^= (st == 0) ? 1u : 0u;
            continue;
        }
        if ((mode & 0xFF) == kModeEnd) break;
        if ((mode & 0xFF) == kModeFlush) {
            // end of a work item: a0_off = first frame of its tile, frange = frames in it
#pragma unroll
            for (int j = 0; j < kFPT; ++j) {
                const uint32_t fl = threadIdx.x + j * kConsumers;
                if (fl < frange) {
                    int32_t* out = bus + (size_t)(a0_off + fl) * OC;
#pragma unroll
                    for (int c = 0; c < OC; ++c) {
                        if (use_atomic) atomicAdd(out + c, acc[j][c]);
                        else out[c] = acc[j][c];
                        acc[j][c] = 0;
                    }
                }
            }
        } else if (kSink && (mode & 0xFF) == kModeReduce) {
            k4_reduce_tile(sink, a0_off);
        } else {
            const uint32_t path = (mode >> 8) & 0xFF;
            const bool fullr = (mode >> 16) & 1;
            if (OC == 2 && path == kPathStereoUnit) {
                auto& a2 = reinterpret_cast<int32_t (&)[kFPT][2]>(acc);
                const bool g1 = gain == 1.0f;
                if (fullr) { if (g1) consume_stereo_unit<true, true>(stage_addr, a0_off, gain, frange, a2); else consume_stereo_unit<true, false>(stage_addr, a0_off, gain, frange, a2); }
                else { if (g1) consume_stereo_unit<false, true>(stage_addr, a0_off, gain, frange, a2); else consume_stereo_unit<false, false>(stage_addr, a0_off, gain, frange, a2); }
            } else if (OC == 2 && path == kPathStereoLerp) {
                auto& a2 = reinterpret_cast<int32_t (&)[kFPT][2]>(acc);
                if ((mode >> 18) & 1u) {
                    if (fullr) consume_stereo_lerp<true, true>(stage_addr, *mp, meta_addr, nz, a2);
                    else consume_stereo_lerp<false, true>(stage_addr, *mp, meta_addr, nz, a2);
                } else {
                    if (fullr) consume_stereo_lerp<true, false>(stage_addr, *mp, meta_addr, nz, a2);
                    else consume_stereo_lerp<false, false>(stage_addr, *mp, meta_addr, nz, a2);
                }
            } else if (OC == 2 && path == kPathStereoMulti) {
                auto& a2 = reinterpret_cast<int32_t (&)[kFPT][2]>(acc);
                const uint32_t ptab = smem_u32(ptabs + st * kMaxPieces);
                if (frange & 1u) consume_stereo_multi<true>(stage_addr, *mp, ptab, a0_off, nz, a2);
                else consume_stereo_multi<false>(stage_addr, *mp, ptab, a0_off, nz, a2);
            } else if ((mode & 0xFF) == kModeStaged) {
                consume_generic<OC, true>(*mp, stage_addr, mp->f0, acc);
            } else {
                consume_generic<OC, false>(*mp, 0u, mp->f0, acc);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_a(sm0 + kEmptyOff + st * 8u);
        st = (st + 1) % kStages;
        phase ^= (st == 0) ? 1u : 0u;
    }
}

// ---------------------------------------------------------------- K4b
// Split frames of voices with Seq processes: a retrigger that landed between the channels of frame f.  K4 rendered
// every channel of that frame from the NEW epoch (home position); the channels before the hit must carry the old
// position's sample instead: add (old - new) for them.  One thread per (voice, split).
__global__ void voice_split_fixup(const VoiceDev* __restrict__ voices, uint32_t n_voices, const Split* __restrict__ splits,
                                  const uint32_t* __restrict__ nsplits, uint32_t oc, int32_t* __restrict__ bus,
                                  const uint32_t* __restrict__ err) {
    if (*err) return;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t vi = idx / (uint32_t)kMaxEvents, si = idx % (uint32_t)kMaxEvents;
    if (vi >= n_voices || si >= nsplits[vi]) return;
    const VoiceDev v = voices[vi];
    const Split s = splits[(size_t)vi * kMaxEvents + si];
    const float home = v.vel >= 0.0f ? 0.0f : (float)v.end;
    const uint32_t i_old = f2u_sat(s.p_old), i_new = f2u_sat(home);
    for (uint32_t ch = 0; ch < s.k && ch < v.nch; ++ch) {
        int32_t delta = 0;
        if (i_old < v.end) delta += voice_sample(v.smp + (size_t)i_old * v.C + ch, v.C, s.p_old, v.vel, v.gain);
        if (i_new < v.end) delta -= voice_sample(v.smp + (size_t)i_new * v.C + ch, v.C, home, v.vel, v.gain);
        if (delta) atomicAdd(bus + (size_t)s.frame * oc + ch, delta);
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

// ---------------------------------------------------------------- K5p: the bus reduction on its own
// The same tile protocol as inside K4 (sink_tile_flushed / sink_reduce_tile) for partial buses that were filled by
// earlier stream work — the Conductor's spans, buses with more than two channels, a rank without voices:
// `peer_publish_tiles` raises this rank's ready flag of every tile at the tile's owner, `bus_reduce_tiles` reduces the
// tiles this rank owns, one CTA each.
__global__ void peer_publish_tiles(const BusSink sink) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= sink.n_tiles) return;
    __threadfence_system();                         // everything this stream wrote before is visible before the flags
    st_release_sys(sink.ready_at[t % sink.world] + t, sink.step);
}

__global__ void __launch_bounds__(256)
bus_reduce_tiles(const BusSink sink) {
    const uint32_t tile = sink.rank + blockIdx.x * sink.world;
    if (tile >= sink.n_tiles) return;
    sink_reduce_tile(sink, tile, threadIdx.x, 256u, [] { __syncthreads(); });
}

// a rank that owns no tile of this step (fewer tiles than ranks) still tells the root "my part is in place" and every
// rank "I am not reading your partial bus"
__global__ void peer_raise(const BusSink sink) {
    __threadfence_system();
    if (threadIdx.x < sink.n_done) st_release_sys(sink.done[threadIdx.x], sink.step);
}

__global__ void flag_wait(const uint32_t* flags, uint32_t n, uint32_t value, uint32_t timeout_ms, uint32_t* err) {
    if (threadIdx.x < n) wait_flag(flags + threadIdx.x, value, timeout_ms, err);
}

}  // namespace

namespace blast_rdr {

void route_voice(VoiceDev& v, uint32_t oc, bool has_seq) {
    uint32_t lo = 0, na = 0;
    if (v.C == 1) {                 // engine.rs:419-422: bus channels 0 and 1 both read (and advance) a mono voice
        v.nch = oc < 2 ? oc : 2;
        v.S = v.nch;
        lo = 0; na = v.nch;
    } else if (oc >= v.C) {         // engine.rs:425-427, 445-447
        v.nch = v.C;
        v.S = 1;
        lo = v.C - 1; na = 1;
    } else {                        // ch == C-1 never happens: the voice never advances
        v.nch = oc;
        v.S = 0;
    }
    v.adv = 0;
    if (has_seq && v.S == 0) {      // a voice that never advances but can be retriggered: steps become calls
        v.adv = oc | (lo << 8) | (na << 16);
        v.S = oc;
    }
}

void free_buffers(RenderBuffers& rb) {
    if (rb.d_voices) cudaFree(rb.d_voices);
    if (rb.d_segs) cudaFree(rb.d_segs);
    if (rb.d_nsegs) cudaFree(rb.d_nsegs);
    if (rb.d_err) cudaFree(rb.d_err);
    if (rb.d_recs) cudaFree(rb.d_recs);
    if (rb.d_pool) cudaFree(rb.d_pool);
    if (rb.d_seqs) cudaFree(rb.d_seqs);
    if (rb.d_events) cudaFree(rb.d_events);
    if (rb.d_nevents) cudaFree(rb.d_nevents);
    if (rb.d_splits) cudaFree(rb.d_splits);
    if (rb.d_nsplits) cudaFree(rb.d_nsplits);
    rb = RenderBuffers{};
}

int reserve_buffers(blast_ctx* ctx, RenderBuffers& rb, uint32_t n_voices, uint32_t n_seqs) {
    const size_t nv = n_voices ? n_voices : 1;
    // voices with Seq processes start a new run of ~20 segments at every retrigger: give them room for ~50 of them
    const uint32_t seg_cap = n_seqs ? (uint32_t)kMaxSegSeq : (uint32_t)kMaxSeg;
    if (nv > rb.voices_cap || seg_cap > rb.seg_cap) {
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (rb.d_voices) cudaFree(rb.d_voices);
        if (rb.d_segs) cudaFree(rb.d_segs);
        if (rb.d_nsegs) cudaFree(rb.d_nsegs);
        rb.d_voices = nullptr; rb.d_segs = nullptr; rb.d_nsegs = nullptr;
        rb.voices_cap = 0;
        BLAST_CUDA_TRY(cudaMalloc(&rb.d_voices, nv * sizeof(VoiceDev)));
        rb.seg_cap = std::max(rb.seg_cap, seg_cap);
        BLAST_CUDA_TRY(cudaMalloc(&rb.d_segs, nv * rb.seg_cap * sizeof(Seg)));
        BLAST_CUDA_TRY(cudaMalloc(&rb.d_nsegs, nv * sizeof(uint32_t)));
        rb.voices_cap = std::max(nv, rb.voices_cap);
        if (rb.d_events) {                       // sized by voices: regrown below
            cudaFree(rb.d_events); cudaFree(rb.d_nevents); cudaFree(rb.d_splits); cudaFree(rb.d_nsplits);
            rb.d_events = nullptr; rb.d_nevents = nullptr; rb.d_splits = nullptr; rb.d_nsplits = nullptr;
        }
    }
    if (!rb.d_err) {
        BLAST_CUDA_TRY(cudaMalloc(&rb.d_err, 4 * sizeof(uint32_t)));                       // see RenderBuffers::d_err
        BLAST_CUDA_TRY(cudaMemsetAsync(rb.d_err, 0, 4 * sizeof(uint32_t), ctx->stream));
    }
    if (n_seqs > 0) {
        if (n_seqs > rb.seqs_cap) {
            BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (rb.d_seqs) cudaFree(rb.d_seqs);
            rb.d_seqs = nullptr;
            rb.seqs_cap = 0;
            BLAST_CUDA_TRY(cudaMalloc(&rb.d_seqs, (size_t)n_seqs * sizeof(SeqDev)));
            rb.seqs_cap = n_seqs;
        }
        if (!rb.d_events) {
            BLAST_CUDA_TRY(cudaMalloc(&rb.d_events, rb.voices_cap * kMaxEvents * sizeof(uint32_t)));
            BLAST_CUDA_TRY(cudaMalloc(&rb.d_nevents, rb.voices_cap * sizeof(uint32_t)));
            BLAST_CUDA_TRY(cudaMalloc(&rb.d_splits, rb.voices_cap * kMaxEvents * sizeof(Split)));
            BLAST_CUDA_TRY(cudaMalloc(&rb.d_nsplits, rb.voices_cap * sizeof(uint32_t)));
        }
    }
    return BLAST_OK;
}

// CUDA loads a kernel's code at its first launch (lazy module loading), and that load can wait for the device to go idle.
// A rank whose kernel waits for a peer must therefore never be the reason a peer's FIRST launch of some kernel cannot
// load — which is exactly what happens when group members share a GPU (measured: a hang until the wait's time-out, only
// in the first step; none with CUDA_MODULE_LOADING=EAGER).  Everything the render and the exchange launch is loaded here,
// once per device, before the first step of a peer bus.
int preload_kernels(blast_ctx* ctx) {
    static bool done[64] = {};
    if (ctx->device >= 0 && ctx->device < 64 && done[ctx->device]) return BLAST_OK;
    cudaFuncAttributes a;
#define BLAST_TOUCH(k) BLAST_CUDA_TRY(cudaFuncGetAttributes(&a, k))
    BLAST_TOUCH(seq_event_scan);
    BLAST_TOUCH(voice_position_scan);
    BLAST_TOUCH(voice_split_fixup);
    BLAST_TOUCH(bus_finalize);
    BLAST_TOUCH(peer_publish_tiles);
    BLAST_TOUCH(bus_reduce_tiles);
    BLAST_TOUCH(peer_raise);
    BLAST_TOUCH(flag_wait);
    BLAST_TOUCH((voice_render_mix_tma<1, false, false>));
    BLAST_TOUCH((voice_render_mix_tma<1, true, false>));
    BLAST_TOUCH((voice_render_mix_tma<2, false, false>));
    BLAST_TOUCH((voice_render_mix_tma<2, false, true>));
    BLAST_TOUCH((voice_render_mix_tma<2, true, false>));
    BLAST_TOUCH((voice_render_mix_tma<2, true, true>));
    BLAST_TOUCH(voice_render_mix<1>);
    BLAST_TOUCH(voice_render_mix<2>);
    BLAST_TOUCH(voice_render_mix<3>);
    BLAST_TOUCH(voice_render_mix<4>);
    BLAST_TOUCH(voice_render_mix<5>);
    BLAST_TOUCH(voice_render_mix<6>);
    BLAST_TOUCH(voice_render_mix<7>);
    BLAST_TOUCH(voice_render_mix<8>);
#undef BLAST_TOUCH
    if (ctx->device >= 0 && ctx->device < 64) done[ctx->device] = true;
    return BLAST_OK;
}

// Everything a render of `frames` frames allocates (grow-only): the per-(tile, voice) records and, when enabled, the
// piece-table pool.  launch_render calls it; hosts that must not allocate while a kernel of theirs waits for a peer (group
// members that share a GPU) call it up front.
int reserve_frames(blast_ctx* ctx, RenderBuffers& rb, uint32_t n_voices, uint32_t oc, uint64_t frames) {
    if (n_voices == 0 || frames == 0) return BLAST_OK;
    const size_t need = (size_t)((frames + kFT - 1) / kFT) * n_voices;
    if (need > rb.recs_cap) {
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (rb.d_recs) BLAST_CUDA_TRY(cudaFree(rb.d_recs));
        rb.d_recs = nullptr;
        rb.recs_cap = 0;
        BLAST_CUDA_TRY(cudaMalloc(&rb.d_recs, need * sizeof(TileRec)));
        rb.recs_cap = need;
    }
    static const bool tables = getenv("BLAST_RENDER_TABLES") != nullptr;
    static const bool legacy = getenv("BLAST_RENDER_LEGACY") != nullptr;
    if (oc == 2 && !legacy && tables) {
        const uint32_t rows = pool_rows_for(rb.seg_cap);
        if (rb.pool_voices < rb.voices_cap || rb.pool_rows != rows) {
            BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (rb.d_pool) cudaFree(rb.d_pool);
            rb.d_pool = nullptr;
            rb.pool_voices = 0;
            BLAST_CUDA_TRY(cudaMalloc(&rb.d_pool, rb.voices_cap * (size_t)rows * sizeof(uint4)));
            rb.pool_voices = rb.voices_cap;
            rb.pool_rows = rows;
        }
    }
    return BLAST_OK;
}

int launch_flag_wait(blast_ctx* ctx, const uint32_t* d_flags, uint32_t n, uint32_t value, uint32_t timeout_ms, uint32_t* d_err,
                     cudaStream_t stream) {
    if (n == 0) return BLAST_OK;
    if (n > 32) return blast::set_error(BLAST_ERR_CAPACITY, "at most 32 flags per wait");
    if (!stream) stream = ctx->stream;
    flag_wait<<<1, 32, 0, stream>>>(d_flags, n, value, timeout_ms, d_err);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

static int launch_peer_raise(blast_ctx* ctx, const BusSink& sink, cudaStream_t stream) {
    peer_raise<<<1, 2 * kMaxPeers, 0, stream>>>(sink);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

int launch_bus_reduce(blast_ctx* ctx, const BusSink& sink, cudaStream_t stream) {
    if (!stream) stream = ctx->stream;
    if (sink.n_tiles) {
        peer_publish_tiles<<<(sink.n_tiles + 255) / 256, 256, 0, stream>>>(sink);
        BLAST_CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
    }
    if (sink.n_my_tiles == 0) return launch_peer_raise(ctx, sink, stream);
    bus_reduce_tiles<<<sink.n_my_tiles, 256, 0, stream>>>(sink);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

int launch_render(blast_ctx* ctx, RenderBuffers& rb, uint32_t n_voices, uint32_t n_seqs, uint32_t oc, uint64_t frames,
                  int32_t* d_partial_bus, const BusSink* sink_in, const VoiceDev* rewind) {
    if (frames == 0) return (sink_in && sink_in->world) ? launch_bus_reduce(ctx, *sink_in) : BLAST_OK;
    if (frames > 0x7FFFFFFFull) return blast::set_error(BLAST_ERR_CAPACITY, "at most 2^31-1 frames per render call");
    if (n_seqs > 0 && frames * oc > 0x7FFFFFFFull)
        return blast::set_error(BLAST_ERR_CAPACITY, "at most 2^31-1 calls (frames x channels) per render call with Seq processes");
    const uint32_t n_tiles = (uint32_t)((frames + kFT - 1) / kFT);
    const size_t slots = (size_t)frames * oc;
    static const bool legacy = getenv("BLAST_RENDER_LEGACY") != nullptr;
    const bool fused = sink_in && sink_in->world && oc <= 2 && !legacy && n_voices > 0 && n_seqs == 0;   // the reduction rides in K4's work queue
    // (K4b patches the partial bus after K4, so voices with Seq processes take the two-kernel reduction)
    if (sink_in && sink_in->world && !fused && sink_in->world > 1)                         // nobody reads the previous partial bus any more
        if (int rc = launch_flag_wait(ctx, sink_in->ack_mine, sink_in->world, sink_in->step - 1u, sink_in->timeout_ms, sink_in->err)) return rc;
    if (n_voices == 0) {
        BLAST_CUDA_TRY(cudaMemsetAsync(d_partial_bus, 0, slots * sizeof(int32_t), ctx->stream));
        BLAST_CUDA_TRY(cudaMemsetAsync(rb.d_err, 0, 4 * sizeof(uint32_t), ctx->stream));   // nothing can overflow
        if (sink_in && sink_in->world) return launch_bus_reduce(ctx, *sink_in);
        return BLAST_OK;
    }
    if (int rc = reserve_frames(ctx, rb, n_voices, oc, frames)) return rc;
    // voice groups: enough work items to fill the GPU a few times over
    uint32_t groups = 1;
    static const uint32_t ctas_per_sm = getenv("BLAST_RENDER_CTAS_PER_SM") ? (uint32_t)atoi(getenv("BLAST_RENDER_CTAS_PER_SM")) : 32u;
    const uint32_t want_ctas = (uint32_t)ctx->sm_count * ctas_per_sm;
    // groups of >= 16 voices: a sharded scene (1,024 files over 8 GPUs = 128 voices per rank) still gets ~6 work items per
    // CTA for the queue to balance; smaller groups pay more per flush than they gain (profiles/r02_group_sweep.json)
    static const uint32_t min_group = getenv("BLAST_RENDER_MIN_GROUP") ? (uint32_t)atoi(getenv("BLAST_RENDER_MIN_GROUP")) : 16u;
    while (n_tiles * groups < want_ctas && n_voices / (groups * 2) >= min_group) groups *= 2;
    // persistent: CTAs per SM (three fit: shared-memory bound) take (tile, group) items, tile-major, from a counter
    const uint32_t resident = (uint32_t)ctx->sm_count * (uint32_t)ctx->render_ctas_per_sm;
    static const uint32_t forced_groups = getenv("BLAST_RENDER_GROUPS") ? (uint32_t)atoi(getenv("BLAST_RENDER_GROUPS")) : 0u;   // development
    if (forced_groups) groups = std::max(1u, std::min(forced_groups, n_voices));
    const uint32_t per_group = (n_voices + groups - 1) / groups;
    groups = (n_voices + per_group - 1) / per_group;
    const int use_atomic = groups > 1;
    BusSink sink{};
    if (fused) {
        sink = *sink_in;
        // tile t is reduced once the queue is `lag` tiles further: by then its last voice group has normally been flushed
        // on every rank (1.5 x the CTAs in flight, plus slack for the skew between ranks), so the reduction seldom waits
        sink.lag = std::min<uint32_t>(n_tiles, (3 * resident / 2 + groups - 1) / groups + 2);
        static const int forced_lag = getenv("BLAST_SINK_LAG") ? atoi(getenv("BLAST_SINK_LAG")) : 0;      // development
        if (forced_lag > 0) sink.lag = std::min<uint32_t>(n_tiles, (uint32_t)forced_lag);
        if (sink.lag < 1) sink.lag = 1;
    }

    // piece-table pool of the tiles with several segments (stereo bus, TMA kernel): grow-only, sized by voices x segments
    static const bool tables = getenv("BLAST_RENDER_TABLES") != nullptr;
    // development switch: bit 0 = one-segment groups of a multi-segment tile as ordinary pieces, bit 1 = two / three segments as
    // one piece (cooperative path), bit 2 = the same cut by the voice's own lane
    static const uint32_t opts = getenv("BLAST_RENDER_OPTS") ? (uint32_t)atoi(getenv("BLAST_RENDER_OPTS")) : 7u;
    uint4* tab_pool = (oc == 2 && !legacy && tables) ? rb.d_pool : nullptr;
    rb.parity ^= 1u;                                         // this render's error word; cleared by the previous render's K3
    uint32_t* d_err = rb.d_err + rb.parity;
    uint32_t* d_work = rb.d_err + 2;
    if (n_seqs > 0) {
        seq_event_scan<<<(n_voices + kSeqScanThreads / 32 - 1) / (kSeqScanThreads / 32), kSeqScanThreads, 0, ctx->stream>>>(
            rb.d_voices, n_voices, rb.d_seqs, (uint32_t)(frames * oc), rb.d_events, rb.d_nevents, d_err);
        BLAST_CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
    }
    {
        // K3's extra blocks clear the NEXT render's error word, K4's work counter and (for atomics) the partial bus
        const uint32_t n_voice_blocks = (n_voices + kScanThreads / 32 - 1) / (kScanThreads / 32);
        const size_t n_zero = use_atomic ? slots : 0;
        const uint32_t n_zero_blocks = (uint32_t)std::max<size_t>(1, (n_zero + kZeroPerBlock - 1) / kZeroPerBlock);
        voice_position_scan<<<n_voice_blocks + n_zero_blocks, kScanThreads, 0, ctx->stream>>>(
            rb.d_voices, n_voices, (uint32_t)frames, rb.d_segs, rb.d_nsegs, d_err, rb.d_events, rb.d_nevents, rb.seg_cap, oc,
            rb.d_splits, rb.d_nsplits, rb.d_recs, n_tiles, n_voice_blocks, rb.d_err + (rb.parity ^ 1u), d_work,
            reinterpret_cast<uint32_t*>(d_partial_bus), n_zero, sink, tab_pool, rb.pool_rows, (uint32_t)kStageBytes, rewind);
        BLAST_CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
    }
    dim3 grid(n_tiles, groups);
    dim3 grid_tma(std::min<uint32_t>(n_tiles * groups, resident), 1);
    if (oc <= 2 && !legacy) {
        // TMA pipeline kernel (one producer warp + eight consumer warps)
#define BLAST_LAUNCH_TMA(OCV, SINK, TAB)                                                                                              \
    do {                                                                                                                                \
        BLAST_CUDA_TRY(cudaFuncSetAttribute(voice_render_mix_tma<OCV, SINK, TAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmem)); \
        voice_render_mix_tma<OCV, SINK, TAB><<<grid_tma, kTmaThreads, kTmaSmem, ctx->stream>>>(                                      \
            rb.d_voices, n_voices, per_group, groups, rb.d_segs, rb.d_nsegs, rb.d_recs, (uint32_t)frames, d_partial_bus, use_atomic, \
            d_err, rb.seg_cap, d_work, sink, -0.0f, tab_pool, rb.pool_rows, opts);                                                       \
    } while (0)
        const bool with_tab = tab_pool != nullptr;
        if (oc == 1) {
            if (fused) BLAST_LAUNCH_TMA(1, true, false); else BLAST_LAUNCH_TMA(1, false, false);
        } else if (fused) {
            if (with_tab) BLAST_LAUNCH_TMA(2, true, true); else BLAST_LAUNCH_TMA(2, true, false);
        } else {
            if (with_tab) BLAST_LAUNCH_TMA(2, false, true); else BLAST_LAUNCH_TMA(2, false, false);
        }
#undef BLAST_LAUNCH_TMA
    } else {
#define BLAST_LAUNCH_MIX(OCV)                                                                              \
    voice_render_mix<OCV><<<grid, kThreads, 0, ctx->stream>>>(rb.d_voices, n_voices, per_group, rb.d_segs,  \
                                                               rb.d_nsegs, rb.d_recs, (uint32_t)frames,      \
                                                               d_partial_bus, use_atomic, d_err, rb.seg_cap)
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
    }
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    if (n_seqs > 0) {
        voice_split_fixup<<<(n_voices * (uint32_t)kMaxEvents + 255) / 256, 256, 0, ctx->stream>>>(rb.d_voices, n_voices, rb.d_splits,
                                                                                              rb.d_nsplits, oc, d_partial_bus, d_err);
        BLAST_CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
    }
    if (sink_in && sink_in->world && !fused) return launch_bus_reduce(ctx, *sink_in);
    if (fused && sink.n_my_tiles == 0) return launch_peer_raise(ctx, sink, ctx->stream);
    return BLAST_OK;
}

}  // namespace blast_rdr

struct blast_scene {
    uint32_t n_voices = 0;
    uint32_t out_channels = 0;
    std::vector<blast_track> tracks;
    std::vector<blast_voice> voices;     // host mirror of the ABI voices (positions refreshed on get)
    RenderBuffers rb;
    VoiceDev* d_voices0 = nullptr;       // the table as last uploaded (blast_scene_restore_dev)
    bool rewind = false;                 // the next render starts from d_voices0 (its position scan copies the voice rows)
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
    route_voice(v, sc->out_channels, false);
    *out = v;
    return BLAST_OK;
}

int upload_voices(blast_ctx* ctx, blast_scene* sc) {
    std::vector<VoiceDev> hv(sc->n_voices);
    for (uint32_t i = 0; i < sc->n_voices; ++i)
        if (int rc = make_voice_dev(sc, sc->voices[i], i, &hv[i])) return rc;
    if (sc->n_voices) {
        BLAST_CUDA_TRY(cudaMemcpyAsync(sc->rb.d_voices, hv.data(), hv.size() * sizeof(VoiceDev), cudaMemcpyHostToDevice, ctx->stream));
        BLAST_CUDA_TRY(cudaMemcpyAsync(sc->d_voices0, sc->rb.d_voices, hv.size() * sizeof(VoiceDev), cudaMemcpyDeviceToDevice, ctx->stream));
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    sc->rewind = false;
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
    if (int rc = reserve_buffers(ctx, sc->rb, n_voices, 0)) return fail(rc);
    if (cudaMalloc(&sc->d_voices0, nv * sizeof(VoiceDev)) != cudaSuccess)
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
    free_buffers(sc->rb);
    if (sc->d_voices0) cudaFree(sc->d_voices0);
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

int blast_scene_restore_dev(blast_ctx* ctx, blast_scene* sc) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc != nullptr, BLAST_ERR_ARG, "blast_scene_restore_dev: null scene");
    sc->rewind = true;                   // no launch: the next render's position scan reads the uploaded rows
    return BLAST_OK;
}

int blast_scene_get_voices(blast_ctx* ctx, blast_scene* sc, blast_voice* out, uint32_t n_voices) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc && out, BLAST_ERR_ARG, "blast_scene_get_voices: null argument");
    BLAST_REQUIRE(n_voices == sc->n_voices, BLAST_ERR_ARG, "blast_scene_get_voices: voice count differs from the scene's");
    std::vector<VoiceDev> hv(sc->n_voices);
    if (sc->n_voices) {
        BLAST_CUDA_TRY(cudaMemcpyAsync(hv.data(), sc->rewind ? sc->d_voices0 : sc->rb.d_voices, hv.size() * sizeof(VoiceDev), cudaMemcpyDeviceToHost, ctx->stream));
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
    const VoiceDev* rewind = (sc->rewind && frames) ? sc->d_voices0 : nullptr;
    if (frames) sc->rewind = false;
    return launch_render(ctx, sc->rb, sc->n_voices, 0, sc->out_channels, frames, d_partial_bus, nullptr, rewind);
}

int blast_scene_reserve(blast_ctx* ctx, blast_scene* sc, uint64_t frames) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc != nullptr, BLAST_ERR_ARG, "blast_scene_reserve: null scene");
    return reserve_frames(ctx, sc->rb, sc->n_voices, sc->out_channels, frames);
}

int blast_scene_check(blast_ctx* ctx, blast_scene* sc) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc != nullptr, BLAST_ERR_ARG, "blast_scene_check: null scene");
    uint32_t e = 0;
    BLAST_CUDA_TRY(cudaMemcpyAsync(&e, sc->rb.err_word(), sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
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

int blast_scene_render_reduce_dev(blast_ctx* ctx, blast_scene* sc, uint64_t frames, blast_peer_bus* pb) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(sc && pb, BLAST_ERR_ARG, "blast_scene_render_reduce_dev: null argument");
    BusSink sink;
    if (int rc = peer_bus_next_step(ctx, pb, frames, sc->out_channels, true, &sink)) return rc;
    const VoiceDev* rewind = (sc->rewind && frames) ? sc->d_voices0 : nullptr;
    if (frames) sc->rewind = false;
    if (!peer_bus_fused(pb)) {
        // Default: render, then the exchange as its own launches — one finalize on a single rank; tile publish + tile
        // reduce over peer memory (after a wait for the peers' acknowledgements of the previous step) on several.  The
        // render kernel can do the exchange itself (blast_peer_bus_set_fused), tile by tile as the tiles complete; measured
        // on B200s that variant is the slower one: its producer warp carries the hand-off and spills
        // (profiles/r02_sink_lag_sweep.json, r02_bench_n2_*.json).
        if (sink.world > 1)
            if (int rc = launch_flag_wait(ctx, sink.ack_mine, sink.world, sink.step - 1u, sink.timeout_ms, sink.err)) return rc;
        if (int rc = launch_render(ctx, sc->rb, sc->n_voices, 0, sc->out_channels, frames, peer_bus_partial(pb), nullptr, rewind)) return rc;
        if (sink.world == 1) return blast_bus_finalize_dev(ctx, peer_bus_partial(pb), blast_peer_bus_bus(pb), frames * sc->out_channels);
        return peer_bus_exchange(ctx, pb, sink);
    }
    return launch_render(ctx, sc->rb, sc->n_voices, 0, sc->out_channels, frames, peer_bus_partial(pb), &sink, rewind);
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
