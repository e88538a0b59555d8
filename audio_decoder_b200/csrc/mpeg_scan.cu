// mpeg_scan.cu — K7 `mpeg_sync_scan` and K8 `mpeg_header_classify`.
//
// Replaces blast/src/file_parsing/mpeg.rs:
//   :17-50    greedy, NON-overlapping sync scan: at cur, if b[cur]==0xFF && (b[cur+1]&0xE0)==0xE0 the
//             4-byte header is recorded and cur += 4, else cur += 1 (a truncated trailing header is dropped)
//   :53-73    reference header = most frequent header VALUE that parse_header accepts
//   :77-127   a candidate is a frame iff its header parses, match_ref()s and has a valid frame length;
//             frames are ordered by file position; the first position of every distinct header value
//             appears twice (`or_insert(vec![fp]).push(fp)`, :39)
//   :367-496, :154-234, :255-303  parse_header / Header::{format,match_ref,compute_frame_len} with their
//             quirks (version low bit = protection bit, bitrate column always 4, "CRC" = 20)
//
// K7 is a single-pass chained scan, HBM-bound at 1 byte read per input byte (+12 B per candidate).
// The greedy rule is a 4-state machine (state = header bytes still to skip); a byte range acts on it as
// a map {0..3} -> {0..3} plus a candidate count per entry state, and maps compose associatively.  Each
// thread owns 64 consecutive bytes: it builds the 64-bit "raw sync" mask with SWAR byte tests, resolves
// the greedy selection for the entry states that can differ, and summarises itself as (map, counts).
// Lanes whose predecessor's map is constant know their entry state at once (the overwhelmingly common
// case); the rest propagate in a short loop.  Warp summaries are combined by warp 0, tile summaries by
// decoupled look-back over 32-tile windows, and every thread then emits its candidates at their final,
// position-ordered indices.
#include <algorithm>
#include <vector>

#include "blast_internal.h"

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kChunk = 64;                                  // bytes per thread per round
constexpr int kRounds = 16;                                 // rounds per tile: a warp walks 16 x 2 KiB = 32 KiB
constexpr int kSpanBytes = kRounds * 32 * kChunk;           // contiguous bytes per warp per tile
constexpr int kTileBytes = kScanWarps * kSpanBytes;         // 256 KiB
constexpr uint32_t kIdentityMap = 0xE4;                     // s -> s for s = 0..3, two bits each

struct TileDesc {             // 64 bytes
    uint32_t flag;            // 0 = nothing, 1 = aggregate valid, 2 = inclusive prefix valid
    uint32_t map;             // aggregate: exit state per entry state (4 x 2 bits)
    uint32_t c[4];            // aggregate: candidates per entry state
    uint32_t state;           // inclusive: exit state of this tile under the true entry state
    uint32_t pad0;
    uint32_t cnt_lo, cnt_hi;  // inclusive: candidates in tiles 0..this
    uint32_t pad[6];
};

struct ScanCtl {
    unsigned long long next_tile;
    unsigned long long total;     // candidates found
    uint32_t panic;               // the reference indexes out of bounds (last byte 0xFF reached with state 0)
    uint32_t pad;
};

// ---- SWAR byte predicates: 0x80 in every byte that satisfies the test
__device__ __forceinline__ uint32_t is_ff(uint32_t w) {
    const uint32_t y = (~w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~y & w & 0x80808080u;
}
__device__ __forceinline__ uint32_t is_e0(uint32_t w) {        // (b & 0xE0) == 0xE0
    const uint32_t t = (w & 0x60606060u) + 0x20202020u;
    return t & w & 0x80808080u;
}
__device__ __forceinline__ uint32_t nibble_of(uint32_t flags) {  // 0x80 flags of 4 bytes -> 4 bits
    return ((flags >> 7) * 0x10204080u) >> 28;
}

// greedy non-overlapping selection on a 64-bit raw mask, skipping the first s_in bytes
__device__ __forceinline__ unsigned long long resolve(unsigned long long M, uint32_t s_in, uint32_t& s_out) {
    unsigned long long m = M & (~0ull << s_in), sel = 0;
    while (m) {
        const int i = __ffsll((long long)m) - 1;
        sel |= 1ull << i;
        m &= ~(0xFull << i);
    }
    if (sel) {
        const int last = 63 - __clzll((long long)sel);
        s_out = last > 60 ? (uint32_t)(last - 60) : 0u;          // bytes of the last header that spill over
    } else {
        s_out = 0;                                               // 64 bytes drain any entry state (<= 3)
    }
    return sel;
}

__device__ __forceinline__ uint32_t map_get(uint32_t map, uint32_t s) { return (map >> (2 * s)) & 3u; }
__device__ __forceinline__ uint32_t map_after(uint32_t first, uint32_t then) {   // s -> then[first[s]]
    uint32_t r = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) r |= map_get(then, map_get(first, s)) << (2 * s);
    return r;
}
__device__ __forceinline__ bool map_is_const(uint32_t map) {
    const uint32_t e = map & 3u;
    return map == e * 0x55u;
}
__device__ __forceinline__ uint32_t cnt16(uint32_t c01, uint32_t c23, uint32_t s) {
    const uint32_t w = (s & 2) ? c23 : c01;
    return (s & 1) ? (w >> 16) : (w & 0xFFFFu);
}

template <typename T>
__device__ __forceinline__ T pick4(T a0, T a1, T a2, T a3, uint32_t i) {       // register-only a[i]
    return i == 0 ? a0 : i == 1 ? a1 : i == 2 ? a2 : a3;
}

struct Agg {
    uint32_t map;
    uint32_t c[4];
};
__device__ __forceinline__ Agg agg_then(const Agg& a, const Agg& b) {            // a first, then b
    Agg r;
    r.map = map_after(a.map, b.map);
#pragma unroll
    for (int s = 0; s < 4; ++s) r.c[s] = a.c[s] + pick4(b.c[0], b.c[1], b.c[2], b.c[3], map_get(a.map, s));
    return r;
}

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kScanThreads)
mpeg_sync_scan(const uint8_t* __restrict__ bytes, unsigned long long n, TileDesc* __restrict__ desc,
               unsigned long long n_tiles, ScanCtl* __restrict__ ctl, unsigned long long* __restrict__ out_pos,
               uint32_t* __restrict__ out_hdr, unsigned long long cap) {
    // per (warp, round, lane): raw-sync mask and the lane's pre-map relative to the warp span's entry state
    __shared__ unsigned long long s_mask[kScanWarps * kRounds * 32];
    __shared__ uint8_t s_tpre[kScanWarps * kRounds * 32];
    __shared__ unsigned long long s_tile;
    __shared__ uint32_t s_wmap[kScanWarps], s_wc[kScanWarps][4];      // warp-span aggregates
    __shared__ uint32_t s_wentry[kScanWarps];                          // true entry state of every warp span
    __shared__ unsigned long long s_wbase[kScanWarps];                 // global candidate index at the span start
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&ctl->next_tile, 1ull);
        __syncthreads();
        const unsigned long long tile = s_tile;
        if (tile >= n_tiles) return;
        const unsigned long long span0 = tile * (unsigned long long)kTileBytes + (unsigned long long)warp * kSpanBytes;

        // ================= pass 1: masks + summaries, one warp walks its contiguous 32 KiB span
        uint32_t wrun = kIdentityMap;          // span entry state -> entry state of the current round (uniform)
        uint32_t wc01 = 0, wc23 = 0;           // candidates so far per span-entry hypothesis (uniform, 4 x 16 bit)
#pragma unroll 1
        for (int r = 0; r < kRounds; ++r) {
            const unsigned long long base = span0 + (unsigned long long)r * (32 * kChunk) + (unsigned long long)lane * kChunk;
            uint32_t w[17];
            {
                const uint4* v = reinterpret_cast<const uint4*>(bytes + base);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 x = make_uint4(0, 0, 0, 0);
                    if (base + 16ull * q < n) x = __ldg(v + q);
                    w[4 * q] = x.x; w[4 * q + 1] = x.y; w[4 * q + 2] = x.z; w[4 * q + 3] = x.w;
                }
                w[16] = (base + 64 < n) ? __ldg(reinterpret_cast<const uint32_t*>(bytes + base + 64)) : 0u;
                if (base + 68 > n) {                                     // the buffer ends inside this window
#pragma unroll
                    for (int k = 0; k < 17; ++k) {
                        const unsigned long long p = base + 4ull * k;
                        if (p >= n) w[k] = 0;
                        else if (p + 4 > n) w[k] &= (1u << (8 * (uint32_t)(n - p))) - 1u;
                    }
                }
            }
            // raw sync mask: bit i <=> b[i]==0xFF && (b[i+1]&0xE0)==0xE0; two words per multiply
            uint32_t mlo, mhi;
            {
                uint32_t raw[16];
                uint32_t e_next = is_e0(w[16]);
#pragma unroll
                for (int k = 15; k >= 0; --k) {
                    const uint32_t e = is_e0(w[k]);
                    raw[k] = is_ff(w[k]) & __funnelshift_r(e, e_next, 8);
                    e_next = e;
                }
                uint32_t b8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) b8[k] = (((raw[2 * k] >> 7) | (raw[2 * k + 1] >> 3)) * 0x01020408u);   // byte 3 = 8 flags
                mlo = __byte_perm(__byte_perm(b8[0], b8[1], 0x0073), __byte_perm(b8[2], b8[3], 0x0073), 0x5410);
                mhi = __byte_perm(__byte_perm(b8[4], b8[5], 0x0073), __byte_perm(b8[6], b8[7], 0x0073), 0x5410);
            }
            const unsigned long long M = ((unsigned long long)mhi << 32) | mlo;
            unsigned long long okmask = ~0ull;
            if (base + 67 >= n) {
                const unsigned long long lim = n > base + 3 ? n - 3 - base : 0;
                okmask = lim >= 64 ? ~0ull : ((1ull << lim) - 1ull);
            }
            // per-thread summary
            uint32_t my_map, my_cnt;
            {
                uint32_t e0;
                const unsigned long long sel0 = resolve(M, 0, e0);
                my_map = e0 * 0x55u;
                my_cnt = (uint32_t)__popcll(sel0 & okmask) * 0x01010101u;
                if (M & 7ull) {
#pragma unroll
                    for (uint32_t s = 1; s < 4; ++s) {
                        uint32_t es;
                        const unsigned long long sel = resolve(M, s, es);
                        my_map = (my_map & ~(3u << (2 * s))) | (es << (2 * s));
                        my_cnt = (my_cnt & ~(0xFFu << (8 * s))) | ((uint32_t)__popcll(sel & okmask) << (8 * s));
                    }
                }
            }
            // lane's pre-map within this round: round entry state -> lane entry state
            uint32_t pre = kIdentityMap;
            {
                bool known = lane == 0;
                const uint32_t pm = __shfl_up_sync(0xFFFFFFFFu, my_map, 1);
                if (lane > 0 && map_is_const(pm)) { pre = pm; known = true; }
                while (!__all_sync(0xFFFFFFFFu, known)) {
                    const uint32_t ppre = __shfl_up_sync(0xFFFFFFFFu, pre, 1);
                    const bool pknown = __shfl_up_sync(0xFFFFFFFFu, (int)known, 1) != 0;
                    if (!known && pknown) { pre = map_after(ppre, pm); known = true; }
                }
            }
            const uint32_t tpre = map_after(wrun, pre);              // span entry -> lane entry
            const uint32_t slot = (warp * kRounds + r) * 32 + lane;
            s_mask[slot] = M;
            s_tpre[slot] = (uint8_t)tpre;
            {
                const uint32_t k0 = (my_cnt >> (8 * map_get(tpre, 0))) & 0xFF, k1 = (my_cnt >> (8 * map_get(tpre, 1))) & 0xFF;
                const uint32_t k2 = (my_cnt >> (8 * map_get(tpre, 2))) & 0xFF, k3 = (my_cnt >> (8 * map_get(tpre, 3))) & 0xFF;
                wc01 += __reduce_add_sync(0xFFFFFFFFu, k0 | (k1 << 16));
                wc23 += __reduce_add_sync(0xFFFFFFFFu, k2 | (k3 << 16));
            }
            const uint32_t round_map = __shfl_sync(0xFFFFFFFFu, map_after(pre, my_map), 31);
            wrun = map_after(wrun, round_map);
        }
        if (lane == 0) {
            s_wmap[warp] = wrun;
            s_wc[warp][0] = wc01 & 0xFFFF; s_wc[warp][1] = wc01 >> 16; s_wc[warp][2] = wc23 & 0xFFFF; s_wc[warp][3] = wc23 >> 16;
        }
        __syncthreads();

        // ================= warp 0: combine the spans, publish the aggregate, look back, publish inclusive
        if (warp == 0) {
            Agg run;
            run.map = kIdentityMap;
            run.c[0] = run.c[1] = run.c[2] = run.c[3] = 0;
            Agg keep = run;                        // prefix before span `lane`
#pragma unroll
            for (int k = 0; k < kScanWarps; ++k) {
                if ((int)lane == k) keep = run;
                Agg wk;
                wk.map = s_wmap[k];
                wk.c[0] = s_wc[k][0]; wk.c[1] = s_wc[k][1]; wk.c[2] = s_wc[k][2]; wk.c[3] = s_wc[k][3];
                run = agg_then(run, wk);
            }
            TileDesc* me = desc + tile;
            uint32_t entry = 0;
            unsigned long long cbase = 0;
            if (tile > 0) {
                if (lane == 0) {
                    me->map = run.map;
                    me->c[0] = run.c[0]; me->c[1] = run.c[1]; me->c[2] = run.c[2]; me->c[3] = run.c[3];
                    st_release(&me->flag, 1u);
                }
                uint32_t fmap = kIdentityMap;
                unsigned long long fc[4] = {0, 0, 0, 0};
                long long look = (long long)tile - 1;
                for (;;) {
                    const long long t = look - (long long)lane;
                    uint32_t flag = 2, amap = kIdentityMap, a0 = 0, a1 = 0, a2 = 0, a3 = 0, st = 0, clo = 0, chi = 0;
                    if (t >= 0) {
                        const TileDesc* dsc = desc + t;
                        do { flag = ld_acquire(&dsc->flag); } while (flag == 0);
                        if (flag == 2) { st = ld_relaxed(&dsc->state); clo = ld_relaxed(&dsc->cnt_lo); chi = ld_relaxed(&dsc->cnt_hi); }
                        else { amap = ld_relaxed(&dsc->map); a0 = ld_relaxed(&dsc->c[0]); a1 = ld_relaxed(&dsc->c[1]); a2 = ld_relaxed(&dsc->c[2]); a3 = ld_relaxed(&dsc->c[3]); }
                    }
                    const uint32_t incm = __ballot_sync(0xFFFFFFFFu, flag == 2);     // lanes past tile 0 count as inclusive(0,0)
                    const int first = incm ? __ffs(incm) - 1 : 32;
                    // "simple" aggregate: constant map, count independent of the entry state (practically every tile)
                    const bool simple = map_is_const(amap) && a0 == a1 && a1 == a2 && a2 == a3;
                    const uint32_t need = first >= 32 ? 0xFFFFFFFFu : ((1u << first) - 1u);
                    const uint32_t simple_m = __ballot_sync(0xFFFFFFFFu, simple) & need;
                    if (simple_m == need) {
                        if (first > 0) {
                            const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, (int)lane < first ? a0 : 0u);
                            const uint32_t e0 = __shfl_sync(0xFFFFFFFFu, amap, 0) & 3u;
                            const unsigned long long nc = tot + pick4(fc[0], fc[1], fc[2], fc[3], e0);
                            fc[0] = fc[1] = fc[2] = fc[3] = nc;
                            fmap = map_get(fmap, e0) * 0x55u;
                        }
                    } else {
                        for (int i = 0; i < first; ++i) {                              // nearest tile first
                            const uint32_t m_i = __shfl_sync(0xFFFFFFFFu, amap, i);
                            const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, a0, i), b1 = __shfl_sync(0xFFFFFFFFu, a1, i);
                            const uint32_t b2 = __shfl_sync(0xFFFFFFFFu, a2, i), b3 = __shfl_sync(0xFFFFFFFFu, a3, i);
                            unsigned long long nc[4];
#pragma unroll
                            for (int s = 0; s < 4; ++s) nc[s] = pick4(b0, b1, b2, b3, (uint32_t)s) + pick4(fc[0], fc[1], fc[2], fc[3], map_get(m_i, s));
#pragma unroll
                            for (int s = 0; s < 4; ++s) fc[s] = nc[s];
                            fmap = map_after(m_i, fmap);
                        }
                    }
                    if (first < 32) {
                        const uint32_t sigma = __shfl_sync(0xFFFFFFFFu, st, first);
                        const uint32_t lo = __shfl_sync(0xFFFFFFFFu, clo, first), hi = __shfl_sync(0xFFFFFFFFu, chi, first);
                        entry = map_get(fmap, sigma);
                        cbase = (((unsigned long long)hi << 32) | lo) + pick4(fc[0], fc[1], fc[2], fc[3], sigma);
                        break;
                    }
                    look -= 32;
                }
            }
            if (lane == 0) {
                const unsigned long long cend = cbase + pick4(run.c[0], run.c[1], run.c[2], run.c[3], entry);
                me->state = map_get(run.map, entry);
                me->cnt_lo = (uint32_t)cend;
                me->cnt_hi = (uint32_t)(cend >> 32);
                st_release(&me->flag, 2u);
                if (tile + 1 == n_tiles) ctl->total = cend;
            }
            if (lane < kScanWarps) {
                s_wentry[lane] = map_get(keep.map, entry);
                s_wbase[lane] = cbase + pick4(keep.c[0], keep.c[1], keep.c[2], keep.c[3], entry);
            }
        }
        __syncthreads();

        // ================= pass 2: true entry states are known; emit candidates at their final indices
        {
            const uint32_t span_entry = s_wentry[warp];
            unsigned long long run_base = s_wbase[warp];
            const bool near_end = span0 + kSpanBytes + 4 > n;       // this span touches the end of the buffer
#pragma unroll 1
            for (int r = 0; r < kRounds; ++r) {
                const unsigned long long base = span0 + (unsigned long long)r * (32 * kChunk) + (unsigned long long)lane * kChunk;
                const uint32_t slot = (warp * kRounds + r) * 32 + lane;
                const unsigned long long M = s_mask[slot];
                const uint32_t my_entry = map_get((uint32_t)s_tpre[slot], span_entry);
                uint32_t dummy;
                const unsigned long long sel_all = M ? resolve(M, my_entry, dummy) : 0ull;
                unsigned long long okmask = ~0ull;
                if (near_end && base + 67 >= n) {
                    const unsigned long long lim = n > base + 3 ? n - 3 - base : 0;
                    okmask = lim >= 64 ? ~0ull : ((1ull << lim) - 1ull);
                }
                unsigned long long sel = sel_all & okmask;
                const uint32_t cnt = (uint32_t)__popcll(sel);
                uint32_t inc = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if ((int)lane >= d) inc += a;
                }
                unsigned long long idx = run_base + (inc - cnt);
                run_base += __shfl_sync(0xFFFFFFFFu, inc, 31);
                // header = 4 bytes at base+i, read as two aligned words (L2 hits: the tile was just scanned).
                // The first two candidates' loads are issued together; more than two per 64 bytes is rare.
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(bytes + base);
                int i0 = -1, i1 = -1;
                uint32_t a0 = 0, b0 = 0, a1 = 0, b1 = 0;
                if (sel) { i0 = __ffsll((long long)sel) - 1; sel &= sel - 1; }
                if (sel) { i1 = __ffsll((long long)sel) - 1; sel &= sel - 1; }
                if (i0 >= 0) { a0 = __ldg(wp + (i0 >> 2)); if (i0 & 3) b0 = __ldg(wp + (i0 >> 2) + 1); }
                if (i1 >= 0) { a1 = __ldg(wp + (i1 >> 2)); if (i1 & 3) b1 = __ldg(wp + (i1 >> 2) + 1); }
                if (i0 >= 0) {
                    if (idx < cap) {
                        out_pos[idx] = base + (unsigned long long)i0;
                        out_hdr[idx] = __byte_perm(__funnelshift_r(a0, b0, 8 * (i0 & 3)), 0, 0x0123);    // big-endian (mpeg.rs:22-38)
                    }
                    idx += 1;
                }
                if (i1 >= 0) {
                    if (idx < cap) {
                        out_pos[idx] = base + (unsigned long long)i1;
                        out_hdr[idx] = __byte_perm(__funnelshift_r(a1, b1, 8 * (i1 & 3)), 0, 0x0123);
                    }
                    idx += 1;
                }
                while (sel) {
                    const int i = __ffsll((long long)sel) - 1;
                    sel &= sel - 1;
                    if (idx < cap) {
                        const uint32_t a = __ldg(wp + (i >> 2)), b = (i & 3) ? __ldg(wp + (i >> 2) + 1) : 0u;
                        out_pos[idx] = base + (unsigned long long)i;
                        out_hdr[idx] = __byte_perm(__funnelshift_r(a, b, 8 * (i & 3)), 0, 0x0123);
                    }
                    idx += 1;
                }
                // mpeg.rs:20: `reader[cur + 1]` with cur == n-1 panics when the scan reaches a trailing 0xFF
                if (near_end && n > 0 && n - 1 >= base && n - 1 < base + kChunk) {
                    const uint32_t il = (uint32_t)(n - 1 - base);
                    if (bytes[n - 1] == 0xFF) {
                        const unsigned long long before = il ? (sel_all & ((1ull << il) - 1ull)) : 0ull;
                        bool skipped = il < my_entry;
                        if (before) skipped = skipped || (il - (uint32_t)(63 - __clzll((long long)before)) <= 3);
                        if (!skipped) ctl->panic = 1;
                    }
                }
            }
        }
        __syncthreads();            // shared scratch is reused by the next tile
    }
}

// ================================================================ K8
// parse_header (mpeg.rs:367-496) as integer logic.  Returns false on Err.
struct HdrInfo {
    uint32_t version, layer, not_prot, ff, chmode, eeee, padded;
};
__host__ __device__ inline bool parse_header_bits(uint32_t h, HdrInfo& o) {
    const uint32_t b1 = (h >> 16) & 0xFF, b2 = (h >> 8) & 0xFF, b3 = h & 0xFF;
    o.version = (((b1 >> 4) & 1u) << 1) | (b1 & 1u);            // low bit is the protection bit (:377-383)
    if (o.version == 1) return false;
    o.layer = (b1 >> 1) & 3u;
    if (o.layer == 0) return false;
    o.not_prot = b1 & 1u;
    o.eeee = b2 >> 4;
    if (o.eeee == 0 || o.eeee == 15) return false;
    o.ff = (b2 & 0xF) >> 2;
    if (o.ff == 3) return false;                                 // sample rate 0 -> InvalidData
    o.padded = (b2 >> 1) & 1u;
    o.chmode = b3 >> 6;
    return true;
}
__host__ __device__ inline bool match_ref_bits(const HdrInfo& a, const HdrInfo& b) {   // mpeg.rs:194-204
    // sr = base(version) * factor(ff): with equal versions, equal sr <=> equal ff
    return a.version == b.version && a.layer == b.layer && a.ff == b.ff && a.chmode == b.chmode && a.not_prot == b.not_prot;
}
// compute_frame_len (mpeg.rs:207-234) in the same f64 arithmetic; false on "Frame length too small"
__host__ __device__ inline bool frame_len_bits(const HdrInfo& o, uint32_t& payload, uint32_t& skip) {
    const uint32_t rates[14] = {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160};   // BITRATES column 4
    const double base = o.version == 3 ? 32000.0 : o.version == 2 ? 16000.0 : 8000.0;
    const double sr = o.ff == 0 ? base * 1.378125 : o.ff == 1 ? base * 1.5 : base;
    const double br = (double)rates[o.eeee - 1] * 1000.0;
    const uint32_t layer = o.layer == 1 ? 3 : o.layer == 2 ? 2 : 1;        // Header::format
    double fl;
    if (layer == 1) { fl = 12.0 * br; fl = fl / sr; fl = fl * 4.0; }
    else { fl = 144.0 * br; fl = fl / sr; }
    if (fl < 20.0) return false;
    const bool prot = o.not_prot == 0;
    payload = (uint32_t)fl - (prot ? 20u : 4u) + (o.padded ? 1u : 0u);
    skip = prot ? 6u : 4u;
    return true;
}

constexpr uint32_t kHdrBins = 1u << 21;         // the 11 sync bits are fixed: 21 free header bits

__global__ void mpeg_hist(const uint32_t* __restrict__ hdr, unsigned long long n, uint32_t* __restrict__ hist) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounds = (n + stride - 1) / stride;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned long long r = 0; r < rounds; ++r, i += stride) {
        const bool on = i < n;
        const uint32_t key = on ? (hdr[i] & (kHdrBins - 1)) : 0xFFFFFFFFu;
        // warp-aggregated atomics: one add per distinct key per warp (the dominant header is hot)
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        if (on && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + key, (uint32_t)__popc(peers));
    }
}

// most frequent header value that parses; ties -> smallest header value (the reference follows HashMap
// iteration order there, i.e. it is nondeterministic; mpeg.rs:53-73)
__global__ void mpeg_pick_ref(const uint32_t* __restrict__ hist, unsigned long long* __restrict__ best) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0;
    if (i < kHdrBins) {
        const uint32_t c = hist[i];
        HdrInfo o;
        if (c && parse_header_bits(0xFFE00000u | i, o)) key = ((unsigned long long)c << 21) | (kHdrBins - 1 - i);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, d);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0 && key) atomicMax(best, key);
}

// per-candidate validity against the reference header; first candidate index of every valid header value
__device__ __forceinline__ bool cand_valid(uint32_t h, const HdrInfo& ref, uint32_t& payload, uint32_t& skip) {
    HdrInfo o;
    if (!parse_header_bits(h, o)) return false;
    if (!match_ref_bits(ref, o)) return false;
    return frame_len_bits(o, payload, skip);
}

__global__ void mpeg_first_index(const uint32_t* __restrict__ hdr, unsigned long long n, uint32_t ref_header,
                                 unsigned long long* __restrict__ first) {
    HdrInfo ref;
    parse_header_bits(ref_header, ref);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t h = hdr[i];
        uint32_t pl, sk;
        if (cand_valid(h, ref, pl, sk)) atomicMin(first + (h & (kHdrBins - 1)), i);
    }
}

constexpr int kClsThreads = 256;
constexpr int kClsPerThread = 8;
constexpr int kClsBlock = kClsThreads * kClsPerThread;      // candidates per block

// pass 1 (emit == 0): per-block output counts.  pass 2 (emit == 1): write frames[*].file_pos in order.
__global__ void __launch_bounds__(kClsThreads)
mpeg_classify(const unsigned long long* __restrict__ pos, const uint32_t* __restrict__ hdr, unsigned long long n,
              uint32_t ref_header, const unsigned long long* __restrict__ first, int compat, unsigned long long file_len,
              unsigned long long* __restrict__ block_counts, const unsigned long long* __restrict__ block_base, int emit,
              unsigned long long* __restrict__ out, unsigned long long cap, uint32_t* __restrict__ err) {
    __shared__ uint32_t s_warp[kClsThreads / 32];
    HdrInfo ref;
    parse_header_bits(ref_header, ref);
    const unsigned long long i0 = (unsigned long long)blockIdx.x * kClsBlock + (unsigned long long)threadIdx.x * kClsPerThread;
    uint32_t cnt[kClsPerThread];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kClsPerThread; ++k) {
        const unsigned long long i = i0 + k;
        cnt[k] = 0;
        if (i < n) {
            const uint32_t h = hdr[i];
            uint32_t pl, sk;
            if (cand_valid(h, ref, pl, sk)) {
                cnt[k] = 1 + ((compat && first[h & (kHdrBins - 1)] == i) ? 1u : 0u);
                if (pos[i] + sk + pl > file_len) atomicExch(err, 1u);       // mpeg.rs:95-97 indexes past EOF
            }
        }
        mine += cnt[k];
    }
    // block-wide exclusive scan of `mine`
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if ((int)lane >= d) inc += a;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kClsThreads / 32; ++k) {
        if (k < (int)warp) wbase += s_warp[k];
        total += s_warp[k];
    }
    if (!emit) {
        if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
        return;
    }
    unsigned long long o = block_base[blockIdx.x] + wbase + (inc - mine);
#pragma unroll
    for (int k = 0; k < kClsPerThread; ++k) {
        for (uint32_t r = 0; r < cnt[k]; ++r) {
            if (o < cap) out[o] = pos[i0 + k];
            o += 1;
        }
    }
}

// exclusive scan of the block counts (one block; the array is small)
__global__ void mpeg_scan_blocks(const unsigned long long* __restrict__ counts, unsigned long long* __restrict__ base,
                                 unsigned long long nb, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_part[1024];
    const unsigned long long per = (nb + blockDim.x - 1) / blockDim.x;
    const unsigned long long a = threadIdx.x * per, b = min(nb, a + per);
    unsigned long long sum = 0;
    for (unsigned long long i = a; i < b; ++i) sum += counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (unsigned int t = 0; t < blockDim.x; ++t) { const unsigned long long v = s_part[t]; s_part[t] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    unsigned long long run = s_part[threadIdx.x];
    for (unsigned long long i = a; i < b; ++i) { base[i] = run; run += counts[i]; }
}

// K9 (mpeg.rs:86-121): gather the payload bytes of the indexed frames, in index order
__global__ void mpeg_payload_sizes(const unsigned long long* __restrict__ offs, unsigned long long n,
                                   const uint8_t* __restrict__ bytes, unsigned long long* __restrict__ sizes) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint8_t* p = bytes + offs[i];
        const uint32_t h = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        HdrInfo o;
        uint32_t pl = 0, sk = 0;
        if (parse_header_bits(h, o)) frame_len_bits(o, pl, sk);
        sizes[i] = pl;
    }
}

__global__ void mpeg_payload_gather(const unsigned long long* __restrict__ offs, const unsigned long long* __restrict__ dst_off,
                                    unsigned long long n, const uint8_t* __restrict__ bytes, uint8_t* __restrict__ out) {
    // one warp per frame, byte copies (frame payloads are a few hundred bytes at arbitrary alignment)
    const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (unsigned long long f = (((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); f < n; f += warps) {
        const uint8_t* p = bytes + offs[f];
        const uint32_t h = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        HdrInfo o;
        uint32_t pl = 0, sk = 0;
        if (parse_header_bits(h, o)) frame_len_bits(o, pl, sk);
        const uint8_t* src = p + sk;
        uint8_t* dst = out + dst_off[f];
        for (uint32_t k = lane; k < pl; k += 32) dst[k] = src[k];
    }
}

struct DevFree {
    std::vector<void*> ptrs;
    ~DevFree() { for (void* p : ptrs) if (p) cudaFree(p); }
    template <typename T> cudaError_t alloc(T** out, size_t bytes) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = reinterpret_cast<T*>(p);
        return e;
    }
};

int run_scan(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, uint64_t* d_pos, uint32_t* d_hdr, uint64_t cap,
             uint64_t* n_out) {
    if (((uintptr_t)d_bytes & 15) != 0) return blast::set_error(BLAST_ERR_ARG, "mpeg scan: d_bytes must be 16-byte aligned");
    *n_out = 0;
    if (len == 0) return BLAST_OK;
    const unsigned long long n_tiles = (len + kTileBytes - 1) / kTileBytes;
    // tile descriptors + control block live in context scratch (no cudaMalloc / cudaFree on this path)
    TileDesc* desc = static_cast<TileDesc*>(blast::scratch(ctx, 0, n_tiles * sizeof(TileDesc) + 256));
    if (!desc) return BLAST_ERR_CUDA;
    ScanCtl* ctl = reinterpret_cast<ScanCtl*>(reinterpret_cast<uint8_t*>(desc) + ((n_tiles * sizeof(TileDesc) + 127) & ~127ull));
    ScanCtl* h_ctl = static_cast<ScanCtl*>(blast::mailbox(ctx));
    if (!h_ctl) return BLAST_ERR_CUDA;
    BLAST_CUDA_TRY(cudaMemsetAsync(desc, 0, ((n_tiles * sizeof(TileDesc) + 127) & ~127ull) + sizeof(ScanCtl), ctx->stream));
    static int per_sm = 0;
    if (per_sm == 0) BLAST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mpeg_sync_scan, kScanThreads, 0));
    const unsigned grid = (unsigned)std::min<unsigned long long>(n_tiles, (unsigned long long)ctx->sm_count * std::max(per_sm, 1));
    mpeg_sync_scan<<<grid, kScanThreads, 0, ctx->stream>>>(d_bytes, len, desc, n_tiles, ctl,
                                                           reinterpret_cast<unsigned long long*>(d_pos), d_hdr, cap);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    BLAST_CUDA_TRY(cudaMemcpyAsync(h_ctl, ctl, sizeof(ScanCtl), cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const ScanCtl h = *h_ctl;
    *n_out = h.total;
    if (h.panic) return blast::set_error(BLAST_ERR_REF_PANIC, "index out of bounds: the scan reaches a trailing 0xFF (mpeg.rs:20)");
    if (h.total > cap) return blast::set_error(BLAST_ERR_CAPACITY, "mpeg scan: %llu candidates, capacity %llu",
                                               (unsigned long long)h.total, (unsigned long long)cap);
    return BLAST_OK;
}

}  // namespace

extern "C" {

int blast_mpeg_header_info(uint32_t header, blast_mpeg_header* out) {
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_mpeg_header_info: null out");
    *out = blast_mpeg_header{};
    HdrInfo o;
    // error classification as parse_header: unsupported version / layer / bitrate, invalid sample rate
    const uint32_t b1 = (header >> 16) & 0xFF, b2 = (header >> 8) & 0xFF;
    const uint32_t version = (((b1 >> 4) & 1u) << 1) | (b1 & 1u);
    if (!parse_header_bits(header, o)) {
        const bool invalid = version != 1 && ((b1 >> 1) & 3u) != 0 && (b2 >> 4) != 0 && (b2 >> 4) != 15;
        out->status = invalid ? BLAST_ERR_INVALID_DATA : BLAST_ERR_UNSUPPORTED_FORMAT;
        return BLAST_OK;
    }
    out->ok = 1;
    out->version_id = o.version;
    out->layer = o.layer == 1 ? 3 : o.layer == 2 ? 2 : 1;
    out->is_protected = o.not_prot == 0;
    out->padded = o.padded;
    out->channel_mode = o.chmode;
    const uint32_t rates[14] = {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160};
    out->bitrate = rates[o.eeee - 1];
    const double base = o.version == 3 ? 32000.0 : o.version == 2 ? 16000.0 : 8000.0;
    out->sample_rate = o.ff == 0 ? base * 1.378125 : o.ff == 1 ? base * 1.5 : base;
    uint32_t pl = 0, sk = 0;
    out->frame_len_ok = frame_len_bits(o, pl, sk) ? 1 : 0;
    out->payload_len = pl;
    out->skip = out->is_protected ? 6 : 4;
    return BLAST_OK;
}

int blast_mpeg_scan_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, uint64_t* d_pos_out, uint32_t* d_hdr_out,
                        uint64_t cap, uint64_t* n_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(n_out && (d_bytes || len == 0) && ((d_pos_out && d_hdr_out) || cap == 0), BLAST_ERR_ARG,
                  "blast_mpeg_scan_dev: null argument");
    return run_scan(ctx, d_bytes, len, d_pos_out, d_hdr_out, cap, n_out);
}

int blast_mpeg_index_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, int reference_compat,
                         uint64_t* d_offsets_out, uint64_t cap, uint64_t* n_offsets_out, uint32_t* ref_header_out,
                         uint64_t* n_candidates_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(n_offsets_out && (d_bytes || len == 0), BLAST_ERR_ARG, "blast_mpeg_index_dev: null argument");
    *n_offsets_out = 0;
    if (n_candidates_out) *n_candidates_out = 0;
    DevFree mem;
    // candidates: one pass with a guessed capacity (an MP3 stream has one sync per ~400 bytes, random
    // bytes one per 2,048); an exact second pass only if the guess was too small
    uint64_t n_cand = 0, guess = len / 32 + 4096;
    unsigned long long* d_pos = nullptr;
    uint32_t* d_hdr = nullptr;
    BLAST_CUDA_TRY(mem.alloc(&d_pos, guess * 8));
    BLAST_CUDA_TRY(mem.alloc(&d_hdr, guess * 4));
    int rc = run_scan(ctx, d_bytes, len, reinterpret_cast<uint64_t*>(d_pos), d_hdr, guess, &n_cand);
    if (rc == BLAST_ERR_CAPACITY) {
        BLAST_CUDA_TRY(mem.alloc(&d_pos, n_cand * 8));
        BLAST_CUDA_TRY(mem.alloc(&d_hdr, n_cand * 4));
        uint64_t n2 = 0;
        rc = run_scan(ctx, d_bytes, len, reinterpret_cast<uint64_t*>(d_pos), d_hdr, n_cand, &n2);
    }
    if (rc != BLAST_OK) return rc;
    if (n_candidates_out) *n_candidates_out = n_cand;
    if (n_cand == 0) return blast::set_error(BLAST_ERR_REF_PANIC, "no sync candidates: the reference indexes an empty list (mpeg.rs:64)");

    uint32_t* d_hist = nullptr;
    unsigned long long *d_best = nullptr, *d_first = nullptr, *d_bc = nullptr, *d_bb = nullptr, *d_total = nullptr;
    uint32_t* d_err = nullptr;
    const unsigned long long n_blocks = (n_cand + kClsBlock - 1) / kClsBlock;
    BLAST_CUDA_TRY(mem.alloc(&d_hist, kHdrBins * sizeof(uint32_t)));
    BLAST_CUDA_TRY(mem.alloc(&d_best, 8));
    BLAST_CUDA_TRY(mem.alloc(&d_first, (size_t)kHdrBins * 8));
    BLAST_CUDA_TRY(mem.alloc(&d_bc, n_blocks * 8));
    BLAST_CUDA_TRY(mem.alloc(&d_bb, n_blocks * 8));
    BLAST_CUDA_TRY(mem.alloc(&d_total, 8));
    BLAST_CUDA_TRY(mem.alloc(&d_err, 4));
    BLAST_CUDA_TRY(cudaMemsetAsync(d_hist, 0, kHdrBins * sizeof(uint32_t), ctx->stream));
    BLAST_CUDA_TRY(cudaMemsetAsync(d_best, 0, 8, ctx->stream));
    BLAST_CUDA_TRY(cudaMemsetAsync(d_first, 0xFF, (size_t)kHdrBins * 8, ctx->stream));
    BLAST_CUDA_TRY(cudaMemsetAsync(d_err, 0, 4, ctx->stream));
    const unsigned g = (unsigned)std::min<unsigned long long>((n_cand + 255) / 256, (unsigned long long)ctx->sm_count * 16);
    mpeg_hist<<<g, 256, 0, ctx->stream>>>(d_hdr, n_cand, d_hist);
    mpeg_pick_ref<<<kHdrBins / 256, 256, 0, ctx->stream>>>(d_hist, d_best);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    unsigned long long best = 0;
    BLAST_CUDA_TRY(cudaMemcpyAsync(&best, d_best, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (best == 0) return blast::set_error(BLAST_ERR_REF_PANIC, "no parsable header: the reference indexes past its candidate list (mpeg.rs:64)");
    const uint32_t ref_header = 0xFFE00000u | (kHdrBins - 1 - (uint32_t)(best & (kHdrBins - 1)));
    if (ref_header_out) *ref_header_out = ref_header;
    if (reference_compat) {
        mpeg_first_index<<<g, 256, 0, ctx->stream>>>(d_hdr, n_cand, ref_header, d_first);
        ctx->launches += 1;
    }
    mpeg_classify<<<(unsigned)n_blocks, kClsThreads, 0, ctx->stream>>>(d_pos, d_hdr, n_cand, ref_header, d_first, reference_compat,
                                                                     len, d_bc, nullptr, 0, nullptr, 0, d_err);
    mpeg_scan_blocks<<<1, 1024, 0, ctx->stream>>>(d_bc, d_bb, n_blocks, d_total);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    unsigned long long total = 0;
    uint32_t err = 0;
    BLAST_CUDA_TRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *n_offsets_out = total;
    if (err && reference_compat)
        return blast::set_error(BLAST_ERR_REF_PANIC, "a frame payload extends past the end of the file (mpeg.rs:96 indexes out of bounds)");
    if (d_offsets_out) {
        if (total > cap) return blast::set_error(BLAST_ERR_CAPACITY, "mpeg index: %llu offsets, capacity %llu", total, (unsigned long long)cap);
        mpeg_classify<<<(unsigned)n_blocks, kClsThreads, 0, ctx->stream>>>(d_pos, d_hdr, n_cand, ref_header, d_first, reference_compat,
                                                                         len, d_bc, d_bb, 1, reinterpret_cast<unsigned long long*>(d_offsets_out),
                                                                         cap, d_err);
        BLAST_CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return BLAST_OK;
}

int blast_mpeg_gather_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, const uint64_t* d_offsets, uint64_t n_offsets,
                          uint8_t* d_payload_out, uint64_t cap, uint64_t* payload_len_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(payload_len_out && (d_bytes || len == 0) && (d_offsets || n_offsets == 0), BLAST_ERR_ARG,
                  "blast_mpeg_gather_dev: null argument");
    *payload_len_out = 0;
    if (n_offsets == 0) return BLAST_OK;
    DevFree mem;
    unsigned long long *d_sizes = nullptr, *d_base = nullptr, *d_total = nullptr;
    BLAST_CUDA_TRY(mem.alloc(&d_sizes, n_offsets * 8));
    BLAST_CUDA_TRY(mem.alloc(&d_base, n_offsets * 8));
    BLAST_CUDA_TRY(mem.alloc(&d_total, 8));
    const unsigned g = (unsigned)std::min<unsigned long long>((n_offsets + 255) / 256, (unsigned long long)ctx->sm_count * 16);
    mpeg_payload_sizes<<<g, 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long*>(d_offsets), n_offsets, d_bytes, d_sizes);
    mpeg_scan_blocks<<<1, 1024, 0, ctx->stream>>>(d_sizes, d_base, n_offsets, d_total);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    unsigned long long total = 0;
    BLAST_CUDA_TRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *payload_len_out = total;
    if (!d_payload_out) return BLAST_OK;
    if (total > cap) return blast::set_error(BLAST_ERR_CAPACITY, "mpeg gather: %llu payload bytes, capacity %llu", total, (unsigned long long)cap);
    const unsigned gw = (unsigned)std::min<unsigned long long>((n_offsets + 7) / 8, (unsigned long long)ctx->sm_count * 16);
    mpeg_payload_gather<<<gw, 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long*>(d_offsets), d_base, n_offsets, d_bytes, d_payload_out);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BLAST_OK;
}

// mpeg::parse drop-in on a host buffer: upload, scan, index, gather, copy back (outputs nullable)
int blast_mpeg_parse(blast_ctx* ctx, const uint8_t* bytes, uint64_t len, int reference_compat, uint64_t* offsets_out,
                     uint64_t offsets_cap, uint64_t* n_offsets_out, uint32_t* ref_header_out, uint64_t* n_candidates_out,
                     uint8_t* payload_out, uint64_t payload_cap, uint64_t* payload_len_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(n_offsets_out && (bytes || len == 0), BLAST_ERR_ARG, "blast_mpeg_parse: null argument");
    DevFree mem;
    uint8_t* d_bytes = nullptr;
    BLAST_CUDA_TRY(mem.alloc(&d_bytes, ((len + 255) & ~255ull) + 256));
    if (len) BLAST_CUDA_TRY(cudaMemcpyAsync(d_bytes, bytes, len, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t n_off = 0;
    int rc = blast_mpeg_index_dev(ctx, d_bytes, len, reference_compat, nullptr, 0, &n_off, ref_header_out, n_candidates_out);
    *n_offsets_out = n_off;
    if (rc != BLAST_OK) return rc;
    const bool want_payload = payload_out != nullptr || payload_len_out != nullptr;
    if (!offsets_out && !want_payload) return BLAST_OK;
    if (offsets_out && n_off > offsets_cap) return blast::set_error(BLAST_ERR_CAPACITY, "blast_mpeg_parse: %llu offsets, capacity %llu",
                                                                    (unsigned long long)n_off, (unsigned long long)offsets_cap);
    uint64_t* d_off = nullptr;
    BLAST_CUDA_TRY(mem.alloc(&d_off, n_off * 8));
    uint64_t n2 = 0;
    if ((rc = blast_mpeg_index_dev(ctx, d_bytes, len, reference_compat, d_off, n_off, &n2, nullptr, nullptr)) != BLAST_OK) return rc;
    if (offsets_out && n_off) BLAST_CUDA_TRY(cudaMemcpyAsync(offsets_out, d_off, n_off * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_payload) {
        uint64_t plen = 0;
        if ((rc = blast_mpeg_gather_dev(ctx, d_bytes, len, d_off, n_off, nullptr, 0, &plen)) != BLAST_OK) return rc;
        if (payload_len_out) *payload_len_out = plen;
        if (payload_out) {
            if (plen > payload_cap) return blast::set_error(BLAST_ERR_CAPACITY, "blast_mpeg_parse: %llu payload bytes, capacity %llu",
                                                            (unsigned long long)plen, (unsigned long long)payload_cap);
            uint8_t* d_pay = nullptr;
            BLAST_CUDA_TRY(mem.alloc(&d_pay, plen));
            if ((rc = blast_mpeg_gather_dev(ctx, d_bytes, len, d_off, n_off, d_pay, plen, &plen)) != BLAST_OK) return rc;
            if (plen) BLAST_CUDA_TRY(cudaMemcpyAsync(payload_out, d_pay, plen, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BLAST_OK;
}

}  // extern "C"
