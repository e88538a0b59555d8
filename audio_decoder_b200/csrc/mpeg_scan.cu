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
// K7 reads every input byte exactly once (1 B/byte of HBM traffic + 12 B per candidate, plus 6 B per candidate
// of temp list written and re-read).  The greedy rule is a 4-state machine (state = header bytes still to
// skip); a byte range acts on it as a map {0..3} -> {0..3} plus a candidate count per entry state, and maps
// compose associatively.  Six launches, none of which ever waits for another thread block:
//   K7a mpeg_walk       one warp per 32 KiB span: finds the span's candidates under entry state 0 and leaves
//                       them in the span's slot of a temp list, plus a 16-byte record (map, counts per entry
//                       state, "head-sensitive" flag)
//   K7b mpeg_span_fold / mpeg_block_chain / mpeg_span_fold<APPLY>   an ordinary three-step scan over the
//                       records (8 MB for 16 GiB of input): true entry state and first global candidate index
//                       of every span
//   K7c mpeg_compact    moves the lists to their final, position-ordered place (coalesced); K7d mpeg_redo re-walks
//                       the few spans whose list does not apply (see below)
// A span's entry state only matters when its very first three bytes hold a raw sync ("head-sensitive",
// ~0.15 % of spans on random data, every span on 0xFF floods): K7a then also counts it under the other three
// entry states, and K7c re-walks it with direct emission if its true entry state is not 0 (same for spans
// whose candidates overflow their slot).
//
// The walk.  16 rounds of 2 KiB per span.  The bytes go global -> shared memory by coalesced 16-byte
// cp.async (LDGSTS: L2 only, no registers), kStages rounds deep, into padded rows from which each lane reads
// ITS 64 consecutive bytes with conflict-free LDS.128 (the per-lane LDG.128 at a 64-byte lane stride of v2
// cost 16 L1 wavefronts per instruction).  Per lane: a 5-instruction-per-word SWAR test gives 0x80 flags,
// two IDP4A per word pair pack them into the 64-bit "raw sync" mask M, and the greedy selection is resolved
// from the lane's entry state — in the common case (no two raw syncs within 4 bytes, no sync pair straddling
// a lane boundary) every raw sync is selected and nothing iterates.  Headers are read from the staged rows:
// no byte is fetched from HBM twice.
//
// Measured dead ends, kept for the record (C5, 16 GiB): a single-pass chained scan with decoupled look-back
// was built first in four variants — v1 16 KiB tiles (0.57 TB/s), v2 256 KiB tiles + two passes over the tile
// (2.0 TB/s, 10 instructions per byte, headers re-read from HBM), v3 warp-per-tile without barriers (1.1 TB/s:
// ~3,500 tiles in lock step, look-back chains 8x longer), v3 with a dedicated chaining warp and
// software-pipelined hand-off (1.8 TB/s).  With the look-back disabled the same walk ran at 4.3 TB/s: all
// tiles of a wave finish together, so every look-back reaches back a whole wave over loaded-L2 round trips
// and every CTA advances at the pace of the slowest one.  Hence three independent launches.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "blast_internal.h"

namespace {

constexpr int kWalkers = 8;                                 // warps per CTA
constexpr int kScanThreads = kWalkers * 32;
constexpr int kCtasPerSm = 3;
constexpr int kChunk = 64;                                  // bytes per lane per round
constexpr int kRowBytes = kChunk + 16;                      // staged chunk + padding: conflict-free LDS.128 at 64-byte lane stride
constexpr int kStageBytes = 32 * kRowBytes;                 // one round of one warp in shared memory
constexpr int kStages = 3;                                  // staged rounds per warp (cp.async groups)
constexpr int kRounds = 16;                                 // rounds per span
constexpr int kRoundBytes = 32 * kChunk;                    // 2 KiB
constexpr int kSpanBytes = kRounds * kRoundBytes;           // 32 KiB per warp
constexpr int kCandCap = 256;                               // candidates per span that fit its slot of the temp lists
constexpr size_t kScanSmem = (size_t)kWalkers * kStages * kStageBytes;
constexpr uint32_t kIdentityMap = 0xE4;                     // s -> s for s = 0..3, two bits each
struct ScanCtl {
    unsigned long long next_tile; // span counter of the walk kernel
    unsigned long long total;     // candidates found (under the entry state of the last chain)
    unsigned long long count[4];  // candidates under entry state s (sharded scans: the range's aggregate)
    uint32_t exit_state[4];       // state after the last byte under entry state s
    uint32_t panic;               // the reference indexes out of bounds (last byte 0xFF reached with state 0)
    uint32_t pad;
};

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
template <typename T>
__device__ __forceinline__ T pick4(T a0, T a1, T a2, T a3, uint32_t i) {       // register-only a[i]
    return i == 0 ? a0 : i == 1 ? a1 : i == 2 ? a2 : a3;
}

// raw sync flags of one word: 0x80 in byte i <=> b[i] == 0xFF && (b[i+1] & 0xE0) == 0xE0, where `next` is
// the following word.  z = b[i] & (b[i+1] | 0x1F) is 0xFF exactly then; the byte-wise == 0xFF test is
// carry-free (low 7 bits + 1 reaches bit 7 iff they are all ones).
__device__ __forceinline__ uint32_t sync_flags(uint32_t w, uint32_t next) {
    const uint32_t sh = __funnelshift_r(w, next, 8);
    const uint32_t z = w & (sh | 0x1F1F1F1Fu);
    const uint32_t t = (z & 0x7F7F7F7Fu) + 0x01010101u;
    return t & z & 0x80808080u;
}

enum : int { kModeCompact = 0, kModeCount = 1, kModeEmit = 2 };

struct SpanOut {
    uint32_t total;        // candidates (uniform)
    uint32_t exit_state;   // state after the last byte of the tile (uniform)
    bool sens;             // the first three bytes of the tile hold a raw sync: the entry state matters
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// One walk over a warp's span with a KNOWN entry state.
//  kModeCompact: candidates -> shared-memory list (s_off, s_hdr) while it has room (total may exceed kCandCap)
//  kModeCount:   counts only
//  kModeEmit:    candidates -> out_pos / out_hdr at index gbase + k
// The bytes travel global -> shared memory with coalesced 16-byte cp.async (LDGSTS, L2 only, no registers),
// kStages rounds deep, into rows of 64 + 16 bytes so that every lane then reads ITS 64 consecutive bytes with
// four conflict-free LDS.128.  NEAR_END instances carry the end-of-buffer handling; interior spans do not.
template <bool NEAR_END>
__device__ __noinline__ SpanOut walk_span(const int MODE, const uint8_t* __restrict__ bytes, const unsigned long long n,
                                          const unsigned long long span0, const uint32_t entry, const uint32_t lane,
                                          const uint32_t ring,           // shared-memory address of this warp's stage ring
                                          uint16_t* __restrict__ s_off, uint32_t* __restrict__ s_hdr,
                                          unsigned long long* __restrict__ out_pos, uint32_t* __restrict__ out_hdr,
                                          const unsigned long long cap, const unsigned long long gbase,
                                          const unsigned long long pos_offset, ScanCtl* __restrict__ ctl) {
    SpanOut o;
    o.total = 0;
    o.sens = false;
    uint32_t carry = entry;                                            // entry state of lane 0 this round (uniform)
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint8_t* src = bytes + span0 + 16u * lane;                   // lane's first 16-byte unit of the next round to fetch
    const uint32_t dst = (lane >> 2) * kRowBytes + (lane & 3) * 16;    // unit j = lane + 32 q -> row (j >> 2) = lane/4 + 8 q
    const uint32_t ring_end = ring + (uint32_t)(kStages * kStageBytes);
    uint32_t st_fill = ring;                                           // stage the next fetched round goes to
    unsigned long long fetched = span0 + 16ull * lane;                 // position of `src` (end-of-buffer test only)

    auto issue = [&](bool any) {
        if (any) {
            const uint32_t st = st_fill + dst;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (!NEAR_END || fetched + 512ull * q < n) cp_async16(st + (uint32_t)q * 8u * kRowBytes, src + 512 * q);
            src += kRoundBytes;
            if (NEAR_END) fetched += kRoundBytes;
            st_fill += kStageBytes;
            if (st_fill == ring_end) st_fill = ring;
        }
        cp_async_commit();
    };
    // first word after the span: lane 31's look-ahead in the last round
    uint32_t after_word = 0;
    if (lane == 31 && span0 + (unsigned long long)kSpanBytes < n)
        after_word = __ldg(reinterpret_cast<const uint32_t*>(bytes + span0 + kSpanBytes));

#pragma unroll
    for (int r = 0; r < kStages - 1; ++r) issue(true);
    uint32_t st_cur = ring;                                            // stage of the round being processed
#pragma unroll 1
    for (int r = 0; r < kRounds; ++r) {
        issue(r + kStages - 1 < kRounds);
        cp_async_wait<kStages - 2>();                                   // rounds r and r + 1 have landed (this lane's part)
        __syncwarp();                                                   // ... and everybody else's
        const uint32_t row = st_cur + lane * kRowBytes;
        uint32_t st_next = st_cur + kStageBytes;
        if (st_next == ring_end) st_next = ring;
        uint32_t w[17];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 v = lds128(row + 16u * q);
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
        {
            const uint32_t down = __shfl_down_sync(0xFFFFFFFFu, w[0], 1);
            uint32_t wrap = after_word;
            if (r + 1 < kRounds) wrap = lds32(st_next);
            w[16] = lane == 31 ? wrap : down;
        }
        const uint32_t rel = (uint32_t)r * kRoundBytes + lane * kChunk;          // offset of the lane's chunk in the span
        if (NEAR_END) {
            const unsigned long long base = span0 + rel;
            if (base + 68 > n) {                                        // the buffer ends inside this window
#pragma unroll
                for (int k = 0; k < 17; ++k) {
                    const unsigned long long p = base + 4ull * k;
                    if (p >= n) w[k] = 0;
                    else if (p + 4 > n) w[k] &= (1u << (8 * (uint32_t)(n - p))) - 1u;
                }
            }
        }
        // ---- raw sync mask M: bit i <=> b[i]==0xFF && (b[i+1]&0xE0)==0xE0
        uint32_t mlo, mhi;
        {
            uint32_t acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t f0 = sync_flags(w[2 * k], w[2 * k + 1]);
                const uint32_t f1 = sync_flags(w[2 * k + 1], w[2 * k + 2]);
                // flags are 0x80 per byte: the dot products leave (8 mask bits) << 7
                acc[k] = __dp4a(f1, 0x80402010u, __dp4a(f0, 0x08040201u, 0u));
            }
            mlo = (acc[0] >> 7) | (acc[1] << 1) | (acc[2] << 9) | (acc[3] << 17);
            mhi = (acc[4] >> 7) | (acc[5] << 1) | (acc[6] << 9) | (acc[7] << 17);
        }
        const unsigned long long M = ((unsigned long long)mhi << 32) | mlo;
        if (r == 0) o.sens = __shfl_sync(0xFFFFFFFFu, mlo & 7u, 0) != 0;

        // ---- greedy selection
        unsigned long long sel = 0;
        uint32_t e = 0;
        const uint32_t anym = __ballot_sync(0xFFFFFFFFu, (mlo | mhi) != 0u);
        if (anym || carry) {
            // fast path (practically always): no two raw syncs closer than 4 bytes inside a chunk, and no chunk whose
            // first 3 bytes hold a raw sync right after a chunk whose last 3 bytes do -> every raw sync is selected
            const uint32_t H = mhi >> 29;
            uint32_t pH = __shfl_up_sync(0xFFFFFFFFu, H, 1);
            if (lane == 0) pH = carry;
            const unsigned long long clash = M & ((M >> 1) | (M >> 2) | (M >> 3));
            const bool cross = pH != 0 && (mlo & 7u) != 0;
            uint32_t x;
            if (!__any_sync(0xFFFFFFFFu, clash != 0ull || cross)) {
                sel = M;
                x = H ? 32u - (uint32_t)__clz((int)H) : 0u;              // H in {1, 2, 4}: spill-over of the last header
                e = lane == 0 ? carry : 0u;                              // (only used by the trailing-0xFF check)
            } else {
                x = 0;
                bool known = lane == 0 || pH == 0 || (mlo & 7u) == 0;
                e = lane == 0 ? carry : 0u;
                if (known && M) sel = resolve(M, e, x);
                while (!__all_sync(0xFFFFFFFFu, known)) {
                    const uint32_t px = __shfl_up_sync(0xFFFFFFFFu, x, 1);
                    const bool pk = __shfl_up_sync(0xFFFFFFFFu, (int)known, 1) != 0;
                    if (!known && pk) { e = px; sel = resolve(M, e, x); known = true; }
                }
            }
            carry = __shfl_sync(0xFFFFFFFFu, x, 31);
        }

        if (NEAR_END) {
            const unsigned long long base = span0 + rel;
            // mpeg.rs:20: `reader[cur + 1]` with cur == n-1 panics when the scan reaches a trailing 0xFF
            if (MODE != kModeCount && n - 1 >= base && n - 1 < base + kChunk) {
                const uint32_t il = (uint32_t)(n - 1 - base);
                if (bytes[n - 1] == 0xFF) {
                    const unsigned long long before = il ? (sel & ((1ull << il) - 1ull)) : 0ull;
                    bool skipped = il < e;
                    if (before) skipped = skipped || (il - (uint32_t)(63 - __clzll((long long)before)) <= 3);
                    if (!skipped) ctl->panic = 1;
                }
            }
            // a header that would run past the end of the buffer is dropped (mpeg.rs:25-37)
            if (base + 67 >= n) {
                const unsigned long long lim = n > base + 3 ? n - 3 - base : 0;
                sel &= lim >= 64 ? ~0ull : ((1ull << lim) - 1ull);
            }
        }

        // ---- count + sink
        const uint32_t cnt = (uint32_t)__popcll(sel);
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, cnt != 0);
        if (bal) {
            uint32_t pre, round_total;
            if (__any_sync(0xFFFFFFFFu, cnt > 1)) {
                uint32_t inc = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if ((int)lane >= d) inc += a;
                }
                pre = inc - cnt;
                round_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            } else {
                pre = (uint32_t)__popc(bal & lt_mask);
                round_total = (uint32_t)__popc(bal);
            }
            if (MODE != kModeCount && cnt) {
                uint32_t k = o.total + pre;
                uint32_t slo = (uint32_t)sel, shi = (uint32_t)(sel >> 32);
                do {
                    // lowest selected bit (one candidate per lane is the rule: no 64-bit loop bookkeeping for it)
                    int i;
                    if (slo) { i = __ffs((int)slo) - 1; slo &= slo - 1; }
                    else { i = 32 + __ffs((int)shi) - 1; shi &= shi - 1; }
                    // header = 4 bytes at chunk offset i: two aligned words of the staged row (the 17th is w[16])
                    const uint32_t a = lds32(row + 4u * (uint32_t)(i >> 2));
                    const uint32_t b = (i >> 2) == 15 ? w[16] : lds32(row + 4u * (uint32_t)(i >> 2) + 4u);
                    const uint32_t h = __byte_perm(__funnelshift_r(a, b, 8 * (i & 3)), 0, 0x0123);     // big-endian (mpeg.rs:22-38)
                    if (MODE == kModeCompact) {
                        if (k < (uint32_t)kCandCap) {
                            s_off[k] = (uint16_t)(rel + (uint32_t)i);
                            s_hdr[k] = h;
                        }
                    } else {
                        const unsigned long long gi = gbase + k;
                        if (gi < cap) {
                            out_pos[gi] = pos_offset + span0 + rel + (unsigned long long)i;
                            out_hdr[gi] = h;
                        }
                    }
                    k += 1;
                } while (slo | shi);
            }
            o.total += round_total;
        }
        __syncwarp();                              // the stage is refilled by the next iteration's cp.async
        st_cur = st_next;
    }
    cp_async_wait<0>();
    o.exit_state = carry;
    return o;
}

__device__ __forceinline__ SpanOut walk_tile(const int MODE, const uint8_t* __restrict__ bytes, const unsigned long long n,
                                             const unsigned long long span0, const uint32_t entry, const uint32_t lane,
                                             const uint32_t ring, uint16_t* __restrict__ s_off, uint32_t* __restrict__ s_hdr,
                                             unsigned long long* __restrict__ out_pos, uint32_t* __restrict__ out_hdr,
                                             const unsigned long long cap, const unsigned long long gbase,
                                             const unsigned long long pos_offset, ScanCtl* __restrict__ ctl) {
    if (span0 + (unsigned long long)kSpanBytes + 80 > n)
        return walk_span<true>(MODE, bytes, n, span0, entry, lane, ring, s_off, s_hdr, out_pos, out_hdr, cap, gbase, pos_offset, ctl);
    return walk_span<false>(MODE, bytes, n, span0, entry, lane, ring, s_off, s_hdr, out_pos, out_hdr, cap, gbase, pos_offset, ctl);
}

// ---------------------------------------------------------------- K7a: walk
// Every warp takes spans from a global counter (the next id is fetched while the current span is walked) and
// leaves, per span: its candidates under entry state 0 in the span's slot of the temp lists, and a 16-byte
// record {map, sens, c0..c3}.  No warp ever waits for another one.
struct SpanRec {               // 16 bytes
    uint32_t meta;             // exit-state map (8 bits) | sens << 8
    uint32_t c01;              // candidates under entry state 0 | 1 << 16
    uint32_t c23;              // ... 2 | 3 << 16
    uint32_t pad;
};

__global__ void __launch_bounds__(kScanThreads, kCtasPerSm)
mpeg_walk(const uint8_t* __restrict__ bytes, unsigned long long n, unsigned long long n_spans, ScanCtl* __restrict__ ctl,
          SpanRec* __restrict__ recs, uint16_t* __restrict__ t_off, uint32_t* __restrict__ t_hdr) {
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ring = smem_addr(smem_dyn) + warp * (uint32_t)(kStages * kStageBytes);
    unsigned long long span = 0;
    if (lane == 0) span = atomicAdd(&ctl->next_tile, 1ull);
    span = __shfl_sync(0xFFFFFFFFu, span, 0);
    while (span < n_spans) {
        unsigned long long next = 0;
        if (lane == 0) next = atomicAdd(&ctl->next_tile, 1ull);       // latency hidden behind the walk
        const unsigned long long span0 = span * (unsigned long long)kSpanBytes;
        uint16_t* g_off = t_off + span * (unsigned long long)kCandCap;
        uint32_t* g_hdr = t_hdr + span * (unsigned long long)kCandCap;
        const SpanOut w0 = walk_tile(kModeCompact, bytes, n, span0, 0u, lane, ring, g_off, g_hdr, nullptr, nullptr, 0ull, 0ull, 0ull, ctl);
        uint32_t amap = w0.exit_state * 0x55u;
        uint32_t c0 = w0.total, c1 = w0.total, c2 = w0.total, c3 = w0.total;
        if (w0.sens) {                                   // rare: the span's first 3 bytes hold a raw sync
            const SpanOut w1 = walk_tile(kModeCount, bytes, n, span0, 1u, lane, ring, g_off, g_hdr, nullptr, nullptr, 0ull, 0ull, 0ull, ctl);
            const SpanOut w2 = walk_tile(kModeCount, bytes, n, span0, 2u, lane, ring, g_off, g_hdr, nullptr, nullptr, 0ull, 0ull, 0ull, ctl);
            const SpanOut w3 = walk_tile(kModeCount, bytes, n, span0, 3u, lane, ring, g_off, g_hdr, nullptr, nullptr, 0ull, 0ull, 0ull, ctl);
            amap = w0.exit_state | (w1.exit_state << 2) | (w2.exit_state << 4) | (w3.exit_state << 6);
            c1 = w1.total; c2 = w2.total; c3 = w3.total;
        }
        if (lane == 0) {
            SpanRec r;
            r.meta = amap | (w0.sens ? 0x100u : 0u);
            r.c01 = c0 | (c1 << 16);
            r.c23 = c2 | (c3 << 16);
            r.pad = 0;
            *reinterpret_cast<uint4*>(recs + span) = *reinterpret_cast<const uint4*>(&r);
        }
        span = __shfl_sync(0xFFFFFFFFu, next, 0);
    }
}

// ---------------------------------------------------------------- K7b: scan over the span records
// (map, counts) aggregates compose associatively, so this is an ordinary three-step scan: fold blocks of 2,048
// records (K7b-1), chain the block aggregates (K7b-2, one warp), replay every block with its true entry state
// (K7b-3 = the same kernel as K7b-1 in APPLY mode) leaving (entry state, first candidate index) per span.
// 16 GiB of input are 524,288 records = 8 MB.
struct Agg {
    uint32_t map;
    uint32_t c[4];
};
__device__ __forceinline__ Agg agg_identity() {
    Agg r;
    r.map = kIdentityMap;
    r.c[0] = r.c[1] = r.c[2] = r.c[3] = 0;
    return r;
}
__device__ __forceinline__ Agg agg_then(const Agg& a, const Agg& b) {            // a first, then b
    Agg r;
    r.map = map_after(a.map, b.map);
#pragma unroll
    for (int s = 0; s < 4; ++s) r.c[s] = a.c[s] + pick4(b.c[0], b.c[1], b.c[2], b.c[3], map_get(a.map, s));
    return r;
}
__device__ __forceinline__ Agg agg_shfl_up(const Agg& a, int d) {
    Agg r;
    r.map = __shfl_up_sync(0xFFFFFFFFu, a.map, d);
#pragma unroll
    for (int s = 0; s < 4; ++s) r.c[s] = __shfl_up_sync(0xFFFFFFFFu, a.c[s], d);
    return r;
}
__device__ __forceinline__ uint32_t rec_count(const SpanRec& r, uint32_t s) {
    const uint32_t w = (s & 2) ? r.c23 : r.c01;
    return (s & 1) ? (w >> 16) : (w & 0xFFFFu);
}
__device__ __forceinline__ Agg agg_of(const SpanRec& r) {
    Agg a;
    a.map = r.meta & 0xFFu;
    a.c[0] = r.c01 & 0xFFFFu; a.c[1] = r.c01 >> 16; a.c[2] = r.c23 & 0xFFFFu; a.c[3] = r.c23 >> 16;
    return a;
}

constexpr int kFoldThreads = 256;
constexpr int kFoldPerThread = 8;
constexpr int kFoldBlock = kFoldThreads * kFoldPerThread;       // records per block

struct BlockAgg {              // 32 bytes
    uint32_t map, c[4];        // K7b-1: the block's aggregate
    uint32_t entry;            // K7b-2: entry state of the block
    unsigned long long base;   // K7b-2: candidates before the block
};

template <bool APPLY>
__global__ void __launch_bounds__(kFoldThreads)
mpeg_span_fold(const SpanRec* __restrict__ recs, unsigned long long n_spans, BlockAgg* __restrict__ blocks,
               uint8_t* __restrict__ span_entry, unsigned long long* __restrict__ span_base) {
    __shared__ Agg s_warp[kFoldThreads / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long i0 = (unsigned long long)blockIdx.x * kFoldBlock + (unsigned long long)threadIdx.x * kFoldPerThread;
    SpanRec r[kFoldPerThread];
    Agg mine = agg_identity();
#pragma unroll
    for (int k = 0; k < kFoldPerThread; ++k) {
        r[k].meta = kIdentityMap; r[k].c01 = 0; r[k].c23 = 0; r[k].pad = 0;
        if (i0 + k < n_spans) {
            const uint4 v = *reinterpret_cast<const uint4*>(recs + i0 + k);
            r[k].meta = v.x; r[k].c01 = v.y; r[k].c23 = v.z;
        }
        mine = agg_then(mine, agg_of(r[k]));
    }
    // inclusive scan over the warp, then over the warps
    Agg inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Agg up = agg_shfl_up(inc, d);
        if ((int)lane >= d) inc = agg_then(up, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    Agg before = agg_identity();                   // everything in the block before this warp
    Agg total = agg_identity();
#pragma unroll
    for (int w = 0; w < kFoldThreads / 32; ++w) {
        if (w == (int)warp) before = total;
        total = agg_then(total, s_warp[w]);
    }
    if (!APPLY) {
        if (threadIdx.x == 0) {
            BlockAgg b;
            b.map = total.map;
            b.c[0] = total.c[0]; b.c[1] = total.c[1]; b.c[2] = total.c[2]; b.c[3] = total.c[3];
            b.entry = 0; b.base = 0;
            blocks[blockIdx.x] = b;
        }
        return;
    }
    // exclusive prefix of this thread inside the block
    Agg ex = agg_shfl_up(inc, 1);
    if (lane == 0) ex = agg_identity();
    ex = agg_then(before, ex);
    const uint32_t b_entry = blocks[blockIdx.x].entry;
    uint32_t state = map_get(ex.map, b_entry);
    unsigned long long cnt = blocks[blockIdx.x].base + pick4(ex.c[0], ex.c[1], ex.c[2], ex.c[3], b_entry);
#pragma unroll
    for (int k = 0; k < kFoldPerThread; ++k) {
        if (i0 + k < n_spans) {
            span_entry[i0 + k] = (uint8_t)state;
            span_base[i0 + k] = cnt;
            cnt += rec_count(r[k], state);
            state = map_get(r[k].meta & 0xFFu, state);
        }
    }
}

// K7b-2: one warp chains the block aggregates, 32 at a time
__global__ void mpeg_block_chain(BlockAgg* __restrict__ blocks, unsigned long long n_blocks, ScanCtl* __restrict__ ctl,
                                 uint32_t entry) {
    const uint32_t lane = threadIdx.x;
    uint32_t state = entry;
    unsigned long long cnt = 0;
    for (unsigned long long b0 = 0; b0 < n_blocks; b0 += 32) {
        const unsigned long long b = b0 + lane;
        Agg mine = agg_identity();
        if (b < n_blocks) {
            mine.map = blocks[b].map;
            mine.c[0] = blocks[b].c[0]; mine.c[1] = blocks[b].c[1]; mine.c[2] = blocks[b].c[2]; mine.c[3] = blocks[b].c[3];
        }
        Agg inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Agg up = agg_shfl_up(inc, d);
            if ((int)lane >= d) inc = agg_then(up, inc);
        }
        Agg ex = agg_shfl_up(inc, 1);
        if (lane == 0) ex = agg_identity();
        if (b < n_blocks) {
            blocks[b].entry = map_get(ex.map, state);
            blocks[b].base = cnt + pick4(ex.c[0], ex.c[1], ex.c[2], ex.c[3], state);
        }
        const uint32_t lmap = __shfl_sync(0xFFFFFFFFu, inc.map, 31);
        const uint32_t l0 = __shfl_sync(0xFFFFFFFFu, inc.c[0], 31), l1 = __shfl_sync(0xFFFFFFFFu, inc.c[1], 31);
        const uint32_t l2 = __shfl_sync(0xFFFFFFFFu, inc.c[2], 31), l3 = __shfl_sync(0xFFFFFFFFu, inc.c[3], 31);
        cnt += pick4(l0, l1, l2, l3, state);
        state = map_get(lmap, state);
    }
    if (lane == 0) {
        ctl->total = cnt;
        ctl->count[entry] = cnt;
        ctl->exit_state[entry] = state;
    }
}

// ---------------------------------------------------------------- K7c: compact
// Moves every span's list to its final, position-ordered place: one warp per span, no shared memory, so the SMs
// hold enough warps to hide the dependent loads (record -> base -> list).  Spans whose list is not valid for their
// true entry state (head-sensitive with entry != 0) or did not fit their slot go to a redo list.
constexpr int kCompactThreads = 256;
__global__ void __launch_bounds__(kCompactThreads)
mpeg_compact(unsigned long long n_spans, const SpanRec* __restrict__ recs, const uint8_t* __restrict__ span_entry,
             const unsigned long long* __restrict__ span_base, const uint16_t* __restrict__ t_off,
             const uint32_t* __restrict__ t_hdr, unsigned long long* __restrict__ out_pos, uint32_t* __restrict__ out_hdr,
             unsigned long long cap, unsigned long long pos_offset, uint32_t* __restrict__ redo_list,
             uint32_t* __restrict__ redo_count) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * (kCompactThreads / 32);
    for (unsigned long long span = (unsigned long long)blockIdx.x * (kCompactThreads / 32) + warp; span < n_spans; span += n_warps) {
        const uint4 v = *reinterpret_cast<const uint4*>(recs + span);
        SpanRec r;
        r.meta = v.x; r.c01 = v.y; r.c23 = v.z; r.pad = 0;
        const uint32_t entry = span_entry[span];
        const unsigned long long base = span_base[span];
        const uint32_t count = rec_count(r, entry);
        const bool sens = (r.meta & 0x100u) != 0;
        if ((!sens || entry == 0) && count <= (uint32_t)kCandCap) {
            const unsigned long long span0 = pos_offset + span * (unsigned long long)kSpanBytes;
            const uint16_t* g_off = t_off + span * (unsigned long long)kCandCap;
            const uint32_t* g_hdr = t_hdr + span * (unsigned long long)kCandCap;
            for (uint32_t k = lane; k < count; k += 32) {
                const unsigned long long gi = base + k;
                if (gi < cap) {
                    out_pos[gi] = span0 + g_off[k];
                    out_hdr[gi] = g_hdr[k];
                }
            }
        } else if (lane == 0) {
            redo_list[atomicAdd(redo_count, 1u)] = (uint32_t)span;
        }
    }
}

// K7d: the spans of the redo list are walked again with their true entry state and emit directly
__global__ void __launch_bounds__(kScanThreads, kCtasPerSm)
mpeg_redo(const uint8_t* __restrict__ bytes, unsigned long long n, ScanCtl* __restrict__ ctl, const uint8_t* __restrict__ span_entry,
          const unsigned long long* __restrict__ span_base, const uint32_t* __restrict__ redo_list,
          const uint32_t* __restrict__ redo_count, unsigned long long* __restrict__ out_pos, uint32_t* __restrict__ out_hdr,
          unsigned long long cap, unsigned long long pos_offset) {
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ring = smem_addr(smem_dyn) + warp * (uint32_t)(kStages * kStageBytes);
    const uint32_t n_redo = *redo_count, n_warps = gridDim.x * (kScanThreads / 32);
    for (uint32_t i = blockIdx.x * (kScanThreads / 32) + warp; i < n_redo; i += n_warps) {
        const unsigned long long span = redo_list[i];
        walk_tile(kModeEmit, bytes, n, span * (unsigned long long)kSpanBytes, span_entry[span], lane, ring, nullptr, nullptr, out_pos,
                  out_hdr, cap, span_base[span], pos_offset, ctl);
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
// match against the reference header (mpeg.rs:194-204): version, layer, sample rate, channel mode and protection must be
// equal; sr = base(version) * factor(ff), so with equal versions equal sr <=> equal ff.  As header bits: kRefMask below.
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

// Two levels of aggregation: one add per distinct key per warp (match.any), and those adds land in a small
// block-local table first — the dominant header of a real stream is shared by most candidates, and one global atomic
// per warp on ONE address serialised in L2 (1.14 ms for C5's 49 M candidates; the 197 MB of headers stream in 40 us).
// A key that finds both of its slots taken by other keys goes to the global histogram directly.
constexpr uint32_t kHotSlots = 512;
constexpr int kHistPerThread = 4;               // headers per thread and round, one 128-bit load: with one 4-byte load per
                                                // thread and round the pass was latency-bound at 1.3 TB/s (155 us on C5)
__global__ void __launch_bounds__(256)
mpeg_hist(const uint32_t* __restrict__ hdr, unsigned long long n, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_key[kHotSlots], s_cnt[kHotSlots];
    for (uint32_t i = threadIdx.x; i < kHotSlots; i += blockDim.x) { s_key[i] = 0xFFFFFFFFu; s_cnt[i] = 0u; }
    __syncthreads();
    const bool aligned = ((unsigned long long)hdr & 15ull) == 0ull;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * kHistPerThread;
    const unsigned long long rounds = (n + stride - 1) / stride;
    unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * kHistPerThread;
    for (unsigned long long r = 0; r < rounds; ++r, i += stride) {
        uint32_t h4[kHistPerThread];
        if (aligned && i + kHistPerThread <= n) {
            const uint4 v = *reinterpret_cast<const uint4*>(hdr + i);
            h4[0] = v.x; h4[1] = v.y; h4[2] = v.z; h4[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < kHistPerThread; ++k) h4[k] = i + k < n ? hdr[i + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < kHistPerThread; ++k) {
            // only a header that parses can win the vote (mpeg_pick_ref): the others — two thirds of the false syncs inside
            // payload bytes, whose scattered global atomics were what this pass was bound by — are not counted at all
            HdrInfo hi_;
            const bool on = i + k < n && parse_header_bits(0xFFE00000u | h4[k], hi_);
            const uint32_t key = on ? (h4[k] & (kHdrBins - 1)) : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
            if (on && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
                const uint32_t c = (uint32_t)__popc(peers);
                uint32_t slot = (key * 2654435761u) >> 23;                      // 9 bits
                bool done = false;
#pragma unroll
                for (int probe = 0; probe < 2 && !done; ++probe, slot ^= 1u) {
                    const uint32_t prev = atomicCAS(&s_key[slot], 0xFFFFFFFFu, key);
                    if (prev == 0xFFFFFFFFu || prev == key) { atomicAdd(&s_cnt[slot], c); done = true; }
                }
                if (!done) atomicAdd(hist + key, c);
            }
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < kHotSlots; k += blockDim.x)
        if (s_cnt[k]) atomicAdd(hist + s_key[k], s_cnt[k]);
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

// per-candidate validity against the reference header = parse_header_bits && match (above) && frame_len_bits, through a
// table: a candidate that matches the reference header (= equality of the version /
// protection, layer, sample-rate and channel-mode bits: mask 0x00170CC0; tests/test_mpeg_lut_model.py) can only differ from it in the bitrate index
// and the padding bit, so compute_frame_len's f64 divisions are done 28 times per block instead of once per candidate.
// Entry: bit 31 valid, payload << 4, skip.
constexpr uint32_t kRefMask = 0x00170CC0u;
__device__ __forceinline__ void build_len_lut(uint32_t ref_header, uint32_t* s_lut) {
    if (threadIdx.x < 32) {
        HdrInfo o;
        uint32_t val = 0;
        const uint32_t e = threadIdx.x >> 1;
        if (parse_header_bits(ref_header, o) && e >= 1 && e <= 14) {
            o.eeee = e;
            o.padded = threadIdx.x & 1u;
            uint32_t pl, sk;
            if (frame_len_bits(o, pl, sk)) val = 0x80000000u | (pl << 4) | sk;
        }
        s_lut[threadIdx.x] = val;
    }
    __syncthreads();
}
__device__ __forceinline__ bool cand_valid_lut(uint32_t h, uint32_t ref_header, const uint32_t* s_lut, uint32_t& payload, uint32_t& skip) {
    if ((h ^ ref_header) & kRefMask) return false;
    const uint32_t val = s_lut[((h >> 12) & 0xFu) * 2u + ((h >> 9) & 1u)];
    payload = (val >> 4) & 0x7FFFFFFu;
    skip = val & 0xFu;
    return (val >> 31) != 0;
}

// first file position of every valid header value (the duplicate-first quirk, mpeg.rs:39); positions rather than
// candidate indices so that the table can be min-reduced across GPUs
__global__ void __launch_bounds__(256)
mpeg_first_pos(const unsigned long long* __restrict__ pos, const uint32_t* __restrict__ hdr, unsigned long long n,
               uint32_t ref_header, unsigned long long* __restrict__ first) {
    // Like mpeg_hist: the dominant header is shared by most candidates, and even a look-before-atomicMin on ONE global
    // word is a hot L2 line for every warp (0.58 ms on C5).  Minima are taken per warp (match.any; candidates are in
    // file order, so the lowest lane of a key holds its smallest position — any other lane that undercuts it adds its
    // own), then per block in a small shared table, and reach the global table once per block and key.
    __shared__ uint32_t s_lut[32];
    __shared__ uint32_t s_key[kHotSlots];
    __shared__ unsigned long long s_min[kHotSlots];
    for (uint32_t i = threadIdx.x; i < kHotSlots; i += blockDim.x) { s_key[i] = 0xFFFFFFFFu; s_min[i] = ~0ull; }
    build_len_lut(ref_header, s_lut);                            // (ends with __syncthreads)
    auto global_min = [&](uint32_t key, unsigned long long p) {
        unsigned long long* slot = first + key;
        if (*reinterpret_cast<volatile unsigned long long*>(slot) > p) atomicMin(slot, p);   // the table only ever decreases
    };
    // four consecutive candidates per thread and round, loaded with three 128-bit loads (as one dependent 4-byte + 8-byte
    // load per thread and round the pass ran at 2.9 TB/s: 207 us on C5).  In every one of the four match rounds the
    // candidates still ascend with the lane, so the lowest lane of a key holds the key's smallest position of the round.
    const bool aligned = (((unsigned long long)pos | (unsigned long long)hdr) & 15ull) == 0ull;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * kHistPerThread;
    const unsigned long long rounds = (n + stride - 1) / stride;
    unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * kHistPerThread;
    const uint32_t lane = threadIdx.x & 31;
    for (unsigned long long r = 0; r < rounds; ++r, i += stride) {
        uint32_t h4[kHistPerThread];
        unsigned long long p4[kHistPerThread];
        if (aligned && i + kHistPerThread <= n) {
            const uint4 v = *reinterpret_cast<const uint4*>(hdr + i);
            h4[0] = v.x; h4[1] = v.y; h4[2] = v.z; h4[3] = v.w;
            const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(pos + i), b2 = *reinterpret_cast<const ulonglong2*>(pos + i + 2);
            p4[0] = a.x; p4[1] = a.y; p4[2] = b2.x; p4[3] = b2.y;
        } else {
#pragma unroll
            for (int k = 0; k < kHistPerThread; ++k) {
                h4[k] = i + k < n ? hdr[i + k] : 0u;                    // 0 has no sync bits: never valid
                p4[k] = i + k < n ? pos[i + k] : ~0ull;
            }
        }
#pragma unroll
        for (int k = 0; k < kHistPerThread; ++k) {
            uint32_t key = 0xFFFFFFFFu;
            unsigned long long p = ~0ull;
            uint32_t pl, sk;
            if (i + k < n && cand_valid_lut(h4[k], ref_header, s_lut, pl, sk)) { key = h4[k] & (kHdrBins - 1); p = p4[k]; }
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
            const int leader = __ffs(peers) - 1;
            const unsigned long long p_lead = __shfl_sync(0xFFFFFFFFu, p, leader);
            if (key != 0xFFFFFFFFu && ((int)lane == leader || p < p_lead)) {
                uint32_t slot = (key * 2654435761u) >> 23;
                bool done = false;
#pragma unroll
                for (int probe = 0; probe < 2 && !done; ++probe, slot ^= 1u) {
                    const uint32_t prev = atomicCAS(&s_key[slot], 0xFFFFFFFFu, key);
                    if (prev == 0xFFFFFFFFu || prev == key) { atomicMin(&s_min[slot], p); done = true; }
                }
                if (!done) global_min(key, p);
            }
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < kHotSlots; k += blockDim.x)
        if (s_key[k] != 0xFFFFFFFFu) global_min(s_key[k], s_min[k]);
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
    __shared__ uint32_t s_lut[32];
    build_len_lut(ref_header, s_lut);
    const unsigned long long i0 = (unsigned long long)blockIdx.x * kClsBlock + (unsigned long long)threadIdx.x * kClsPerThread;
    uint32_t cnt[kClsPerThread];
    uint32_t mine = 0;
    // the thread's eight candidates as four 128-bit loads up front (headers: 32 B, positions: 64 B, both aligned because
    // i0 is a multiple of 8): as one dependent scalar load per candidate the two passes ran at half the HBM rate
    uint32_t h8[kClsPerThread];
    unsigned long long p8[kClsPerThread];
    const bool aligned = (((unsigned long long)pos | (unsigned long long)hdr) & 15ull) == 0ull;      // caller-provided device pointers
    if (aligned && i0 + kClsPerThread <= n) {
        const uint4* hv = reinterpret_cast<const uint4*>(hdr + i0);
        const uint4 ha = hv[0], hb = hv[1];
        h8[0] = ha.x; h8[1] = ha.y; h8[2] = ha.z; h8[3] = ha.w; h8[4] = hb.x; h8[5] = hb.y; h8[6] = hb.z; h8[7] = hb.w;
        const ulonglong2* pv = reinterpret_cast<const ulonglong2*>(pos + i0);
#pragma unroll
        for (int k = 0; k < kClsPerThread / 2; ++k) { const ulonglong2 t = pv[k]; p8[2 * k] = t.x; p8[2 * k + 1] = t.y; }
    } else {
#pragma unroll
        for (int k = 0; k < kClsPerThread; ++k) {
            const bool in = i0 + k < n;
            h8[k] = in ? hdr[i0 + k] : 0u;                       // 0 has no sync bits: never valid
            p8[k] = in ? pos[i0 + k] : 0ull;
        }
    }
#pragma unroll
    for (int k = 0; k < kClsPerThread; ++k) {
        cnt[k] = 0;
        const uint32_t h = h8[k];
        uint32_t pl, sk;
        if (i0 + k < n && cand_valid_lut(h, ref_header, s_lut, pl, sk)) {
            const unsigned long long p = p8[k];
            cnt[k] = 1 + ((compat && first[h & (kHdrBins - 1)] == p) ? 1u : 0u);
            if (p + sk + pl > file_len) atomicExch(err, 1u);            // mpeg.rs:95-97 indexes past EOF
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
            if (o < cap) out[o] = p8[k];
            o += 1;
        }
    }
}

// One pass for the duplicate-first table AND the per-block counts of the single-GPU index (blast_mpeg_index_dev with
// reference_compat): the blocks are mpeg_classify's (kClsBlock consecutive candidates, eight per thread, 128-bit loads),
// so the count pass of mpeg_classify is not needed — its counts are the valid candidates of the block (written here) plus
// one per header value whose FIRST position lies in the block (mpeg_dup_blocks, once the table is complete).
__global__ void __launch_bounds__(kClsThreads)
mpeg_first_count(const unsigned long long* __restrict__ pos, const uint32_t* __restrict__ hdr, unsigned long long n,
                 uint32_t ref_header, unsigned long long file_len, unsigned long long* __restrict__ first,
                 unsigned long long* __restrict__ block_counts, uint32_t* __restrict__ err) {
    __shared__ uint32_t s_lut[32];
    __shared__ uint32_t s_key[kHotSlots];
    __shared__ unsigned long long s_min[kHotSlots];
    __shared__ uint32_t s_valid;
    for (uint32_t i = threadIdx.x; i < kHotSlots; i += blockDim.x) { s_key[i] = 0xFFFFFFFFu; s_min[i] = ~0ull; }
    if (threadIdx.x == 0) s_valid = 0u;
    build_len_lut(ref_header, s_lut);                            // (ends with __syncthreads)
    auto global_min = [&](uint32_t key, unsigned long long p) {
        unsigned long long* slot = first + key;
        if (*reinterpret_cast<volatile unsigned long long*>(slot) > p) atomicMin(slot, p);   // the table only ever decreases
    };
    const unsigned long long i0 = (unsigned long long)blockIdx.x * kClsBlock + (unsigned long long)threadIdx.x * kClsPerThread;
    uint32_t h8[kClsPerThread];
    unsigned long long p8[kClsPerThread];
    const bool aligned = (((unsigned long long)pos | (unsigned long long)hdr) & 15ull) == 0ull;
    if (aligned && i0 + kClsPerThread <= n) {
        const uint4* hv = reinterpret_cast<const uint4*>(hdr + i0);
        const uint4 ha = hv[0], hb = hv[1];
        h8[0] = ha.x; h8[1] = ha.y; h8[2] = ha.z; h8[3] = ha.w; h8[4] = hb.x; h8[5] = hb.y; h8[6] = hb.z; h8[7] = hb.w;
        const ulonglong2* pv = reinterpret_cast<const ulonglong2*>(pos + i0);
#pragma unroll
        for (int k = 0; k < kClsPerThread / 2; ++k) { const ulonglong2 t = pv[k]; p8[2 * k] = t.x; p8[2 * k + 1] = t.y; }
    } else {
#pragma unroll
        for (int k = 0; k < kClsPerThread; ++k) {
            const bool in = i0 + k < n;
            h8[k] = in ? hdr[i0 + k] : 0u;                       // 0 has no sync bits: never valid
            p8[k] = in ? pos[i0 + k] : ~0ull;
        }
    }
    const uint32_t lane = threadIdx.x & 31;
    uint32_t mine = 0;
    // a thread's eight candidates ascend, and so do the k-th candidates of the lanes: in each of the eight match rounds the
    // lowest lane of a key holds the key's smallest position of the round
#pragma unroll
    for (int k = 0; k < kClsPerThread; ++k) {
        uint32_t key = 0xFFFFFFFFu;
        unsigned long long p = ~0ull;
        uint32_t pl, sk;
        if (i0 + k < n && cand_valid_lut(h8[k], ref_header, s_lut, pl, sk)) {
            key = h8[k] & (kHdrBins - 1);
            p = p8[k];
            mine += 1;
            if (p + sk + pl > file_len) atomicExch(err, 1u);            // mpeg.rs:95-97 indexes past EOF
        }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        const int leader = __ffs(peers) - 1;
        const unsigned long long p_lead = __shfl_sync(0xFFFFFFFFu, p, leader);
        if (key != 0xFFFFFFFFu && ((int)lane == leader || p < p_lead)) {
            uint32_t slot = (key * 2654435761u) >> 23;
            bool done = false;
#pragma unroll
            for (int probe = 0; probe < 2 && !done; ++probe, slot ^= 1u) {
                const uint32_t prev = atomicCAS(&s_key[slot], 0xFFFFFFFFu, key);
                if (prev == 0xFFFFFFFFu || prev == key) { atomicMin(&s_min[slot], p); done = true; }
            }
            if (!done) global_min(key, p);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
    if (lane == 0 && mine) atomicAdd(&s_valid, mine);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < kHotSlots; k += blockDim.x)
        if (s_key[k] != 0xFFFFFFFFu) global_min(s_key[k], s_min[k]);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = s_valid;
}

// every header value with a first position adds one output (the duplicate, mpeg.rs:39) to the block that holds the position:
// the largest block whose first candidate is not behind it
__global__ void mpeg_dup_blocks(const unsigned long long* __restrict__ first, const unsigned long long* __restrict__ pos,
                                unsigned long long n, unsigned long long n_blocks, unsigned long long* __restrict__ block_counts) {
    const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= kHdrBins) return;
    const unsigned long long p = first[key];
    if (p == ~0ull) return;
    unsigned long long lo = 0, hi = n_blocks;                    // invariant: pos[lo * kClsBlock] <= p < pos[hi * kClsBlock] (hi == n_blocks: +inf)
    while (hi - lo > 1) {
        const unsigned long long mid = (lo + hi) / 2;
        if (pos[mid * kClsBlock] <= p) lo = mid; else hi = mid;
    }
    atomicAdd(block_counts + lo, 1ull);
}

// exclusive scan of the block counts (one block of 1,024 threads; the array is small): per-thread sums, a two-level
// shuffle scan across the block (one thread walking the 1,024 partial sums took 15 of the kernel's 39 us), per-thread write
__global__ void mpeg_scan_blocks(const unsigned long long* __restrict__ counts, unsigned long long* __restrict__ base,
                                 unsigned long long nb, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_warp[32];
    const unsigned long long per = (nb + blockDim.x - 1) / blockDim.x;
    const unsigned long long a = min(nb, threadIdx.x * per), b = min(nb, a + per);
    unsigned long long sum = 0;
    for (unsigned long long i = a; i < b; ++i) sum += counts[i];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if ((int)lane >= d) inc += up;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t n_warps = (blockDim.x + 31) / 32;
        unsigned long long w = lane < n_warps ? s_warp[lane] : 0ull;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if ((int)lane >= d) w += up;
        }
        s_warp[lane] = w;                                   // inclusive over the warps
    }
    __syncthreads();
    unsigned long long run = (inc - sum) + (warp ? s_warp[warp - 1] : 0ull);
    if (threadIdx.x == blockDim.x - 1) *total = run + sum;
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

struct ScanBufs {
    unsigned long long n_spans = 0, n_blocks = 0;
    SpanRec* recs = nullptr;
    unsigned long long* span_base = nullptr;
    uint8_t* span_entry = nullptr;
    BlockAgg* blocks = nullptr;
    uint32_t* redo_list = nullptr;     // spans that must be re-walked (n_spans entries at most)
    uint32_t* redo_count = nullptr;
    ScanCtl* ctl = nullptr;
    uint32_t* t_hdr = nullptr;
    uint16_t* t_off = nullptr;
    unsigned grid = 0;
};

// span records, per-span (entry, base), block aggregates and the control block live in context scratch slot 0, the
// temp candidate lists in slot 1 (grow-only: no cudaMalloc / cudaFree on this path after the first call of a size)
int scan_buffers(blast_ctx* ctx, uint64_t own_len, ScanBufs& sb) {
    sb.n_spans = (own_len + kSpanBytes - 1) / kSpanBytes;
    sb.n_blocks = (sb.n_spans + kFoldBlock - 1) / kFoldBlock;
    const size_t rec_b = (sb.n_spans * sizeof(SpanRec) + 255) & ~255ull, base_b = (sb.n_spans * 8 + 255) & ~255ull;
    const size_t entry_b = (sb.n_spans + 255) & ~255ull, blk_b = (sb.n_blocks * sizeof(BlockAgg) + 255) & ~255ull;
    const size_t redo_b = (sb.n_spans * sizeof(uint32_t) + 255) & ~255ull;
    uint8_t* s0 = static_cast<uint8_t*>(blast::scratch(ctx, 0, rec_b + base_b + entry_b + blk_b + redo_b + 512));
    if (!s0) return BLAST_ERR_CUDA;
    sb.recs = reinterpret_cast<SpanRec*>(s0);
    sb.span_base = reinterpret_cast<unsigned long long*>(s0 + rec_b);
    sb.span_entry = s0 + rec_b + base_b;
    sb.blocks = reinterpret_cast<BlockAgg*>(s0 + rec_b + base_b + entry_b);
    sb.redo_list = reinterpret_cast<uint32_t*>(s0 + rec_b + base_b + entry_b + blk_b);
    sb.redo_count = reinterpret_cast<uint32_t*>(s0 + rec_b + base_b + entry_b + blk_b + redo_b);
    sb.ctl = reinterpret_cast<ScanCtl*>(s0 + rec_b + base_b + entry_b + blk_b + redo_b + 256);
    const size_t hdr_b = (sb.n_spans * kCandCap * sizeof(uint32_t) + 255) & ~255ull;
    uint8_t* s1 = static_cast<uint8_t*>(blast::scratch(ctx, 1, hdr_b + sb.n_spans * kCandCap * sizeof(uint16_t)));
    if (!s1) return BLAST_ERR_CUDA;
    sb.t_hdr = reinterpret_cast<uint32_t*>(s1);
    sb.t_off = reinterpret_cast<uint16_t*>(s1 + hdr_b);
    if (ctx->mpeg_ctas_per_sm == 0) {                      // function attributes are per device: cached per context
        int per_sm = 0;
        BLAST_CUDA_TRY(cudaFuncSetAttribute(mpeg_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmem));
        BLAST_CUDA_TRY(cudaFuncSetAttribute(mpeg_redo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmem));
        BLAST_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mpeg_walk, kScanThreads, kScanSmem));
        ctx->mpeg_ctas_per_sm = std::max(per_sm, 1);
    }
    const int per_sm = ctx->mpeg_ctas_per_sm;
    const unsigned long long want = (sb.n_spans + kWalkers - 1) / kWalkers;
    sb.grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)ctx->sm_count * std::max(per_sm, 1));
    return BLAST_OK;
}

// phase 1: walk the own spans (`readable` >= own_len bytes may be read: a following range's first bytes are real data)
// and fold the records into block aggregates
int scan_walk(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t readable, const ScanBufs& sb) {
    BLAST_CUDA_TRY(cudaMemsetAsync(sb.ctl, 0, sizeof(ScanCtl), ctx->stream));
    mpeg_walk<<<sb.grid, kScanThreads, kScanSmem, ctx->stream>>>(d_bytes, readable, sb.n_spans, sb.ctl, sb.recs, sb.t_off, sb.t_hdr);
    mpeg_span_fold<false><<<(unsigned)sb.n_blocks, kFoldThreads, 0, ctx->stream>>>(sb.recs, sb.n_spans, sb.blocks, sb.span_entry, sb.span_base);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    return BLAST_OK;
}

// phase 2: chain the blocks from a known entry state, give every span its (entry, base), move the lists out
int scan_emit(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t readable, const ScanBufs& sb, uint32_t entry, uint64_t pos_offset,
              uint64_t* d_pos, uint32_t* d_hdr, uint64_t cap, uint64_t* n_out) {
    ScanCtl* h_ctl = static_cast<ScanCtl*>(blast::mailbox(ctx));
    if (!h_ctl) return BLAST_ERR_CUDA;
    mpeg_block_chain<<<1, 32, 0, ctx->stream>>>(sb.blocks, sb.n_blocks, sb.ctl, entry);
    mpeg_span_fold<true><<<(unsigned)sb.n_blocks, kFoldThreads, 0, ctx->stream>>>(sb.recs, sb.n_spans, sb.blocks, sb.span_entry, sb.span_base);
    BLAST_CUDA_TRY(cudaMemsetAsync(sb.redo_count, 0, sizeof(uint32_t), ctx->stream));
    const unsigned cgrid = (unsigned)std::min<unsigned long long>((sb.n_spans + 7) / 8, (unsigned long long)ctx->sm_count * 8);
    mpeg_compact<<<cgrid, kCompactThreads, 0, ctx->stream>>>(sb.n_spans, sb.recs, sb.span_entry, sb.span_base, sb.t_off, sb.t_hdr,
                                                             reinterpret_cast<unsigned long long*>(d_pos), d_hdr, cap, pos_offset,
                                                             sb.redo_list, sb.redo_count);
    mpeg_redo<<<sb.grid, kScanThreads, kScanSmem, ctx->stream>>>(d_bytes, readable, sb.ctl, sb.span_entry, sb.span_base, sb.redo_list,
                                                                sb.redo_count, reinterpret_cast<unsigned long long*>(d_pos), d_hdr, cap,
                                                                pos_offset);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 4;
    BLAST_CUDA_TRY(cudaMemcpyAsync(h_ctl, sb.ctl, sizeof(ScanCtl), cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const ScanCtl h = *h_ctl;
    *n_out = h.total;
    if (h.panic) return blast::set_error(BLAST_ERR_REF_PANIC, "index out of bounds: the scan reaches a trailing 0xFF (mpeg.rs:20)");
    if (h.total > cap) return blast::set_error(BLAST_ERR_CAPACITY, "mpeg scan: %llu candidates, capacity %llu",
                                               (unsigned long long)h.total, (unsigned long long)cap);
    return BLAST_OK;
}

int run_scan(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, uint64_t* d_pos, uint32_t* d_hdr, uint64_t cap,
             uint64_t* n_out) {
    if (((uintptr_t)d_bytes & 15) != 0) return blast::set_error(BLAST_ERR_ARG, "mpeg scan: d_bytes must be 16-byte aligned");
    *n_out = 0;
    if (len == 0) return BLAST_OK;
    ScanBufs sb;
    if (int rc = scan_buffers(ctx, len, sb)) return rc;
    if (int rc = scan_walk(ctx, d_bytes, len, sb)) return rc;
    return scan_emit(ctx, d_bytes, len, sb, 0u, 0ull, d_pos, d_hdr, cap, n_out);
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

// first-position table + per-block counts in ONE pass over the candidates, then the duplicates per block, the scan of the
// counts and mpeg_classify's emit pass (blast_mpeg_classify_dev's count pass is not run)
static int index_compat_fused(blast_ctx* ctx, const uint64_t* d_pos, const uint32_t* d_hdr, uint64_t n, uint32_t ref_header,
                              uint64_t* d_first, uint64_t stream_len, uint64_t* d_offsets_out, uint64_t cap, uint64_t* n_offsets_out) {
    const unsigned long long n_blocks = (n + kClsBlock - 1) / kClsBlock;
    const size_t bc_b = (n_blocks * 8 + 255) & ~255ull;
    uint8_t* s7 = static_cast<uint8_t*>(blast::scratch(ctx, 7, 2 * bc_b + 256));
    unsigned long long* h_box = static_cast<unsigned long long*>(blast::mailbox(ctx));
    if (!s7 || !h_box) return BLAST_ERR_CUDA;
    h_box += 64;
    unsigned long long* d_bc = reinterpret_cast<unsigned long long*>(s7);
    unsigned long long* d_bb = reinterpret_cast<unsigned long long*>(s7 + bc_b);
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(s7 + 2 * bc_b);
    uint32_t* d_err = reinterpret_cast<uint32_t*>(d_total + 1);
    BLAST_CUDA_TRY(cudaMemsetAsync(d_total, 0, 16, ctx->stream));
    const unsigned long long* pos = reinterpret_cast<const unsigned long long*>(d_pos);
    unsigned long long* first = reinterpret_cast<unsigned long long*>(d_first);
    mpeg_first_count<<<(unsigned)n_blocks, kClsThreads, 0, ctx->stream>>>(pos, d_hdr, n, ref_header, stream_len, first, d_bc, d_err);
    mpeg_dup_blocks<<<kHdrBins / 256, 256, 0, ctx->stream>>>(first, pos, n, n_blocks, d_bc);
    mpeg_scan_blocks<<<1, 1024, 0, ctx->stream>>>(d_bc, d_bb, n_blocks, d_total);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 3;
    BLAST_CUDA_TRY(cudaMemcpyAsync(h_box, d_total, 16, cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const unsigned long long total = h_box[0];
    const uint32_t err = (uint32_t)h_box[1];
    *n_offsets_out = total;
    if (err)
        return blast::set_error(BLAST_ERR_REF_PANIC, "a frame payload extends past the end of the file (mpeg.rs:96 indexes out of bounds)");
    if (d_offsets_out) {
        if (total > cap) return blast::set_error(BLAST_ERR_CAPACITY, "mpeg index: %llu offsets, capacity %llu", total, (unsigned long long)cap);
        mpeg_classify<<<(unsigned)n_blocks, kClsThreads, 0, ctx->stream>>>(pos, d_hdr, n, ref_header, first, 1, stream_len, d_bc, d_bb, 1,
                                                                         reinterpret_cast<unsigned long long*>(d_offsets_out), cap, d_err);
        BLAST_CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return BLAST_OK;
}

int blast_mpeg_index_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t len, int reference_compat,
                         uint64_t* d_offsets_out, uint64_t cap, uint64_t* n_offsets_out, uint32_t* ref_header_out,
                         uint64_t* n_candidates_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(n_offsets_out && (d_bytes || len == 0), BLAST_ERR_ARG, "blast_mpeg_index_dev: null argument");
    *n_offsets_out = 0;
    if (n_candidates_out) *n_candidates_out = 0;
    // candidates: one pass with a guessed capacity (an MP3 stream has one sync per ~400 bytes, random bytes one
    // per 2,048); an exact second pass only if the guess was too small.  All buffers are context scratch
    // (grow-only): a second call of the same size allocates nothing.
    uint64_t n_cand = 0, guess = len / 128 + 4096;
    auto cand_buffers = [&](uint64_t count, uint64_t** pos, uint32_t** hdr) -> bool {
        *pos = static_cast<uint64_t*>(blast::scratch(ctx, 2, count * 8));
        *hdr = static_cast<uint32_t*>(blast::scratch(ctx, 3, count * 4));
        return *pos && *hdr;
    };
    uint64_t* d_pos = nullptr;
    uint32_t* d_hdr = nullptr;
    if (!cand_buffers(guess, &d_pos, &d_hdr)) return BLAST_ERR_CUDA;
    int rc = run_scan(ctx, d_bytes, len, d_pos, d_hdr, guess, &n_cand);
    if (rc == BLAST_ERR_CAPACITY) {
        if (!cand_buffers(n_cand, &d_pos, &d_hdr)) return BLAST_ERR_CUDA;
        uint64_t n2 = 0;
        rc = run_scan(ctx, d_bytes, len, d_pos, d_hdr, n_cand, &n2);
    }
    if (rc != BLAST_OK) return rc;
    if (n_candidates_out) *n_candidates_out = n_cand;
    if (n_cand == 0) return blast::set_error(BLAST_ERR_REF_PANIC, "no sync candidates: the reference indexes an empty list (mpeg.rs:64)");

    const size_t hist_b = kHdrBins * sizeof(uint32_t), first_b = (size_t)kHdrBins * 8;
    uint8_t* s4 = static_cast<uint8_t*>(blast::scratch(ctx, 4, hist_b + first_b));
    if (!s4) return BLAST_ERR_CUDA;
    uint32_t* d_hist = reinterpret_cast<uint32_t*>(s4);
    uint64_t* d_first = reinterpret_cast<uint64_t*>(s4 + hist_b);
    BLAST_CUDA_TRY(cudaMemsetAsync(d_hist, 0, hist_b, ctx->stream));
    if ((rc = blast_mpeg_hist_dev(ctx, d_hdr, n_cand, d_hist)) != BLAST_OK) return rc;
    uint32_t ref_header = 0;
    if ((rc = blast_mpeg_pick_ref_dev(ctx, d_hist, &ref_header)) != BLAST_OK) return rc;
    if (ref_header_out) *ref_header_out = ref_header;
    if (reference_compat) {
        BLAST_CUDA_TRY(cudaMemsetAsync(d_first, 0xFF, first_b, ctx->stream));
        return index_compat_fused(ctx, d_pos, d_hdr, n_cand, ref_header, d_first, len, d_offsets_out, cap, n_offsets_out);
    }
    return blast_mpeg_classify_dev(ctx, d_pos, d_hdr, n_cand, ref_header, nullptr, len, d_offsets_out, cap, n_offsets_out);
}

// ---- sharded scan (SURVEY §8 e): one contiguous byte range per GPU, one small exchange between the two phases
int blast_mpeg_shard_walk_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t own_len, uint64_t halo_len,
                              blast_mpeg_shard_agg* agg_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(agg_out && (d_bytes || own_len == 0), BLAST_ERR_ARG, "blast_mpeg_shard_walk_dev: null argument");
    if (((uintptr_t)d_bytes & 15) != 0) return blast::set_error(BLAST_ERR_ARG, "mpeg scan: d_bytes must be 16-byte aligned");
    if (halo_len != 0 && (own_len % kSpanBytes != 0 || halo_len < 16))
        return blast::set_error(BLAST_ERR_ARG, "a range that is followed by another one must be a multiple of %d bytes and carry >= 16 halo bytes", kSpanBytes);
    for (int s = 0; s < 4; ++s) { agg_out->exit_state[s] = (uint32_t)s; agg_out->count[s] = 0; }
    if (own_len == 0) return BLAST_OK;
    ScanBufs sb;
    if (int rc = scan_buffers(ctx, own_len, sb)) return rc;
    if (int rc = scan_walk(ctx, d_bytes, own_len + halo_len, sb)) return rc;
    for (uint32_t s = 0; s < 4; ++s) mpeg_block_chain<<<1, 32, 0, ctx->stream>>>(sb.blocks, sb.n_blocks, sb.ctl, s);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 4;
    ScanCtl* h_ctl = static_cast<ScanCtl*>(blast::mailbox(ctx));
    if (!h_ctl) return BLAST_ERR_CUDA;
    BLAST_CUDA_TRY(cudaMemcpyAsync(h_ctl, sb.ctl, sizeof(ScanCtl), cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < 4; ++s) { agg_out->exit_state[s] = h_ctl->exit_state[s]; agg_out->count[s] = h_ctl->count[s]; }
    return BLAST_OK;
}

int blast_mpeg_shard_emit_dev(blast_ctx* ctx, const uint8_t* d_bytes, uint64_t own_len, uint64_t halo_len, uint32_t entry_state,
                              uint64_t pos_offset, uint64_t* d_pos_out, uint32_t* d_hdr_out, uint64_t cap, uint64_t* n_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(n_out && (d_bytes || own_len == 0) && ((d_pos_out && d_hdr_out) || cap == 0), BLAST_ERR_ARG,
                  "blast_mpeg_shard_emit_dev: null argument");
    BLAST_REQUIRE(entry_state < 4, BLAST_ERR_ARG, "blast_mpeg_shard_emit_dev: entry_state must be 0..3");
    *n_out = 0;
    if (own_len == 0) return BLAST_OK;
    ScanBufs sb;
    if (int rc = scan_buffers(ctx, own_len, sb)) return rc;           // same sizes as the walk: the same scratch, untouched
    return scan_emit(ctx, d_bytes, own_len + halo_len, sb, entry_state, pos_offset, d_pos_out, d_hdr_out, cap, n_out);
}

// ---- the header vote and the frame filter as separate steps (blast_mpeg_index_dev = these four in sequence)
int blast_mpeg_hist_dev(blast_ctx* ctx, const uint32_t* d_hdr, uint64_t n, uint32_t* d_hist) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(d_hist && (d_hdr || n == 0), BLAST_ERR_ARG, "blast_mpeg_hist_dev: null argument");
    if (n == 0) return BLAST_OK;
    const unsigned g = (unsigned)std::min<unsigned long long>((n + 256 * kHistPerThread - 1) / (256 * kHistPerThread), (unsigned long long)ctx->sm_count * 16);
    mpeg_hist<<<g, 256, 0, ctx->stream>>>(d_hdr, n, d_hist);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

int blast_mpeg_pick_ref_dev(blast_ctx* ctx, const uint32_t* d_hist, uint32_t* ref_header_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(d_hist && ref_header_out, BLAST_ERR_ARG, "blast_mpeg_pick_ref_dev: null argument");
    unsigned long long* d_best = static_cast<unsigned long long*>(blast::scratch(ctx, 6, 256));
    unsigned long long* h_box = static_cast<unsigned long long*>(blast::mailbox(ctx));
    if (!d_best || !h_box) return BLAST_ERR_CUDA;
    h_box += 64;
    BLAST_CUDA_TRY(cudaMemsetAsync(d_best, 0, 8, ctx->stream));
    mpeg_pick_ref<<<kHdrBins / 256, 256, 0, ctx->stream>>>(d_hist, d_best);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    BLAST_CUDA_TRY(cudaMemcpyAsync(h_box, d_best, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_box[0] == 0) return blast::set_error(BLAST_ERR_REF_PANIC, "no parsable header: the reference indexes past its candidate list (mpeg.rs:64)");
    *ref_header_out = 0xFFE00000u | (kHdrBins - 1 - (uint32_t)(h_box[0] & (kHdrBins - 1)));
    return BLAST_OK;
}

int blast_mpeg_first_pos_dev(blast_ctx* ctx, const uint64_t* d_pos, const uint32_t* d_hdr, uint64_t n, uint32_t ref_header,
                             uint64_t* d_first) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(d_first && ((d_pos && d_hdr) || n == 0), BLAST_ERR_ARG, "blast_mpeg_first_pos_dev: null argument");
    if (n == 0) return BLAST_OK;
    const unsigned g = (unsigned)std::min<unsigned long long>((n + 256 * kHistPerThread - 1) / (256 * kHistPerThread), (unsigned long long)ctx->sm_count * 16);
    mpeg_first_pos<<<g, 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long*>(d_pos), d_hdr, n, ref_header,
                                               reinterpret_cast<unsigned long long*>(d_first));
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

int blast_mpeg_classify_dev(blast_ctx* ctx, const uint64_t* d_pos, const uint32_t* d_hdr, uint64_t n, uint32_t ref_header,
                            const uint64_t* d_first, uint64_t stream_len, uint64_t* d_offsets_out, uint64_t cap,
                            uint64_t* n_offsets_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(n_offsets_out && ((d_pos && d_hdr) || n == 0), BLAST_ERR_ARG, "blast_mpeg_classify_dev: null argument");
    *n_offsets_out = 0;
    if (n == 0) return BLAST_OK;
    const int compat = d_first != nullptr;
    const unsigned long long n_blocks = (n + kClsBlock - 1) / kClsBlock;
    const size_t bc_b = (n_blocks * 8 + 255) & ~255ull;
    uint8_t* s7 = static_cast<uint8_t*>(blast::scratch(ctx, 7, 2 * bc_b + 256));
    unsigned long long* h_box = static_cast<unsigned long long*>(blast::mailbox(ctx));
    if (!s7 || !h_box) return BLAST_ERR_CUDA;
    h_box += 64;
    unsigned long long* d_bc = reinterpret_cast<unsigned long long*>(s7);
    unsigned long long* d_bb = reinterpret_cast<unsigned long long*>(s7 + bc_b);
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(s7 + 2 * bc_b);
    uint32_t* d_err = reinterpret_cast<uint32_t*>(d_total + 1);
    BLAST_CUDA_TRY(cudaMemsetAsync(d_total, 0, 16, ctx->stream));
    const unsigned long long* pos = reinterpret_cast<const unsigned long long*>(d_pos);
    const unsigned long long* first = reinterpret_cast<const unsigned long long*>(d_first);
    mpeg_classify<<<(unsigned)n_blocks, kClsThreads, 0, ctx->stream>>>(pos, d_hdr, n, ref_header, first, compat, stream_len, d_bc, nullptr, 0,
                                                                     nullptr, 0, d_err);
    mpeg_scan_blocks<<<1, 1024, 0, ctx->stream>>>(d_bc, d_bb, n_blocks, d_total);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    BLAST_CUDA_TRY(cudaMemcpyAsync(h_box, d_total, 16, cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const unsigned long long total = h_box[0];
    const uint32_t err = (uint32_t)h_box[1];
    *n_offsets_out = total;
    if (err && compat)
        return blast::set_error(BLAST_ERR_REF_PANIC, "a frame payload extends past the end of the file (mpeg.rs:96 indexes out of bounds)");
    if (d_offsets_out) {
        if (total > cap) return blast::set_error(BLAST_ERR_CAPACITY, "mpeg index: %llu offsets, capacity %llu", total, (unsigned long long)cap);
        mpeg_classify<<<(unsigned)n_blocks, kClsThreads, 0, ctx->stream>>>(pos, d_hdr, n, ref_header, first, compat, stream_len, d_bc, d_bb, 1,
                                                                         reinterpret_cast<unsigned long long*>(d_offsets_out), cap, d_err);
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
