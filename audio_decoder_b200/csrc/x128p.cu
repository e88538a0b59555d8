// x128p.cu — K6 `x128p_streams`: xoroshiro128+ parameter streams with Lemire multiply-shift ranges.
//
// Replaces blast/src/audio_processing/blast_rand.rs:4-60 (X128P::{new,next_u64,next_i64_range}):
//   next_u64:  r = s0 + s1;  t = s1 ^ s0;  s0 = rotl(s0,55) ^ t ^ (t << 14);  s1 = rotl(t,36)   (55/14/36)
//   next_i64_range(lo,hi): lo + ((r as u128 * |hi-lo| as u128) >> 64)      (no rejection; hi<lo allowed)
//
// The reference has no jump().  The state transition is GF(2)-linear, state_n = T^n * state_0 for a
// 128x128 bit matrix T, so stream s of a batch starts at T^(s*stride) * base: the host squares T
// into the handful of matrices J^(2^b) (J = T^stride) and each GPU thread applies the ones selected
// by the bits of its stream index.  Every draw index then reproduces the CPU's sequential sequence.
//
// Fill kernel: one thread per stream (32 consecutive streams per warp), 64-bit integer ALU work
// (~25 32-bit ops per draw) against 8 or 16 bytes written per draw: HBM-write-bound when the
// streams are materialised, ALU-bound when only checksums are kept.  Draws are staged through a
// padded shared-memory tile and written out as 128-byte contiguous rows per stream.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "blast_internal.h"

namespace {

struct U128 {
    uint64_t lo, hi;      // lo = s0, hi = s1
};

// ---- host: GF(2) matrices as 128 columns
using Mat = std::vector<U128>;

inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

inline U128 step_state(U128 s) {
    uint64_t t = s.hi ^ s.lo;
    return U128{rotl64(s.lo, 55) ^ t ^ (t << 14), rotl64(t, 36)};
}

Mat transition() {
    Mat m(128);
    for (int i = 0; i < 128; ++i) {
        U128 e{i < 64 ? 1ull << i : 0, i >= 64 ? 1ull << (i - 64) : 0};
        m[i] = step_state(e);
    }
    return m;
}

inline U128 mat_vec(const Mat& m, U128 v) {
    U128 acc{0, 0};
    for (int i = 0; i < 64; ++i) {
        if ((v.lo >> i) & 1) { acc.lo ^= m[i].lo; acc.hi ^= m[i].hi; }
        if ((v.hi >> i) & 1) { acc.lo ^= m[64 + i].lo; acc.hi ^= m[64 + i].hi; }
    }
    return acc;
}

Mat mat_mul(const Mat& a, const Mat& b) {       // (a*b) v = a (b v)
    Mat c(128);
    for (int i = 0; i < 128; ++i) c[i] = mat_vec(a, b[i]);
    return c;
}

Mat mat_pow(Mat base, uint64_t n) {
    Mat r(128);
    for (int i = 0; i < 128; ++i) r[i] = U128{i < 64 ? 1ull << i : 0, i >= 64 ? 1ull << (i - 64) : 0};
    while (n) {
        if (n & 1) r = mat_mul(base, r);
        n >>= 1;
        if (n) base = mat_mul(base, base);
    }
    return r;
}

// ---- device
__device__ __forceinline__ uint64_t rotl_d(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

__device__ __forceinline__ uint64_t next_u64(uint64_t& s0, uint64_t& s1) {
    const uint64_t r = s0 + s1;
    const uint64_t t = s1 ^ s0;
    s0 = rotl_d(s0, 55) ^ t ^ (t << 14);
    s1 = rotl_d(t, 36);
    return r;
}

// d_mats: n_mats matrices of 128 columns (uint4 = {lo.lo32, lo.hi32, hi.lo32, hi.hi32})
__global__ void x128p_jump_states(const uint4* __restrict__ mats, int n_mats, uint64_t base_lo, uint64_t base_hi,
                                  uint64_t n_streams, U128* __restrict__ out) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    uint64_t lo = base_lo, hi = base_hi;
    for (int b = 0; b < n_mats; ++b) {
        if (!((s >> b) & 1)) continue;
        const uint4* __restrict__ m = mats + (size_t)b * 128;
        uint64_t alo = 0, ahi = 0;
#pragma unroll 4
        for (int i = 0; i < 64; ++i) {
            const uint4 c0 = __ldg(m + i), c1 = __ldg(m + 64 + i);
            const uint64_t m0 = 0 - ((lo >> i) & 1), m1 = 0 - ((hi >> i) & 1);
            alo ^= (((uint64_t)c0.y << 32) | c0.x) & m0;
            ahi ^= (((uint64_t)c0.w << 32) | c0.z) & m0;
            alo ^= (((uint64_t)c1.y << 32) | c1.x) & m1;
            ahi ^= (((uint64_t)c1.w << 32) | c1.z) & m1;
        }
        lo = alo;
        hi = ahi;
    }
    out[s] = U128{lo, hi};
}

// Sub-stream start states: sub k of stream s starts at J^k * state_s, J = T^(draws / split).  mats = J^(2^b).
__global__ void x128p_split_states(const uint4* __restrict__ mats, int n_mats, const U128* __restrict__ states,
                                   uint64_t n_streams, uint32_t split, U128* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_streams * split) return;
    const uint64_t s = t / split;
    const uint32_t k = (uint32_t)(t - s * split);
    uint64_t lo = states[s].lo, hi = states[s].hi;
    for (int b = 0; b < n_mats; ++b) {
        if (!((k >> b) & 1)) continue;
        const uint4* __restrict__ m = mats + (size_t)b * 128;
        uint64_t alo = 0, ahi = 0;
#pragma unroll 4
        for (int i = 0; i < 64; ++i) {
            const uint4 c0 = __ldg(m + i), c1 = __ldg(m + 64 + i);
            const uint64_t m0 = 0 - ((lo >> i) & 1), m1 = 0 - ((hi >> i) & 1);
            alo ^= (((uint64_t)c0.y << 32) | c0.x) & m0;
            ahi ^= (((uint64_t)c0.w << 32) | c0.z) & m0;
            alo ^= (((uint64_t)c1.y << 32) | c1.x) & m1;
            ahi ^= (((uint64_t)c1.w << 32) | c1.z) & m1;
        }
        lo = alo;
        hi = ahi;
    }
    out[t] = U128{lo, hi};
}

// after the fill: stream state = state of its last sub-stream; checks = xor / wrapping sum over the sub-streams
__global__ void x128p_merge_split(const U128* __restrict__ sub_states, const uint64_t* __restrict__ sub_checks,
                                  uint64_t n_streams, uint32_t split, U128* __restrict__ states,
                                  uint64_t* __restrict__ checks) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    states[s] = sub_states[s * split + split - 1];
    if (checks) {
        uint64_t xr = 0, sr = 0, xg = 0, sg = 0;
        for (uint32_t k = 0; k < split; ++k) {
            const uint64_t* c = sub_checks + (s * split + k) * 4;
            xr ^= c[0]; sr += c[1]; xg ^= c[2]; sg += c[3];
        }
        uint64_t* o = checks + s * 4;
        o[0] = xr; o[1] = sr; o[2] = xg; o[3] = sg;
    }
}

constexpr int kWarps = 8;
constexpr int kRound = 32;                 // draws staged per stream per round (one 256-byte row)
constexpr int kPitch = kRound + 1;         // u64 row pitch: conflict-free column writes

template <bool kRaw, bool kRanged, bool kChecks>
__global__ void __launch_bounds__(kWarps * 32)
x128p_streams(U128* __restrict__ states, uint64_t n_streams, uint64_t draws, int64_t lo, uint64_t range,
              uint64_t* __restrict__ raw, int64_t* __restrict__ ranged, uint64_t* __restrict__ checks) {
    extern __shared__ __align__(16) uint64_t tile_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* tile_raw = tile_all + (size_t)warp * 32 * kPitch * ((kRaw ? 1 : 0) + (kRanged ? 1 : 0));
    uint64_t* tile_rng = tile_raw + (kRaw ? 32 * kPitch : 0);
    const uint64_t stream0 = ((uint64_t)blockIdx.x * kWarps + warp) * 32;      // first stream of this warp
    if (stream0 >= n_streams) return;
    const uint64_t stream = stream0 + lane;
    const bool live = stream < n_streams;
    uint64_t s0 = 0, s1 = 0;
    if (live) { const U128 st = states[stream]; s0 = st.lo; s1 = st.hi; }
    uint64_t xr = 0, sr = 0, xg = 0, sg = 0;
    const uint32_t rows = (uint32_t)min((uint64_t)32, n_streams - stream0);
    // 16-byte stores need 16-byte aligned rows in memory: row start = (stream * draws + j0) * 8
    const bool vec_ok = (draws % 2 == 0) && (((uintptr_t)raw | (uintptr_t)ranged) & 15) == 0;
    const bool range32 = (range >> 32) == 0;
    const uint32_t rg = (uint32_t)range;

    for (uint64_t j0 = 0; j0 < draws; j0 += kRound) {
        const int nj = (int)min((uint64_t)kRound, draws - j0);
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
            const uint64_t r = next_u64(s0, s1);
            // blast_rand.rs:57-58: lower + ((r as u128 * range as u128) >> 64).  A range below 2^32 (every range the
            // reference draws from) needs two 32 x 32 -> 64 multiplies instead of the four of a full 64 x 64 high product.
            uint64_t m;
            if (range32) {
                const uint64_t t = (uint64_t)(uint32_t)r * rg;
                m = ((uint64_t)(uint32_t)(r >> 32) * rg + (t >> 32)) >> 32;
            } else {
                m = __umul64hi(r, range);
            }
            const uint64_t v = (uint64_t)lo + m;
            if (kChecks) { xr ^= r; sr += r; xg ^= v; sg += v; }
            if (kRaw) tile_raw[lane * kPitch + j] = r;
            if (kRanged) tile_rng[lane * kPitch + j] = v;
        }
        if (kRaw || kRanged) {
            __syncwarp();
            // row `q` of the tile = nj consecutive draws of stream stream0+q: 256 contiguous bytes in memory.
            // (Measured alternative: every lane storing its own draws straight from registers, 32 bytes at a time,
            // trusting the L2 to assemble lines — 19.0 ms instead of 8.2 ms for the raw C4 fill.)
            if (vec_ok && nj == kRound) {
                // 16 lanes x 16 bytes per row, two rows per store instruction
                const uint32_t jl = (lane & 15) * 2;
                for (uint32_t q2 = 0; q2 < rows; q2 += 2) {
                    const uint32_t q = q2 + (lane >> 4);
                    if (q < rows) {
                        const size_t base = (size_t)(stream0 + q) * draws + j0 + jl;
                        if (kRaw) {
                            const uint64_t a = tile_raw[q * kPitch + jl], b = tile_raw[q * kPitch + jl + 1];
                            blast::st_stream(reinterpret_cast<uint4*>(raw + base),
                                             make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32)));
                        }
                        if (kRanged) {
                            const uint64_t a = tile_rng[q * kPitch + jl], b = tile_rng[q * kPitch + jl + 1];
                            blast::st_stream(reinterpret_cast<uint4*>(ranged + base),
                                             make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32)));
                        }
                    }
                }
            } else {
                for (uint32_t q = 0; q < rows; ++q) {
                    if (lane < nj) {
                        const size_t base = (size_t)(stream0 + q) * draws + j0;
                        if (kRaw) raw[base + lane] = tile_raw[q * kPitch + lane];
                        if (kRanged) ranged[base + lane] = (int64_t)tile_rng[q * kPitch + lane];
                    }
                }
            }
            __syncwarp();
        }
    }
    if (live) {
        states[stream] = U128{s0, s1};
        if (kChecks) {
            uint64_t* c = checks + stream * 4;
            c[0] = xr; c[1] = sr; c[2] = xg; c[3] = sg;
        }
    }
}

}  // namespace

extern "C" {

// X128P::new (blast_rand.rs:10-24): two consecutive SplitMix64 outputs
void blast_x128p_seed(uint64_t seed, blast_x128p* out) {
    auto mix = [](uint64_t x) {
        x += 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        return x ^ (x >> 31);
    };
    out->s0 = mix(seed);
    out->s1 = mix(seed + 0x9E3779B97F4A7C15ull);
}

// host-side single-state jump (same matrices the device kernel consumes)
int blast_x128p_advance(const blast_x128p* in, uint64_t n_draws, blast_x128p* out) {
    BLAST_REQUIRE(in && out, BLAST_ERR_ARG, "blast_x128p_advance: null argument");
    const U128 r = mat_vec(mat_pow(transition(), n_draws), U128{in->s0, in->s1});
    out->s0 = r.lo;
    out->s1 = r.hi;
    return BLAST_OK;
}

int blast_x128p_jump_dev(blast_ctx* ctx, const blast_x128p* base, uint64_t stride, uint64_t n_streams,
                         blast_x128p* d_states_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(base && (d_states_out || n_streams == 0), BLAST_ERR_ARG, "blast_x128p_jump_dev: null argument");
    if (n_streams == 0) return BLAST_OK;
    int n_mats = 0;
    while (n_mats < 64 && ((n_streams - 1) >> n_mats) != 0) ++n_mats;
    std::vector<uint4> h((size_t)std::max(n_mats, 1) * 128);
    if (n_mats) {
        Mat j = mat_pow(transition(), stride);                  // J = T^stride
        for (int b = 0; b < n_mats; ++b) {
            for (int i = 0; i < 128; ++i)
                h[(size_t)b * 128 + i] = make_uint4((uint32_t)j[i].lo, (uint32_t)(j[i].lo >> 32), (uint32_t)j[i].hi,
                                                    (uint32_t)(j[i].hi >> 32));
            if (b + 1 < n_mats) j = mat_mul(j, j);              // J^(2^(b+1))
        }
    }
    uint4* d_mats = nullptr;
    BLAST_CUDA_TRY(cudaMalloc(&d_mats, h.size() * sizeof(uint4)));
    cudaError_t e = cudaMemcpyAsync(d_mats, h.data(), h.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((n_streams + 127) / 128);
        x128p_jump_states<<<blocks, 128, 0, ctx->stream>>>(d_mats, n_mats, base->s0, base->s1, n_streams,
                                                           reinterpret_cast<U128*>(d_states_out));
        e = cudaGetLastError();
        ctx->launches += 1;
    }
    cudaError_t s = cudaStreamSynchronize(ctx->stream);         // h / d_mats lifetime
    cudaFree(d_mats);
    if (e != cudaSuccess) return blast::set_error(BLAST_ERR_CUDA, "x128p jump failed: %s", cudaGetErrorString(e));
    if (s != cudaSuccess) return blast::set_error(BLAST_ERR_CUDA, "x128p jump failed: %s", cudaGetErrorString(s));
    return BLAST_OK;
}

namespace {

int launch_fill(blast_ctx* ctx, U128* st, uint64_t n_streams, uint64_t draws, int64_t lower, uint64_t range, uint64_t* d_raw,
                int64_t* d_ranged, uint64_t* d_checks) {
    const unsigned blocks = (unsigned)((n_streams + kWarps * 32 - 1) / (kWarps * 32));
    const int sel = (d_raw ? 1 : 0) | (d_ranged ? 2 : 0) | (d_checks ? 4 : 0);
    const size_t smem = (size_t)kWarps * 32 * kPitch * sizeof(uint64_t) * ((d_raw ? 1 : 0) + (d_ranged ? 1 : 0));
#define BLAST_FILL(R, G, K)                                                                                  \
    do {                                                                                                     \
        auto kern = x128p_streams<R, G, K>;                                                                  \
        if (smem > 48 * 1024)                                                                                \
            BLAST_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<blocks, kWarps * 32, smem, ctx->stream>>>(st, n_streams, draws, lower, range, d_raw, d_ranged, d_checks); \
    } while (0)
    switch (sel) {
        case 0: BLAST_FILL(false, false, false); break;     // advance only
        case 1: BLAST_FILL(true, false, false); break;
        case 2: BLAST_FILL(false, true, false); break;
        case 3: BLAST_FILL(true, true, false); break;
        case 4: BLAST_FILL(false, false, true); break;
        case 5: BLAST_FILL(true, false, true); break;
        case 6: BLAST_FILL(false, true, true); break;
        default: BLAST_FILL(true, true, true); break;
    }
#undef BLAST_FILL
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

}  // namespace

int blast_x128p_fill_dev(blast_ctx* ctx, blast_x128p* d_states, uint64_t n_streams, uint64_t draws_per_stream,
                         int64_t lower, int64_t upper, uint64_t* d_raw, int64_t* d_ranged, uint64_t* d_checks) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(d_states || n_streams == 0, BLAST_ERR_ARG, "blast_x128p_fill_dev: null states");
    if (n_streams == 0 || draws_per_stream == 0) return BLAST_OK;
    // blast_rand.rs:52-55: range = |upper - lower|
    const uint64_t range = upper > lower ? (uint64_t)upper - (uint64_t)lower : (uint64_t)lower - (uint64_t)upper;
    U128* st = reinterpret_cast<U128*>(d_states);

    // One thread per stream leaves most of the machine idle when there are few, long streams (C4: 65,536 streams are
    // 22 % of the resident threads, each a serial chain of 65,536 draws).  The generator is GF(2)-linear, so a stream
    // can be cut into `split` sub-streams that start at J^k * state (J = T^(draws/split)); draw j of sub-stream k is
    // draw k*draws/split + j of the stream and lands at the same address.
    uint32_t split = 1;
    const uint64_t want_threads = (uint64_t)ctx->sm_count * 1536;
    while (split < 64 && n_streams * split < want_threads && draws_per_stream % (2ull * split) == 0 &&
           draws_per_stream / (2ull * split) >= 1024)
        split *= 2;
    // Materialised draws are staged through shared memory: 24 (one output) or 8 (two) warps are resident per SM and the
    // warps of a launch come in rounds of that many, the last one partly filled.  One more doubling makes that last round
    // a smaller part of the whole (C4 on a B200: raw 5.65 -> 5.51 ms, raw + ranged 11.9 -> 11.2 ms; two more doublings
    // cost more in sub-stream set-up than they gain).
    if ((d_raw || d_ranged) && split > 1 && split < 64 && draws_per_stream % (2ull * split) == 0 && draws_per_stream / (2ull * split) >= 1024)
        split *= 2;
    static const uint32_t forced = getenv("BLAST_X128P_SPLIT") ? (uint32_t)atoi(getenv("BLAST_X128P_SPLIT")) : 0u;   // development
    if (forced) {
        split = 1;
        while (split < forced && split < 64 && draws_per_stream % (2ull * split) == 0 && draws_per_stream / (2ull * split) >= 1024) split *= 2;
    }
    if (split == 1) return launch_fill(ctx, st, n_streams, draws_per_stream, lower, range, d_raw, d_ranged, d_checks);

    const uint64_t sub = draws_per_stream / split, n_sub = n_streams * split;
    int n_mats = 0;
    while ((1u << n_mats) < split) ++n_mats;
    const size_t mats_b = (size_t)n_mats * 128 * sizeof(uint4), st_b = (n_sub * sizeof(U128) + 255) & ~255ull;
    uint8_t* sc = static_cast<uint8_t*>(blast::scratch(ctx, 5, mats_b + st_b + (d_checks ? n_sub * 32 : 0)));
    if (!sc) return BLAST_ERR_CUDA;
    uint4* d_mats = reinterpret_cast<uint4*>(sc);
    U128* d_sub = reinterpret_cast<U128*>(sc + mats_b);
    uint64_t* d_subchecks = d_checks ? reinterpret_cast<uint64_t*>(sc + mats_b + st_b) : nullptr;
    // J^(2^b): cached per (sub, n_mats) — squaring 128x128 bit matrices costs about a millisecond on the host
    if (ctx->x128p_split_sub != sub || ctx->x128p_split_mats != n_mats || ctx->x128p_split_ptr != d_mats) {
        std::vector<uint4> h((size_t)n_mats * 128, make_uint4(0, 0, 0, 0));
        Mat j = mat_pow(transition(), sub);
        for (int b = 0; b < n_mats; ++b) {
            for (int i = 0; i < 128; ++i)
                h[(size_t)b * 128 + i] = make_uint4((uint32_t)j[i].lo, (uint32_t)(j[i].lo >> 32), (uint32_t)j[i].hi,
                                                    (uint32_t)(j[i].hi >> 32));
            if (b + 1 < n_mats) j = mat_mul(j, j);
        }
        BLAST_CUDA_TRY(cudaMemcpyAsync(d_mats, h.data(), mats_b, cudaMemcpyHostToDevice, ctx->stream));
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));            // `h` is pageable and dies here
        ctx->x128p_split_sub = sub; ctx->x128p_split_mats = n_mats; ctx->x128p_split_ptr = d_mats;
    }
    x128p_split_states<<<(unsigned)((n_sub + 127) / 128), 128, 0, ctx->stream>>>(d_mats, n_mats, st, n_streams, split, d_sub);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    if (int rc = launch_fill(ctx, d_sub, n_sub, sub, lower, range, d_raw, d_ranged, d_subchecks)) return rc;
    x128p_merge_split<<<(unsigned)((n_streams + 127) / 128), 128, 0, ctx->stream>>>(d_sub, d_subchecks, n_streams, split, st, d_checks);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

int blast_x128p_fill(blast_ctx* ctx, uint64_t seed, uint64_t stride, uint64_t n_streams, uint64_t draws_per_stream,
                     int64_t lower, int64_t upper, uint64_t* raw_out, int64_t* ranged_out, uint64_t* checks_out) {
    if (int rc = blast::bind(ctx)) return rc;
    const uint64_t total = n_streams * draws_per_stream;
    if (total == 0) return BLAST_OK;
    blast_x128p base;
    blast_x128p_seed(seed, &base);
    blast_x128p* d_states = nullptr;
    uint64_t *d_raw = nullptr, *d_checks = nullptr;
    int64_t* d_rng = nullptr;
    int rc = BLAST_OK;
    auto done = [&](int code) {
        cudaStreamSynchronize(ctx->stream);
        if (d_states) cudaFree(d_states);
        if (d_raw) cudaFree(d_raw);
        if (d_rng) cudaFree(d_rng);
        if (d_checks) cudaFree(d_checks);
        return code;
    };
    if (cudaMalloc(&d_states, n_streams * sizeof(blast_x128p)) != cudaSuccess ||
        (raw_out && cudaMalloc(&d_raw, total * 8) != cudaSuccess) ||
        (ranged_out && cudaMalloc(&d_rng, total * 8) != cudaSuccess) ||
        (checks_out && cudaMalloc(&d_checks, n_streams * 32) != cudaSuccess))
        return done(blast::set_error(BLAST_ERR_CUDA, "blast_x128p_fill: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())));
    if ((rc = blast_x128p_jump_dev(ctx, &base, stride, n_streams, d_states)) != BLAST_OK) return done(rc);
    if ((rc = blast_x128p_fill_dev(ctx, d_states, n_streams, draws_per_stream, lower, upper, d_raw, d_rng, d_checks)) != BLAST_OK)
        return done(rc);
    cudaError_t e = cudaSuccess;
    if (raw_out) e = cudaMemcpyAsync(raw_out, d_raw, total * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && ranged_out) e = cudaMemcpyAsync(ranged_out, d_rng, total * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && checks_out) e = cudaMemcpyAsync(checks_out, d_checks, n_streams * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return done(blast::set_error(BLAST_ERR_CUDA, "blast_x128p_fill: %s", cudaGetErrorString(e)));
    return done(BLAST_OK);
}

}  // extern "C"
