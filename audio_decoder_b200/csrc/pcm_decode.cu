// pcm_decode.cu — K1 `pcm16_decode_batch` and K2 `pcm24_unpack_batch`.
//
// Replaces the per-pair sample loops of the reference:
//   blast/src/file_parsing/wav.rs:143-154   samples.push(i16::from_le_bytes([b[i], b[i+1]]))
//   blast/src/file_parsing/aiff.rs:159-170  samples.push(i16::from_be_bytes([b[i], b[i+1]]))
// Both ignore bits_per_sample: the payload is consumed as byte PAIRS.
//
// HBM-bound streaming kernel (roofline: 2 B read + 2 B written per i16 word).  One launch
// covers a whole ragged batch: a tile table maps each 16 KiB output tile to (job, tile index);
// a persistent grid (multiple of the SM count) strides over the tiles.  Every thread moves
// four independent 128-bit vectors per tile (all loads issued before the first store), the
// endian swap and the source misalignment (payloads start at +44 / +54 in their file image,
// and may start at odd addresses) are folded into one PRMT per 32-bit word.
#include <algorithm>
#include <cstring>
#include <vector>

#include "blast_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;                               // 4 x 16 B in flight per thread
constexpr int kTileVecs = kThreads * kVecPerThread;            // 1024 vectors
constexpr int kTileBytes = kTileVecs * 16;                     // 16 KiB of output per tile
constexpr int kCtasPerSm = 4;

struct JobDev {
    const uint8_t* src;    // first byte of the vector region (any alignment)
    uint4* dst;            // 16-byte aligned
    uint64_t n_vecs;       // full 16-byte output vectors
    // scalar fringe: `head` words before the vector region, `tail` words after it
    const uint8_t* head_src;
    int16_t* head_dst;
    const uint8_t* tail_src;
    int16_t* tail_dst;
    uint32_t head_words;
    uint32_t tail_words;
    uint32_t big_endian;
    uint32_t pad;
};

struct TileRef {
    uint32_t job;
    uint32_t tile;         // tile index inside the job
};

__device__ __forceinline__ int16_t pair_to_i16(const uint8_t* p, bool be) {
    uint32_t a = p[0], b = p[1];
    return (int16_t)(be ? ((a << 8) | b) : ((b << 8) | a));
}

// window of 8 consecutive 32-bit words (two aligned 16-byte vectors)
template <int W>
__device__ __forceinline__ uint4 splice(const uint4& lo, const uint4& hi, uint32_t sel) {
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint4 o;
    o.x = __byte_perm(w[W + 0], w[W + 1], sel);
    o.y = __byte_perm(w[W + 1], w[W + 2], sel);
    o.z = __byte_perm(w[W + 2], w[W + 3], sel);
    o.w = __byte_perm(w[W + 3], w[W + 4], sel);
    return o;
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
pcm16_decode_batch(const JobDev* __restrict__ jobs, const TileRef* __restrict__ tiles, uint32_t n_tiles) {
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const TileRef ref = tiles[t];
        const JobDev job = jobs[ref.job];
        const bool be = job.big_endian != 0;

        if (ref.tile == 0) {
            // scalar fringe of this job: at most 7 + 7 words
            if (threadIdx.x < job.head_words)
                job.head_dst[threadIdx.x] = pair_to_i16(job.head_src + 2 * threadIdx.x, be);
            else if (threadIdx.x >= 32 && threadIdx.x - 32 < job.tail_words)
                job.tail_dst[threadIdx.x - 32] = pair_to_i16(job.tail_src + 2 * (threadIdx.x - 32), be);
        }

        const uint64_t v0 = (uint64_t)ref.tile * kTileVecs;
        const uint64_t remaining = job.n_vecs - v0;
        const uint32_t nv = remaining < (uint64_t)kTileVecs ? (uint32_t)remaining : (uint32_t)kTileVecs;

        const uintptr_t s = (uintptr_t)job.src;
        const uint32_t mis = (uint32_t)(s & 15);
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(s - mis) + v0;
        uint4* __restrict__ dst = job.dst + v0;

        if (mis == 0) {
            // aligned fast path: copy (LE) or swap bytes in each half-word (BE)
            const uint32_t sel = be ? 0x2301u : 0x3210u;
            uint4 v[kVecPerThread];
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                uint32_t i = threadIdx.x + j * kThreads;
                if (i < nv) v[j] = blast::ld_stream(src + i);
            }
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                uint32_t i = threadIdx.x + j * kThreads;
                if (i < nv) {
                    uint4 o;
                    o.x = __byte_perm(v[j].x, 0, sel);
                    o.y = __byte_perm(v[j].y, 0, sel);
                    o.z = __byte_perm(v[j].z, 0, sel);
                    o.w = __byte_perm(v[j].w, 0, sel);
                    blast::st_stream(dst + i, o);
                }
            }
        } else {
            // misaligned source: output vector i needs bytes [mis, mis+16) of aligned vectors
            // (i, i+1).  One PRMT per word does the funnel shift and the endian swap.
            const uint32_t sh = mis & 3, wsel = mis >> 2;
            const uint32_t sel = be ? ((sh + 1) | (sh << 4) | ((sh + 3) << 8) | ((sh + 2) << 12))
                                    : (sh | ((sh + 1) << 4) | ((sh + 2) << 8) | ((sh + 3) << 12));
            uint4 lo[kVecPerThread], hi[kVecPerThread];
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                uint32_t i = threadIdx.x + j * kThreads;
                if (i < nv) {
                    lo[j] = blast::ld_cached(src + i);
                    hi[j] = blast::ld_cached(src + i + 1);
                }
            }
#pragma unroll
            for (int j = 0; j < kVecPerThread; ++j) {
                uint32_t i = threadIdx.x + j * kThreads;
                if (i < nv) {
                    uint4 o;
                    switch (wsel) {
                        case 0: o = splice<0>(lo[j], hi[j], sel); break;
                        case 1: o = splice<1>(lo[j], hi[j], sel); break;
                        case 2: o = splice<2>(lo[j], hi[j], sel); break;
                        default: o = splice<3>(lo[j], hi[j], sel); break;
                    }
                    blast::st_stream(dst + i, o);
                }
            }
        }
    }
}

// ---- K2: packed 24-bit -> i32 (sign-extended) or i16 (top 16 bits); extension, not in the reference (SURVEY §8 a4,
// north_star: "24-bit packed samples byte-swapped and unpacked in registers or shared-memory-staged tiles").
// A tile is 4,096 samples = 12,288 source bytes.  Global -> shared by 16-byte cp.async (LDGSTS, L1 bypassed) from the
// 16-byte boundary below the tile, two tiles deep, so the loads of tile i+1 are in flight while tile i is unpacked.  A
// thread unpacks 4 consecutive samples (12 bytes) per 1,024-sample slab: four conflict-free LDS.32 (the 12-byte thread
// stride is 3 banks), three funnel shifts that remove the tile's byte misalignment (0..3, the rest is whole words), and
// ONE PRMT per sample: the selector picks the three bytes in endian order and replicates the sign of the top byte
// (selector nibble | 8), i.e. byte swap and sign extension are the same instruction.  Stores are 16 bytes (i32) or
// 8 bytes (i16) per thread, coalesced.  Roofline: HBM, 3 B read + 4 B (or 2 B) written per sample.
struct Job24Dev {
    const uint8_t* src;
    void* dst;
    uint64_t n_samples;
    uint32_t big_endian;
    uint32_t out_kind;
    uint32_t tile0;            // index of the job's first tile in the batch (ascending: tiles find their job by bisection)
    uint32_t pad;
};

constexpr int k24Threads = 256;
constexpr int k24Slabs = 4;                                          // 1,024-sample slabs per tile
constexpr int k24SamplesPerTile = k24Threads * 4 * k24Slabs;         // 4,096 samples = 12,288 source bytes
constexpr int k24StageBytes = k24SamplesPerTile * 3 + 32;            // + the misalignment in front, + the last vector's slack
constexpr int k24CtasPerSm = 8;                                     // 8 x 24.6 KB of staging fit an SM; 256-thread blocks: 64 warps

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// the job of tile t: the last one whose first tile is <= t (the job table is a few KB and stays in L1 / L2; a tile table
// of the whole batch — a quarter of a million entries for C2 — would have to be built and uploaded by the host per call)
__device__ __forceinline__ uint32_t job_of_tile(const Job24Dev* __restrict__ jobs, uint32_t n_jobs, uint32_t t) {
    uint32_t lo = 0, hi = n_jobs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (jobs[mid].tile0 <= t) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(k24Threads)
pcm24_unpack_batch(const Job24Dev* __restrict__ jobs, uint32_t n_jobs, uint32_t n_tiles) {
    __shared__ __align__(16) uint8_t stage[2][k24StageBytes];
    __shared__ uint32_t s_job[2];
    auto issue = [&](uint32_t t, int buf) {
        const uint32_t j = job_of_tile(jobs, n_jobs, t);
        if (threadIdx.x == 0) s_job[buf] = j;                          // read after the barrier that follows the copies
        const Job24Dev job = jobs[j];
        const uint64_t s0 = (uint64_t)(t - job.tile0) * k24SamplesPerTile;
        const uint64_t left = job.n_samples - s0;
        const uint32_t ns = left < (uint64_t)k24SamplesPerTile ? (uint32_t)left : (uint32_t)k24SamplesPerTile;
        const uint8_t* base = job.src + s0 * 3;
        const uint32_t mis = (uint32_t)((uintptr_t)base & 15);
        const uint8_t* vsrc = base - mis;
        const uint32_t nvec = (ns * 3 + mis + 15) / 16;
        const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(stage[buf]);
        for (uint32_t i = threadIdx.x; i < nvec; i += k24Threads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + i * 16u), "l"(vsrc + (size_t)i * 16) : "memory");
    };
    uint32_t t = blockIdx.x;
    int buf = 0;
    if (t < n_tiles) issue(t, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (; t < n_tiles; t += gridDim.x, buf ^= 1) {
        if (t + gridDim.x < n_tiles) issue(t + gridDim.x, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();                                               // tile t is staged
        const Job24Dev job = jobs[s_job[buf]];
        const uint64_t s0 = (uint64_t)(t - job.tile0) * k24SamplesPerTile;
        const uint64_t left = job.n_samples - s0;
        const uint32_t ns = left < (uint64_t)k24SamplesPerTile ? (uint32_t)left : (uint32_t)k24SamplesPerTile;
        const uint32_t mis = (uint32_t)((uintptr_t)(job.src + s0 * 3) & 15);
        const uint32_t sh = (mis & 3u) * 8u;                           // byte misalignment inside a word: uniform over the tile
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage[buf]) + (mis & ~3u);
        const bool be = job.big_endian != 0;
        // selectors over the funnel-shifted words (y0, y1, y2) = the thread's 12 bytes b0..b11; nibble | 8 = sign of that byte
        const uint32_t w0 = be ? 0x8012u : 0xA210u, w1 = be ? 0xB345u : 0xD543u, w2 = be ? 0xA234u : 0xC432u, w3 = be ? 0x9123u : 0xB321u;
        const uint32_t h01 = be ? 0x3401u : 0x5421u, h23 = be ? 0x5623u : 0x7643u;
#pragma unroll
        for (int g = 0; g < k24Slabs; ++g) {
            const uint32_t q = (uint32_t)g * (k24Threads * 4) + threadIdx.x * 4;       // first sample of this thread in the slab
            if (q >= ns) break;
            const uint32_t a = sbase + q * 3u;                         // word aligned: q * 3 is a multiple of 4
            uint32_t x0, x1, x2, x3;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x0) : "r"(a));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(x1) : "r"(a));
            asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(x2) : "r"(a));
            asm volatile("ld.shared.u32 %0, [%1+12];" : "=r"(x3) : "r"(a));
            const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh);
            if (job.out_kind == 0) {
                const int4 o = make_int4((int32_t)prmt(y0, 0u, w0), (int32_t)prmt(y0, y1, w1), (int32_t)prmt(y1, y2, w2), (int32_t)prmt(y2, 0u, w3));
                int32_t* d = reinterpret_cast<int32_t*>(job.dst) + s0 + q;
                if (q + 4 <= ns && ((uintptr_t)d & 15) == 0) {
                    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(d), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                } else {
                    const int32_t v[4] = {o.x, o.y, o.z, o.w};
                    for (int k = 0; k < 4 && q + k < ns; ++k) d[k] = v[k];
                }
            } else {
                const uint32_t p0 = prmt(y0, y1, h01), p1 = prmt(y1, y2, h23);   // two top-16 samples per word
                int16_t* d = reinterpret_cast<int16_t*>(job.dst) + s0 + q;
                if (q + 4 <= ns && ((uintptr_t)d & 7) == 0) {
                    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(d), "r"(p0), "r"(p1) : "memory");
                } else {
                    const uint32_t v[2] = {p0, p1};
                    for (int k = 0; k < 4 && q + k < ns; ++k) d[k] = (int16_t)(v[k >> 1] >> ((k & 1) * 16));
                }
            }
        }
        __syncthreads();                                               // tile t consumed: its stage may be refilled
    }
}

}  // namespace

struct blast_pcm_plan {
    JobDev* d_jobs = nullptr;
    TileRef* d_tiles = nullptr;
    uint32_t n_jobs = 0;
    uint32_t n_tiles = 0;
    uint64_t words = 0;
    int grid = 0;
};

namespace {

// Translates ABI jobs into device job records + the tile table.  `hj` must hold n_jobs
// entries; tiles are appended to `ht`.
int fill_tables(const blast_pcm_job* jobs, uint32_t n_jobs, JobDev* hj, std::vector<TileRef>& ht, uint64_t* words_out) {
    uint64_t words = 0;
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const blast_pcm_job& in = jobs[j];
        if (in.n_words && (!in.d_src || !in.d_dst)) return blast::set_error(BLAST_ERR_ARG, "pcm job %u: null pointer", j);
        if ((uintptr_t)in.d_dst & 1) return blast::set_error(BLAST_ERR_ARG, "pcm job %u: d_dst is not 2-byte aligned", j);
        JobDev d{};
        d.big_endian = in.big_endian;
        // peel words until the destination is 16-byte aligned
        uint64_t head = (((16 - ((uintptr_t)in.d_dst & 15)) & 15) / 2);
        if (head > in.n_words) head = in.n_words;
        uint64_t body = in.n_words - head;
        uint64_t n_vecs = body / 8;
        uint64_t tail = body - n_vecs * 8;
        d.head_src = in.d_src;
        d.head_dst = in.d_dst;
        d.head_words = (uint32_t)head;
        d.src = in.d_src + 2 * head;
        d.dst = reinterpret_cast<uint4*>(in.d_dst + head);
        d.n_vecs = n_vecs;
        d.tail_src = d.src + 16 * n_vecs;
        d.tail_dst = in.d_dst + head + 8 * n_vecs;
        d.tail_words = (uint32_t)tail;
        hj[j] = d;
        uint64_t nt = (n_vecs + kTileVecs - 1) / kTileVecs;
        if (nt == 0 && (head || tail)) nt = 1;   // a fringe-only job still needs one tile
        if (nt > 0xFFFFFFFFull) return blast::set_error(BLAST_ERR_CAPACITY, "pcm job %u too large", j);
        for (uint64_t t = 0; t < nt; ++t) ht.push_back(TileRef{j, (uint32_t)t});
        words += in.n_words;
    }
    if (ht.size() > 0xFFFFFFFFull) return blast::set_error(BLAST_ERR_CAPACITY, "too many tiles in one batch");
    *words_out = words;
    return BLAST_OK;
}

int build_plan(blast_ctx* ctx, const blast_pcm_job* jobs, uint32_t n_jobs, cudaStream_t stream, blast_pcm_plan* plan) {
    std::vector<JobDev> hj(n_jobs);
    std::vector<TileRef> ht;
    uint64_t words = 0;
    if (int rc = fill_tables(jobs, n_jobs, hj.data(), ht, &words)) return rc;
    plan->n_jobs = n_jobs;
    plan->n_tiles = (uint32_t)ht.size();
    plan->words = words;
    if (plan->n_tiles == 0) return BLAST_OK;
    BLAST_CUDA_TRY(cudaMalloc(&plan->d_jobs, hj.size() * sizeof(JobDev)));
    BLAST_CUDA_TRY(cudaMalloc(&plan->d_tiles, ht.size() * sizeof(TileRef)));
    BLAST_CUDA_TRY(cudaMemcpyAsync(plan->d_jobs, hj.data(), hj.size() * sizeof(JobDev), cudaMemcpyHostToDevice, stream));
    BLAST_CUDA_TRY(cudaMemcpyAsync(plan->d_tiles, ht.data(), ht.size() * sizeof(TileRef), cudaMemcpyHostToDevice, stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(stream));   // hj / ht go out of scope
    int cap = ctx->sm_count * kCtasPerSm;
    plan->grid = (int)std::min<uint64_t>(plan->n_tiles, (uint64_t)cap);
    return BLAST_OK;
}

void free_plan(blast_pcm_plan* plan) {
    if (plan->d_jobs) cudaFree(plan->d_jobs);
    if (plan->d_tiles) cudaFree(plan->d_tiles);
    plan->d_jobs = nullptr;
    plan->d_tiles = nullptr;
}

int run_plan(blast_ctx* ctx, const blast_pcm_plan* plan, cudaStream_t stream) {
    if (plan->n_tiles == 0) return BLAST_OK;
    pcm16_decode_batch<<<plan->grid, kThreads, 0, stream>>>(plan->d_jobs, plan->d_tiles, plan->n_tiles);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

}  // namespace

extern "C" {

int blast_pcm_plan_create(blast_ctx* ctx, const blast_pcm_job* jobs, uint32_t n_jobs, blast_pcm_plan** out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(out != nullptr && (jobs != nullptr || n_jobs == 0), BLAST_ERR_ARG, "blast_pcm_plan_create: null argument");
    blast_pcm_plan* plan = new blast_pcm_plan();
    int rc = build_plan(ctx, jobs, n_jobs, ctx->stream, plan);
    if (rc != BLAST_OK) {
        free_plan(plan);
        delete plan;
        return rc;
    }
    *out = plan;
    return BLAST_OK;
}

int blast_pcm_plan_run_dev(blast_ctx* ctx, blast_pcm_plan* plan) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(plan != nullptr, BLAST_ERR_ARG, "blast_pcm_plan_run_dev: null plan");
    return run_plan(ctx, plan, ctx->stream);
}

void blast_pcm_plan_destroy(blast_ctx* ctx, blast_pcm_plan* plan) {
    if (!plan) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    free_plan(plan);
    delete plan;
}

uint64_t blast_pcm_plan_words(const blast_pcm_plan* plan) { return plan ? plan->words : 0; }

int blast_pcm_decode_dev(blast_ctx* ctx, const blast_pcm_job* jobs, uint32_t n_jobs) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(jobs != nullptr || n_jobs == 0, BLAST_ERR_ARG, "blast_pcm_decode_dev: null jobs");
    std::vector<JobDev> hj(n_jobs);
    std::vector<TileRef> ht;
    uint64_t words = 0;
    if (int rc = fill_tables(jobs, n_jobs, hj.data(), ht, &words)) return rc;
    if (ht.empty()) return BLAST_OK;
    // the tables go into the context's grow-only scratch (shared with the 24-bit unpack): no cudaMalloc / cudaFree and no
    // stream synchronisation once warm; the uploads come from pageable vectors, staged before cudaMemcpyAsync returns
    JobDev* d_jobs = static_cast<JobDev*>(blast::scratch(ctx, 8, hj.size() * sizeof(JobDev)));
    TileRef* d_tiles = static_cast<TileRef*>(blast::scratch(ctx, 9, ht.size() * sizeof(TileRef)));
    if (!d_jobs || !d_tiles) return BLAST_ERR_CUDA;
    ctx->tab8_ptr = nullptr;                                           // slot 8 no longer holds a 24-bit job table
    BLAST_CUDA_TRY(cudaMemcpyAsync(d_jobs, hj.data(), hj.size() * sizeof(JobDev), cudaMemcpyHostToDevice, ctx->stream));
    BLAST_CUDA_TRY(cudaMemcpyAsync(d_tiles, ht.data(), ht.size() * sizeof(TileRef), cudaMemcpyHostToDevice, ctx->stream));
    const int grid = (int)std::min<uint64_t>(ht.size(), (uint64_t)ctx->sm_count * kCtasPerSm);
    pcm16_decode_batch<<<grid, kThreads, 0, ctx->stream>>>(d_jobs, d_tiles, (uint32_t)ht.size());
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

// Host-buffer batch decode.  Every file's payload is cut into pieces of <= 8 MiB; pieces are
// grouped into chunks of <= 32 MiB; chunk c runs on pipeline lane c % kPipe (its own stream):
// table upload -> H2D of the pieces (16-byte aligned slots in the lane's staging slab) -> ONE
// decode launch -> D2H of the words.  Lanes overlap, so PCIe traffic in both directions and
// the kernels of neighbouring chunks run concurrently.  All staging memory is owned by the
// context (grow-only), nothing is allocated or freed inside the loop.
int blast_pcm_decode_batch(blast_ctx* ctx, uint32_t n, const uint8_t* const* files, const size_t* lens,
                           const blast_pcm_desc* descs, int16_t* const* host_out, int16_t* const* d_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(n == 0 || (files && lens && descs), BLAST_ERR_ARG, "blast_pcm_decode_batch: null argument");
    BLAST_REQUIRE(host_out || d_out, BLAST_ERR_ARG, "blast_pcm_decode_batch: no output requested");
    if (n == 0) return BLAST_OK;

    // the reference's bounds rule, checked for the whole batch before the GPU is touched
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t w = (descs[i].data_len + 1) / 2;
        if (w && (!files[i] || descs[i].data_off + 2 * w > (uint64_t)lens[i]))
            return blast::set_error(BLAST_ERR_UNEXPECTED_EOF, "file %u: UnexpectedEof in sample data", i);
    }

    constexpr uint64_t kPieceWords = 4ull << 20;            // 8 MiB of payload
    constexpr uint64_t kChunkBytes = 32ull << 20;
    constexpr uint32_t kChunkPieces = 4096;
    if (int rc = blast::ensure_pipe(ctx, kChunkBytes, kChunkPieces, kChunkBytes / kTileBytes + kChunkPieces,
                                    sizeof(JobDev), sizeof(TileRef)))
        return rc;

    struct Piece { uint32_t file; uint64_t word0, words; };
    std::vector<Piece> pieces;        // pieces of the chunk being assembled
    std::vector<blast_pcm_job> jobs;
    std::vector<TileRef> ht;
    uint64_t chunk_bytes = 0;
    uint32_t chunk_index = 0;
    int rc = BLAST_OK;

    // Host copies are coalesced: consecutive pieces whose source bytes are (almost) adjacent in host memory — the next
    // piece of the same file, or the next file when the two images are back to back (the bytes in between are then
    // that file's own header) — travel in ONE cudaMemcpyAsync, and so do results whose host and device destinations
    // are both adjacent.  An asset directory read into one buffer is 2 copies per 32 MiB chunk instead of 2 per file.
    constexpr uint64_t kMaxGap = 4096;
    auto src_of = [&](const Piece& p) { return files[p.file] + descs[p.file].data_off + 2 * p.word0; };
    auto adjacent = [&](const Piece& a, const Piece& b, uint64_t* gap) -> bool {
        const uint8_t* end_a = src_of(a) + 2 * a.words;
        const uint8_t* beg_b = src_of(b);
        const bool same_file = a.file == b.file && b.word0 == a.word0 + a.words;
        const bool back_to_back = b.file == a.file + 1 && files[a.file] + lens[a.file] == files[b.file];
        if (!(same_file || back_to_back) || beg_b < end_a || (uint64_t)(beg_b - end_a) > kMaxGap) return false;
        *gap = (uint64_t)(beg_b - end_a);
        return true;
    };

    auto flush = [&]() -> int {
        if (pieces.empty()) return BLAST_OK;
        blast_ctx::Lane& lane = ctx->lane[chunk_index % blast_ctx::kPipe];
        chunk_index += 1;
        cudaStream_t st = lane.stream;
        BLAST_CUDA_TRY(cudaStreamSynchronize(st));          // lane buffers are free again
        jobs.clear();
        ht.clear();
        struct Run { const uint8_t* src; uint64_t dev_off, bytes; };
        std::vector<Run> runs;
        uint64_t in_off = 0, out_off = 0;
        for (size_t k = 0; k < pieces.size(); ++k) {
            const Piece& p = pieces[k];
            uint64_t gap = 0;
            if (k > 0 && adjacent(pieces[k - 1], p, &gap)) {
                Run& r = runs.back();                        // device layout mirrors the host layout inside a run
                in_off = r.dev_off + r.bytes + gap;
                r.bytes += gap + 2 * p.words;
            } else {
                in_off = (in_off + 15) & ~15ull;
                runs.push_back(Run{src_of(p), in_off, 2 * p.words});
            }
            int16_t* dst = (d_out && d_out[p.file]) ? d_out[p.file] + p.word0 : (int16_t*)lane.d_tmp + out_off;
            jobs.push_back(blast_pcm_job{lane.d_in + in_off, dst, p.words, descs[p.file].big_endian, 0});
            in_off += 2 * p.words;
            out_off += (p.words + 7) & ~7ull;
        }
        uint64_t words = 0;
        if (int r = fill_tables(jobs.data(), (uint32_t)jobs.size(), (JobDev*)lane.h_jobs, ht, &words)) return r;
        std::memcpy(lane.h_tiles, ht.data(), ht.size() * sizeof(TileRef));
        BLAST_CUDA_TRY(cudaMemcpyAsync(lane.d_jobs, lane.h_jobs, jobs.size() * sizeof(JobDev), cudaMemcpyHostToDevice, st));
        BLAST_CUDA_TRY(cudaMemcpyAsync(lane.d_tiles, lane.h_tiles, ht.size() * sizeof(TileRef), cudaMemcpyHostToDevice, st));
        for (const Run& r : runs)
            BLAST_CUDA_TRY(cudaMemcpyAsync(lane.d_in + r.dev_off, r.src, r.bytes, cudaMemcpyHostToDevice, st));
        int grid = (int)std::min<uint64_t>(ht.size(), (uint64_t)ctx->sm_count * kCtasPerSm);
        pcm16_decode_batch<<<grid, kThreads, 0, st>>>((const JobDev*)lane.d_jobs, (const TileRef*)lane.d_tiles, (uint32_t)ht.size());
        BLAST_CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
        if (host_out) {
            int16_t* h_run = nullptr;
            const int16_t* d_run = nullptr;
            uint64_t run_words = 0;
            auto send = [&]() -> int {
                if (run_words) BLAST_CUDA_TRY(cudaMemcpyAsync(h_run, d_run, 2 * run_words, cudaMemcpyDeviceToHost, st));
                run_words = 0;
                return BLAST_OK;
            };
            for (size_t k = 0; k < pieces.size(); ++k) {
                const Piece& p = pieces[k];
                if (!host_out[p.file]) continue;
                int16_t* h = host_out[p.file] + p.word0;
                const int16_t* d = jobs[k].d_dst;
                if (run_words && h == h_run + run_words && d == d_run + run_words) {
                    run_words += p.words;
                } else {
                    if (int r = send()) return r;
                    h_run = h; d_run = d; run_words = p.words;
                }
            }
            if (int r = send()) return r;
        }
        pieces.clear();
        chunk_bytes = 0;
        return BLAST_OK;
    };

    for (uint32_t i = 0; i < n && rc == BLAST_OK; ++i) {
        uint64_t total = (descs[i].data_len + 1) / 2;
        for (uint64_t w0 = 0; w0 < total && rc == BLAST_OK; w0 += kPieceWords) {
            uint64_t w = std::min(kPieceWords, total - w0);
            uint64_t slot = ((2 * w + 15) & ~15ull) + kMaxGap;     // room for a coalesced gap in front of the piece
            if (!pieces.empty() && (chunk_bytes + slot > kChunkBytes || pieces.size() >= kChunkPieces)) rc = flush();
            pieces.push_back(Piece{i, w0, w});
            chunk_bytes += slot;
        }
    }
    if (rc == BLAST_OK) rc = flush();
    for (int l = 0; l < blast_ctx::kPipe; ++l) {
        cudaError_t e = cudaStreamSynchronize(ctx->lane[l].stream);
        if (e != cudaSuccess && rc == BLAST_OK)
            rc = blast::set_error(BLAST_ERR_CUDA, "decode pipeline failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

int blast_pcm24_unpack_dev(blast_ctx* ctx, const blast_pcm24_job* jobs, uint32_t n_jobs) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(jobs != nullptr || n_jobs == 0, BLAST_ERR_ARG, "blast_pcm24_unpack_dev: null jobs");
    std::vector<Job24Dev> hj;
    hj.reserve(n_jobs);
    uint64_t n_tiles = 0;
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const blast_pcm24_job& in = jobs[j];
        if (in.n_samples && (!in.d_src || !in.d_dst)) return blast::set_error(BLAST_ERR_ARG, "pcm24 job %u: null pointer", j);
        if (in.out_kind > 1) return blast::set_error(BLAST_ERR_ARG, "pcm24 job %u: out_kind must be 0 or 1", j);
        if ((uintptr_t)in.d_dst & (in.out_kind == 0 ? 3 : 1)) return blast::set_error(BLAST_ERR_ARG, "pcm24 job %u: misaligned d_dst", j);
        const uint64_t nt = (in.n_samples + k24SamplesPerTile - 1) / k24SamplesPerTile;
        if (nt == 0) continue;                                         // empty jobs own no tile
        if (n_tiles + nt > 0xFFFFFFFFull) return blast::set_error(BLAST_ERR_CAPACITY, "pcm24 batch too large");
        hj.push_back(Job24Dev{in.d_src, in.d_dst, in.n_samples, in.big_endian, in.out_kind, (uint32_t)n_tiles, 0u});
        n_tiles += nt;
    }
    if (hj.empty()) return BLAST_OK;
    // the job table lives in the context's grow-only scratch (never cudaMalloc once warm); the upload comes from a
    // pageable vector, which the runtime stages before cudaMemcpyAsync returns
    Job24Dev* d_jobs = static_cast<Job24Dev*>(blast::scratch(ctx, 8, hj.size() * sizeof(Job24Dev)));
    if (!d_jobs) return BLAST_ERR_CUDA;
    // the same batch again (the same buffers, step after step): the table in slot 8 is still the one uploaded last time
    const size_t tab_bytes = hj.size() * sizeof(Job24Dev);
    const bool same = ctx->tab8_ptr == d_jobs && ctx->tab8_copy.size() == tab_bytes &&
                      memcmp(ctx->tab8_copy.data(), hj.data(), tab_bytes) == 0;
    if (!same) {
        ctx->tab8_ptr = nullptr;
        BLAST_CUDA_TRY(cudaMemcpyAsync(d_jobs, hj.data(), tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->tab8_copy.assign(reinterpret_cast<const uint8_t*>(hj.data()), reinterpret_cast<const uint8_t*>(hj.data()) + tab_bytes);
        ctx->tab8_ptr = d_jobs;
    }
    const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * k24CtasPerSm);
    pcm24_unpack_batch<<<grid, k24Threads, 0, ctx->stream>>>(d_jobs, (uint32_t)hj.size(), (uint32_t)n_tiles);
    BLAST_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BLAST_OK;
}

}  // extern "C"
