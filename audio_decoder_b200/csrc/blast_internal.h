// blast_internal.h — shared internals of libblast_cuda.so (not part of the ABI)
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/blast_cuda.h"

struct blast_ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    uint64_t launches = 0;
    // Pipeline lanes for the host-buffer entry points (created lazily, grow-only, freed at
    // destroy): each lane has its own stream, device staging slabs and pinned table buffers.
    static constexpr int kPipe = 4;
    struct Lane {
        cudaStream_t stream = nullptr;
        uint8_t* d_in = nullptr;      // staged input bytes
        uint8_t* d_tmp = nullptr;     // scratch output when the caller keeps nothing on the device
        void* d_jobs = nullptr;
        void* d_tiles = nullptr;
        void* h_jobs = nullptr;       // pinned
        void* h_tiles = nullptr;      // pinned
    };
    Lane lane[kPipe];
    size_t lane_bytes = 0, lane_jobs = 0, lane_tiles = 0;
    // grow-only device scratch slots + a small pinned mailbox: hot entry points never cudaMalloc / cudaFree
    static constexpr int kScratch = 12;
    void* scratch[kScratch] = {nullptr};
    size_t scratch_cap[kScratch] = {0};
    void* mailbox = nullptr;          // 4 KiB pinned host memory for small read-backs
    // per-context (= per-device) caches: function attributes are per device, scratch contents per context
    int mpeg_ctas_per_sm = 0;         // occupancy of mpeg_walk once its dynamic shared-memory limit has been raised
    uint64_t x128p_split_sub = 0;     // the sub-stream jump matrices J^(2^b) held in scratch slot 5
    int x128p_split_mats = 0;
    void* x128p_split_ptr = nullptr;
    std::vector<uint8_t> tab8_copy;   // the 24-bit unpack's job table as last uploaded into scratch slot 8 (a batch that is
    const void* tab8_ptr = nullptr;   // unpacked again — the same buffers, step after step — uploads nothing)
    int render_ctas_per_sm = 3;       // K4's persistent CTAs per SM (3 fit); lowered when several contexts that wait for
                                      // each other inside K4 share one GPU (blast_group over a repeated device id)
};

struct blast_event {
    cudaEvent_t ev = nullptr;
    int device = -1;
};

namespace blast {

int set_error(int code, const char* fmt, ...);

#define BLAST_CUDA_TRY(expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return blast::set_error(BLAST_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    } while (0)

#define BLAST_REQUIRE(cond, code, msg)                          \
    do {                                                        \
        if (!(cond)) return blast::set_error((code), "%s", (msg)); \
    } while (0)

inline int bind(blast_ctx* ctx) {
    if (!ctx) return set_error(BLAST_ERR_ARG, "null context");
    BLAST_CUDA_TRY(cudaSetDevice(ctx->device));
    return BLAST_OK;
}

int ensure_pipe(blast_ctx* ctx, size_t chunk_bytes, size_t max_jobs, size_t max_tiles, size_t job_size,
                size_t tile_size);
void release_pipe(blast_ctx* ctx);
// device scratch slot `slot` with at least `bytes` bytes (contents undefined); nullptr + error on failure
void* scratch(blast_ctx* ctx, int slot, size_t bytes);
void* mailbox(blast_ctx* ctx);

// 128-bit streaming accessors (read-only path, no L1 allocation: every byte is touched once)
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_cached(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

}  // namespace blast
