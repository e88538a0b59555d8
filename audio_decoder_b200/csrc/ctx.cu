// ctx.cu — context lifecycle, memory helpers, event timing for libblast_cuda.so
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "blast_internal.h"

namespace blast {

static thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

void release_pipe(blast_ctx* ctx) {
    for (int i = 0; i < blast_ctx::kPipe; ++i) {
        blast_ctx::Lane& l = ctx->lane[i];
        if (l.stream) cudaStreamSynchronize(l.stream);
        if (l.d_in) cudaFree(l.d_in);
        if (l.d_tmp) cudaFree(l.d_tmp);
        if (l.d_jobs) cudaFree(l.d_jobs);
        if (l.d_tiles) cudaFree(l.d_tiles);
        if (l.h_jobs) cudaFreeHost(l.h_jobs);
        if (l.h_tiles) cudaFreeHost(l.h_tiles);
        cudaStream_t keep = l.stream;
        l = blast_ctx::Lane();
        l.stream = keep;
    }
    ctx->lane_bytes = ctx->lane_jobs = ctx->lane_tiles = 0;
}

int ensure_pipe(blast_ctx* ctx, size_t chunk_bytes, size_t max_jobs, size_t max_tiles, size_t job_size,
                size_t tile_size) {
    for (int i = 0; i < blast_ctx::kPipe; ++i)
        if (!ctx->lane[i].stream)
            BLAST_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->lane[i].stream, cudaStreamNonBlocking));
    if (ctx->lane_bytes >= chunk_bytes && ctx->lane_jobs >= max_jobs * job_size && ctx->lane_tiles >= max_tiles * tile_size)
        return BLAST_OK;
    release_pipe(ctx);
    for (int i = 0; i < blast_ctx::kPipe; ++i) {
        blast_ctx::Lane& l = ctx->lane[i];
        BLAST_CUDA_TRY(cudaMalloc(&l.d_in, chunk_bytes + 256));
        BLAST_CUDA_TRY(cudaMalloc(&l.d_tmp, chunk_bytes + 256));
        BLAST_CUDA_TRY(cudaMalloc(&l.d_jobs, max_jobs * job_size));
        BLAST_CUDA_TRY(cudaMalloc(&l.d_tiles, max_tiles * tile_size));
        BLAST_CUDA_TRY(cudaHostAlloc(&l.h_jobs, max_jobs * job_size, cudaHostAllocDefault));
        BLAST_CUDA_TRY(cudaHostAlloc(&l.h_tiles, max_tiles * tile_size, cudaHostAllocDefault));
    }
    ctx->lane_bytes = chunk_bytes;
    ctx->lane_jobs = max_jobs * job_size;
    ctx->lane_tiles = max_tiles * tile_size;
    return BLAST_OK;
}

void* scratch(blast_ctx* ctx, int slot, size_t bytes) {
    if (slot < 0 || slot >= blast_ctx::kScratch) { set_error(BLAST_ERR_ARG, "bad scratch slot"); return nullptr; }
    if (ctx->scratch_cap[slot] >= bytes && ctx->scratch[slot]) return ctx->scratch[slot];
    if (ctx->scratch[slot]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->scratch[slot]);
        ctx->scratch[slot] = nullptr;
        ctx->scratch_cap[slot] = 0;
        // slot 5 caches the RNG's sub-stream jump matrices: a new allocation may come back at the same address with
        // undefined contents, so the cache key (which includes the pointer) is dropped with the memory
        if (slot == 5) ctx->x128p_split_ptr = nullptr;
        if (slot == 8) ctx->tab8_ptr = nullptr;
    }
    size_t want = (bytes + (bytes >> 2) + 255) & ~(size_t)255;      // 25 % head-room
    if (cudaMalloc(&ctx->scratch[slot], want) != cudaSuccess) {
        ctx->scratch[slot] = nullptr;
        set_error(BLAST_ERR_CUDA, "scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    ctx->scratch_cap[slot] = want;
    return ctx->scratch[slot];
}

void* mailbox(blast_ctx* ctx) {
    if (!ctx->mailbox && cudaHostAlloc(&ctx->mailbox, 4096, cudaHostAllocDefault) != cudaSuccess) {
        ctx->mailbox = nullptr;
        set_error(BLAST_ERR_CUDA, "pinned mailbox allocation failed");
    }
    return ctx->mailbox;
}

}  // namespace blast

using blast::set_error;

extern "C" {

int blast_abi_version(void) { return BLAST_ABI_VERSION; }

const char* blast_last_error(void) { return blast::g_last_error.c_str(); }

int blast_ctx_create(blast_ctx** out, int device) {
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_ctx_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(BLAST_ERR_NO_DEVICE,
                         "no CUDA device (%s); libblast_cuda has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return set_error(BLAST_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    BLAST_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(BLAST_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                         prop.major, prop.minor);
    BLAST_CUDA_TRY(cudaSetDevice(device));
    blast_ctx* ctx = new blast_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t se = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
        delete ctx;
        return set_error(BLAST_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(se));
    }
    ctx->owns_stream = true;
    if (const char* e = getenv("BLAST_RENDER_MAX_CTAS_PER_SM")) ctx->render_ctas_per_sm = std::min(3, std::max(1, atoi(e)));
    *out = ctx;
    return BLAST_OK;
}

void blast_ctx_destroy(blast_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    blast::release_pipe(ctx);
    for (int i = 0; i < blast_ctx::kPipe; ++i)
        if (ctx->lane[i].stream) cudaStreamDestroy(ctx->lane[i].stream);
    for (int i = 0; i < blast_ctx::kScratch; ++i)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int blast_ctx_trim(blast_ctx* ctx) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < blast_ctx::kScratch; ++i) {
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
        ctx->scratch[i] = nullptr;
        ctx->scratch_cap[i] = 0;
    }
    ctx->x128p_split_ptr = nullptr;                 // the cached jump matrices lived in scratch
    ctx->tab8_ptr = nullptr;                        // and so did the cached 24-bit job table
    blast::release_pipe(ctx);
    return BLAST_OK;
}

int blast_ctx_set_stream(blast_ctx* ctx, void* cuda_stream) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->owns_stream) BLAST_CUDA_TRY(cudaStreamDestroy(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->owns_stream = false;
    return BLAST_OK;
}

void* blast_ctx_stream(blast_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int blast_ctx_sync(blast_ctx* ctx) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BLAST_OK;
}

int blast_ctx_device(const blast_ctx* ctx) { return ctx ? ctx->device : -1; }
int blast_ctx_sm_count(const blast_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t blast_ctx_launch_count(const blast_ctx* ctx) { return ctx ? ctx->launches : 0; }

int blast_dev_alloc(blast_ctx* ctx, size_t bytes, void** out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_dev_alloc: out is null");
    size_t padded = (bytes + 255) & ~(size_t)255;
    if (padded == 0) padded = 256;
    BLAST_CUDA_TRY(cudaMalloc(out, padded));
    return BLAST_OK;
}

int blast_dev_free(blast_ctx* ctx, void* p) {
    if (int rc = blast::bind(ctx)) return rc;
    if (p) BLAST_CUDA_TRY(cudaFree(p));
    return BLAST_OK;
}

int blast_host_alloc(blast_ctx* ctx, size_t bytes, void** out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_host_alloc: out is null");
    BLAST_CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return BLAST_OK;
}

int blast_host_free(blast_ctx* ctx, void* p) {
    if (int rc = blast::bind(ctx)) return rc;
    if (p) BLAST_CUDA_TRY(cudaFreeHost(p));
    return BLAST_OK;
}

int blast_memcpy_h2d(blast_ctx* ctx, void* d_dst, const void* src, size_t bytes) {
    if (int rc = blast::bind(ctx)) return rc;
    if (bytes) BLAST_CUDA_TRY(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return BLAST_OK;
}

int blast_memcpy_d2h(blast_ctx* ctx, void* dst, const void* d_src, size_t bytes) {
    if (int rc = blast::bind(ctx)) return rc;
    if (bytes) BLAST_CUDA_TRY(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return BLAST_OK;
}

int blast_memcpy_d2d(blast_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
    if (int rc = blast::bind(ctx)) return rc;
    if (bytes) BLAST_CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return BLAST_OK;
}

int blast_memset_dev(blast_ctx* ctx, void* d_dst, int value, size_t bytes) {
    if (int rc = blast::bind(ctx)) return rc;
    if (bytes) BLAST_CUDA_TRY(cudaMemsetAsync(d_dst, value, bytes, ctx->stream));
    return BLAST_OK;
}

int blast_event_create(blast_ctx* ctx, blast_event** out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_event_create: out is null");
    blast_event* ev = new blast_event();
    ev->device = ctx->device;
    cudaError_t e = cudaEventCreate(&ev->ev);
    if (e != cudaSuccess) {
        delete ev;
        return set_error(BLAST_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e));
    }
    *out = ev;
    return BLAST_OK;
}

void blast_event_destroy(blast_event* ev) {
    if (!ev) return;
    cudaSetDevice(ev->device);
    cudaEventDestroy(ev->ev);
    delete ev;
}

int blast_event_record(blast_ctx* ctx, blast_event* ev) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(ev != nullptr, BLAST_ERR_ARG, "blast_event_record: null event");
    BLAST_CUDA_TRY(cudaEventRecord(ev->ev, ctx->stream));
    return BLAST_OK;
}

int blast_event_elapsed_ms(blast_event* start, blast_event* stop, float* ms_out) {
    BLAST_REQUIRE(start && stop && ms_out, BLAST_ERR_ARG, "blast_event_elapsed_ms: null argument");
    BLAST_CUDA_TRY(cudaSetDevice(stop->device));
    BLAST_CUDA_TRY(cudaEventSynchronize(stop->ev));
    BLAST_CUDA_TRY(cudaEventElapsedTime(ms_out, start->ev, stop->ev));
    return BLAST_OK;
}

}  // extern "C"
