// peer_bus.cu — the window a rank shares with its peers for the render's one exchange step (SURVEY.md §8 e): the int32
// partial bus, the S16 result bus and the flag tables of the tile protocol, mapped into every peer's address space by
// CUDA IPC (one process per GPU) or by peer access (several GPUs driven by one process: blast_group).  The kernels of
// the protocol are in render.cu (sink_tile_flushed / sink_reduce_tile, inside K4 or as two small kernels); this file is
// the host side: allocation, mapping, step counting and the sink the kernels take.
//
// Replaces, across GPUs, the accumulate of Conductor::coordinate (blast/src/audio_processing/engine.rs:46-81: every
// voice adds into one i16 slot, engine.rs:441): i16 wrapping addition is addition mod 2^16, so ranks sum int32 partial
// buses in any order and the low 16 bits are the reference's result.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "render_internal.h"

using namespace blast_rdr;

struct blast_peer_bus {
    blast_ctx* ctx = nullptr;
    uint32_t rank = 0, world = 1, root = 0;
    uint64_t n_slots = 0;
    uint32_t max_tiles = 0;
    uint8_t* window = nullptr;                  // this rank's allocation
    size_t off_bus = 0, off_ready = 0, off_done = 0, off_ack = 0, off_count = 0, off_red = 0, off_err = 0, bytes = 0;
    uint8_t* peer[kMaxPeers] = {};              // every rank's window as mapped here (peer[rank] == window)
    bool ipc[kMaxPeers] = {};                   // opened with cudaIpcOpenMemHandle (to be closed)
    bool connected = false;
    uint32_t step = 0;
    uint32_t timeout_ms = 20000;
    bool fused = false;                         // blast_scene_render_reduce_dev: the exchange inside the render kernel
    // The exchange runs on a stream of its own, behind the render and beside whatever the caller enqueues next: a rank
    // waits for its slowest peer there, not in the stream that carries the next batch's decode.
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_render = nullptr, ev_exch = nullptr;
    bool exch_pending = false;
};

namespace {

constexpr uint32_t kSlotTile = 4096;            // tile of the stand-alone reduction, in bus slots

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void layout(blast_peer_bus* pb) {
    // flags first would be friendlier to read; the partial bus first keeps it 256-byte aligned for free
    size_t o = 0;
    o = align_up(o + pb->n_slots * sizeof(int32_t), 256);
    pb->off_bus = o;
    o = align_up(o + pb->n_slots * sizeof(int16_t), 256);
    pb->off_ready = o;
    o = align_up(o + (size_t)pb->world * pb->max_tiles * sizeof(uint32_t), 256);
    pb->off_done = o;
    o += 256;
    pb->off_ack = o;
    o += 256;
    pb->off_count = o;
    o = align_up(o + (size_t)pb->max_tiles * sizeof(uint32_t), 256);
    pb->off_red = o;
    o += 128;
    pb->off_err = o;
    o += 128;
    pb->bytes = o;
}

}  // namespace

namespace blast_rdr {

int32_t* peer_bus_partial(blast_peer_bus* pb) { return reinterpret_cast<int32_t*>(pb->window); }
bool peer_bus_fused(const blast_peer_bus* pb) { return pb->fused; }

int peer_bus_exchange(blast_ctx* ctx, blast_peer_bus* pb, const BusSink& sink) {
    if (pb->world == 1) return launch_bus_reduce(ctx, sink);
    BLAST_CUDA_TRY(cudaEventRecord(pb->ev_render, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamWaitEvent(pb->aux, pb->ev_render, 0));
    if (int rc = launch_bus_reduce(ctx, sink, pb->aux)) return rc;
    if (pb->rank == pb->root)                   // the root's bus is complete when every rank has stored its tiles
        if (int rc = launch_flag_wait(ctx, reinterpret_cast<const uint32_t*>(pb->window + pb->off_done), pb->world, sink.step,
                                      pb->timeout_ms, reinterpret_cast<uint32_t*>(pb->window + pb->off_err), pb->aux))
            return rc;
    BLAST_CUDA_TRY(cudaEventRecord(pb->ev_exch, pb->aux));
    pb->exch_pending = true;
    return BLAST_OK;
}

int peer_bus_next_step(blast_ctx* ctx, blast_peer_bus* pb, uint64_t frames, uint32_t oc, bool in_render, BusSink* out) {
    if (pb->ctx != ctx) return blast::set_error(BLAST_ERR_ARG, "the peer bus belongs to another context");
    if (!pb->connected) return blast::set_error(BLAST_ERR_ARG, "the peer bus is not connected (blast_peer_bus_connect_*)");
    const uint64_t slots = frames * oc;
    if (slots > pb->n_slots) return blast::set_error(BLAST_ERR_CAPACITY, "%llu bus slots asked, the peer bus holds %llu",
                                                     (unsigned long long)slots, (unsigned long long)pb->n_slots);
    BusSink s{};
    s.world = pb->world;
    s.rank = pb->rank;
    s.step = ++pb->step;
    s.lag = 1;
    s.max_tiles = pb->max_tiles;
    s.tile_slots = in_render ? (uint32_t)kFT * oc : kSlotTile;
    s.n_slots = slots;
    s.n_tiles = (uint32_t)((slots + s.tile_slots - 1) / s.tile_slots);
    s.n_my_tiles = s.n_tiles > s.rank ? (s.n_tiles - s.rank + s.world - 1) / s.world : 0;
    s.timeout_ms = pb->timeout_ms;
    for (uint32_t r = 0; r < pb->world; ++r) {
        s.part[r] = reinterpret_cast<const int32_t*>(pb->peer[r]);
        s.ready_at[r] = reinterpret_cast<uint32_t*>(pb->peer[r] + pb->off_ready) + (size_t)pb->rank * pb->max_tiles;
    }
    s.ready_mine = reinterpret_cast<const uint32_t*>(pb->window + pb->off_ready);
    s.out = reinterpret_cast<int16_t*>(pb->peer[pb->root] + pb->off_bus);
    s.tile_count = reinterpret_cast<uint32_t*>(pb->window + pb->off_count);
    s.red_count = reinterpret_cast<uint32_t*>(pb->window + pb->off_red);
    s.ack_mine = reinterpret_cast<const uint32_t*>(pb->window + pb->off_ack);
    s.err = reinterpret_cast<uint32_t*>(pb->window + pb->off_err);
    // when all my tiles are reduced: "my part of the bus is in place" to the root, "I am done reading your partial bus"
    // to every rank (myself included: the next step's wait is then uniform)
    uint32_t n = 0;
    s.done[n++] = reinterpret_cast<uint32_t*>(pb->peer[pb->root] + pb->off_done) + pb->rank;
    for (uint32_t r = 0; r < pb->world; ++r) s.done[n++] = reinterpret_cast<uint32_t*>(pb->peer[r] + pb->off_ack) + pb->rank;
    s.n_done = n;
    *out = s;
    return BLAST_OK;
}

}  // namespace blast_rdr

extern "C" {

int blast_peer_bus_create(blast_ctx* ctx, uint64_t n_slots, uint32_t rank, uint32_t world, uint32_t root, blast_peer_bus** out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_peer_bus_create: out is null");
    *out = nullptr;
    if (world < 1 || world > (uint32_t)kMaxPeers || rank >= world || root >= world)
        return blast::set_error(BLAST_ERR_ARG, "blast_peer_bus_create: need rank, root < world <= %d", kMaxPeers);
    BLAST_REQUIRE(n_slots >= 1 && n_slots < (1ull << 40), BLAST_ERR_ARG, "blast_peer_bus_create: bad slot count");
    auto* pb = new blast_peer_bus();
    pb->ctx = ctx;
    pb->rank = rank;
    pb->world = world;
    pb->root = root;
    pb->n_slots = n_slots;
    pb->max_tiles = (uint32_t)((n_slots + kFT - 1) / kFT) + 1;          // a 1-channel bus has the smallest tiles
    if (const char* e = getenv("BLAST_PEER_TIMEOUT_MS")) pb->timeout_ms = (uint32_t)std::max(0, atoi(e));
    layout(pb);
    if (cudaMalloc(&pb->window, pb->bytes) != cudaSuccess) {
        const int rc = blast::set_error(BLAST_ERR_CUDA, "blast_peer_bus_create: cudaMalloc of %zu bytes failed: %s", pb->bytes,
                                        cudaGetErrorString(cudaGetLastError()));
        delete pb;
        return rc;
    }
    // flags start at step 0; the partial bus is cleared so that a reduction never reads undefined memory
    if (cudaMemsetAsync(pb->window, 0, pb->bytes, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        const int rc = blast::set_error(BLAST_ERR_CUDA, "blast_peer_bus_create: clearing the window failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(pb->window);
        delete pb;
        return rc;
    }
    pb->peer[rank] = pb->window;
    pb->connected = world == 1;
    if (world > 1) {
        if (int rc = preload_kernels(ctx)) { blast_peer_bus_destroy(ctx, pb); return rc; }
        if (cudaStreamCreateWithFlags(&pb->aux, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&pb->ev_render, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&pb->ev_exch, cudaEventDisableTiming) != cudaSuccess) {
            const int rc = blast::set_error(BLAST_ERR_CUDA, "blast_peer_bus_create: stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            blast_peer_bus_destroy(ctx, pb);
            return rc;
        }
    }
    *out = pb;
    return BLAST_OK;
}

void blast_peer_bus_destroy(blast_ctx* ctx, blast_peer_bus* pb) {
    if (!pb) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (pb->aux) { cudaStreamSynchronize(pb->aux); cudaStreamDestroy(pb->aux); }
    if (pb->ev_render) cudaEventDestroy(pb->ev_render);
    if (pb->ev_exch) cudaEventDestroy(pb->ev_exch);
    for (uint32_t r = 0; r < pb->world; ++r)
        if (pb->ipc[r] && pb->peer[r]) cudaIpcCloseMemHandle(pb->peer[r]);
    if (pb->window) cudaFree(pb->window);
    delete pb;
}

int blast_peer_bus_export(blast_ctx* ctx, blast_peer_bus* pb, uint8_t handle_out[BLAST_PEER_HANDLE_BYTES]) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(pb && handle_out, BLAST_ERR_ARG, "blast_peer_bus_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == BLAST_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    BLAST_CUDA_TRY(cudaIpcGetMemHandle(&h, pb->window));
    std::memcpy(handle_out, &h, sizeof(h));
    return BLAST_OK;
}

int blast_peer_bus_connect_ipc(blast_ctx* ctx, blast_peer_bus* pb, const uint8_t* handles) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(pb && handles, BLAST_ERR_ARG, "blast_peer_bus_connect_ipc: null argument");
    BLAST_REQUIRE(!pb->connected || pb->world == 1, BLAST_ERR_ARG, "blast_peer_bus_connect_ipc: already connected");
    for (uint32_t r = 0; r < pb->world; ++r) {
        if (r == pb->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * BLAST_PEER_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (uint32_t q = 0; q < r; ++q)
                if (pb->ipc[q]) { cudaIpcCloseMemHandle(pb->peer[q]); pb->ipc[q] = false; pb->peer[q] = nullptr; }
            return blast::set_error(BLAST_ERR_CUDA, "cudaIpcOpenMemHandle of rank %u's window failed: %s", r, cudaGetErrorString(e));
        }
        pb->peer[r] = static_cast<uint8_t*>(p);
        pb->ipc[r] = true;
    }
    pb->connected = true;
    return BLAST_OK;
}

int blast_peer_bus_connect_local(blast_peer_bus* const* all, uint32_t world) {
    BLAST_REQUIRE(all != nullptr && world >= 1 && world <= (uint32_t)kMaxPeers, BLAST_ERR_ARG, "blast_peer_bus_connect_local: bad arguments");
    for (uint32_t r = 0; r < world; ++r) {
        BLAST_REQUIRE(all[r] != nullptr, BLAST_ERR_ARG, "blast_peer_bus_connect_local: null peer bus");
        if (all[r]->world != world || all[r]->rank != r || all[r]->n_slots != all[0]->n_slots || all[r]->root != all[0]->root)
            return blast::set_error(BLAST_ERR_ARG, "blast_peer_bus_connect_local: all[%u] is not rank %u of %u over the same bus", r, r, world);
    }
    for (uint32_t r = 0; r < world; ++r) {
        const int dev = all[r]->ctx->device;
        BLAST_CUDA_TRY(cudaSetDevice(dev));
        for (uint32_t q = 0; q < world; ++q) {
            const int other = all[q]->ctx->device;
            if (other != dev) {
                int can = 0;
                BLAST_CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, other));
                if (!can) return blast::set_error(BLAST_ERR_UNSUPPORTED, "GPU %d cannot map the memory of GPU %d (no peer access)", dev, other);
                cudaError_t e = cudaDeviceEnablePeerAccess(other, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return blast::set_error(BLAST_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", dev, other, cudaGetErrorString(e));
            }
            all[r]->peer[q] = all[q]->window;
        }
        all[r]->connected = true;
    }
    return BLAST_OK;
}

int blast_peer_bus_set_fused(blast_peer_bus* pb, int fused) {
    BLAST_REQUIRE(pb != nullptr, BLAST_ERR_ARG, "blast_peer_bus_set_fused: null peer bus");
    pb->fused = fused != 0;
    return BLAST_OK;
}

int32_t* blast_peer_bus_partial(blast_peer_bus* pb) { return pb ? reinterpret_cast<int32_t*>(pb->window) : nullptr; }
int16_t* blast_peer_bus_bus(blast_peer_bus* pb) { return pb ? reinterpret_cast<int16_t*>(pb->window + pb->off_bus) : nullptr; }

int blast_peer_bus_begin_dev(blast_ctx* ctx, blast_peer_bus* pb) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(pb && pb->ctx == ctx && pb->connected, BLAST_ERR_ARG, "blast_peer_bus_begin_dev: no connected peer bus of this context");
    if (pb->world == 1) return BLAST_OK;
    return launch_flag_wait(ctx, reinterpret_cast<const uint32_t*>(pb->window + pb->off_ack), pb->world, pb->step, pb->timeout_ms,
                            reinterpret_cast<uint32_t*>(pb->window + pb->off_err));
}

int blast_peer_bus_reduce_dev(blast_ctx* ctx, blast_peer_bus* pb, uint64_t n_slots_used) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(pb != nullptr, BLAST_ERR_ARG, "blast_peer_bus_reduce_dev: null peer bus");
    BusSink sink;
    if (int rc = peer_bus_next_step(ctx, pb, n_slots_used, 1, false, &sink)) return rc;
    return peer_bus_exchange(ctx, pb, sink);
}

int blast_peer_bus_wait_dev(blast_ctx* ctx, blast_peer_bus* pb) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(pb && pb->ctx == ctx, BLAST_ERR_ARG, "blast_peer_bus_wait_dev: no peer bus of this context");
    if (pb->world == 1) return BLAST_OK;
    if (pb->exch_pending) {                     // the exchange ran beside the stream: join it
        BLAST_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, pb->ev_exch, 0));
        pb->exch_pending = false;
        return BLAST_OK;
    }
    if (pb->rank != pb->root) return BLAST_OK;
    return launch_flag_wait(ctx, reinterpret_cast<const uint32_t*>(pb->window + pb->off_done), pb->world, pb->step, pb->timeout_ms,
                            reinterpret_cast<uint32_t*>(pb->window + pb->off_err));   // (fused: the exchange was in the render kernel)
}

int blast_peer_bus_flags(blast_ctx* ctx, blast_peer_bus* pb, uint32_t* out, uint32_t cap) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(pb && out && pb->ctx == ctx, BLAST_ERR_ARG, "blast_peer_bus_flags: bad argument");
    // layout of the dump: step, world, then per rank r: ready[r][tile 0..3], done[r], ack[r]; then the error word and the reduce count
    std::vector<uint32_t> v;
    v.push_back(pb->step);
    v.push_back(pb->world);
    cudaStream_t s = nullptr;
    BLAST_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));       // beside whatever is still waiting
    auto rd = [&](size_t off) { uint32_t x = 0xDEADBEEFu; cudaMemcpyAsync(&x, pb->window + off, 4, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); return x; };
    for (uint32_t r = 0; r < pb->world; ++r) {
        for (uint32_t t = 0; t < 4; ++t) v.push_back(rd(pb->off_ready + ((size_t)r * pb->max_tiles + t) * 4));
        v.push_back(rd(pb->off_done + r * 4));
        v.push_back(rd(pb->off_ack + r * 4));
    }
    v.push_back(rd(pb->off_err));
    v.push_back(rd(pb->off_red));
    cudaStreamDestroy(s);
    for (uint32_t i = 0; i < cap && i < v.size(); ++i) out[i] = v[i];
    return BLAST_OK;
}

int blast_peer_bus_check(blast_ctx* ctx, blast_peer_bus* pb) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(pb && pb->ctx == ctx, BLAST_ERR_ARG, "blast_peer_bus_check: no peer bus of this context");
    uint32_t* mb = static_cast<uint32_t*>(blast::mailbox(ctx));
    if (!mb) return BLAST_ERR_CUDA;
    if (pb->aux) BLAST_CUDA_TRY(cudaStreamSynchronize(pb->aux));
    BLAST_CUDA_TRY(cudaMemcpyAsync(mb, pb->window + pb->off_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (*mb & 4u) {
        BLAST_CUDA_TRY(cudaMemsetAsync(pb->window + pb->off_err, 0, sizeof(uint32_t), ctx->stream));
        return blast::set_error(BLAST_ERR_TIMEOUT, "rank %u gave up waiting for a peer GPU after %u ms (step %u): the bus of that step is incomplete",
                                pb->rank, pb->timeout_ms, pb->step);
    }
    return BLAST_OK;
}

}  // extern "C"
