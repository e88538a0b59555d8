// group.cu — several GPUs of one box driven by ONE host process through the C ABI (no torch, no launcher, no
// collective library): what a single-process host like the reference's main() (blast/src/main.rs:13-128) needs to use
// every GPU.  A blast_group owns one blast_ctx per member and maps the members' memory into each other (peer access over
// NVLink / NVSwitch).  The path shards as SURVEY.md §8(e) says:
//   decode   file i  -> member i mod n, the decoded track stays in that member's HBM           (no exchange)
//   render   a voice is rendered where its track lives; one exchange step: the bus reduction, tile by tile inside the
//            render kernel over peer memory (blast_peer_bus, render.cu)
//   RNG      stream s -> member s mod n                                                        (no exchange)
//   MPEG     one stream cut into byte ranges; 48-byte range aggregates folded on the host, header histogram and
//            first-position table reduced through peer memory
// Calls that block (decode, conductor spans) run one host thread per member so that the GPUs work concurrently.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "blast_internal.h"

struct blast_group {
    uint32_t n = 0;
    std::vector<blast_ctx*> ctx;
    std::vector<int> device;
    std::vector<blast_peer_bus*> pb;          // one per member, connected; sized for pb_slots
    uint64_t pb_slots = 0;
    bool fused = false;                       // blast_group_render: the exchange inside the render kernel
    std::vector<std::vector<void*>> slabs;    // device memory handed out as tracks, per member
};

struct blast_group_conductor {
    blast_group* g = nullptr;
    std::vector<blast_conductor*> c;
    uint32_t out_channels = 0;
};

namespace {

// fn(member) on every member, one host thread each; the first failure (by member order) is reported to the caller
template <typename F>
int for_members(blast_group* g, F fn) {
    if (g->n == 1) return fn(0u);
    std::vector<int> rc(g->n, BLAST_OK);
    std::vector<std::string> msg(g->n);
    std::vector<std::thread> th;
    th.reserve(g->n);
    for (uint32_t m = 0; m < g->n; ++m)
        th.emplace_back([&, m] {
            rc[m] = fn(m);
            if (rc[m] != BLAST_OK) msg[m] = blast_last_error();     // the message is thread-local
        });
    for (auto& t : th) t.join();
    for (uint32_t m = 0; m < g->n; ++m)
        if (rc[m] != BLAST_OK) return blast::set_error(rc[m], "member %u (GPU %d): %s", m, g->device[m], msg[m].c_str());
    return BLAST_OK;
}

void free_peer_buses(blast_group* g) {
    for (uint32_t m = 0; m < g->pb.size(); ++m)
        if (g->pb[m]) blast_peer_bus_destroy(g->ctx[m], g->pb[m]);
    g->pb.clear();
    g->pb_slots = 0;
}

int ensure_peer_buses(blast_group* g, uint64_t n_slots) {
    if (g->pb_slots >= n_slots && !g->pb.empty()) return BLAST_OK;
    free_peer_buses(g);
    g->pb.assign(g->n, nullptr);
    for (uint32_t m = 0; m < g->n; ++m)
        if (int rc = blast_peer_bus_create(g->ctx[m], std::max<uint64_t>(n_slots, 1), m, g->n, 0, &g->pb[m])) { free_peer_buses(g); return rc; }
    if (int rc = blast_peer_bus_connect_local(g->pb.data(), g->n)) { free_peer_buses(g); return rc; }
    for (uint32_t m = 0; m < g->n; ++m) blast_peer_bus_set_fused(g->pb[m], g->fused ? 1 : 0);
    g->pb_slots = std::max<uint64_t>(n_slots, 1);
    return BLAST_OK;
}

// dst[i] += src[i] / dst[i] = min(dst[i], src[i]); src may be a peer GPU's memory
__global__ void add_u32_from(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
__global__ void min_u64_from(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long a = dst[i], b = src[i];
        if (b < a) dst[i] = b;
    }
}

constexpr uint64_t kMpegSpan = 32768, kMpegHalo = 16;

}  // namespace

extern "C" {

int blast_group_create(blast_group** out, const int* device_ids, uint32_t n_devices) {
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_group_create: out is null");
    *out = nullptr;
    BLAST_REQUIRE(device_ids != nullptr && n_devices >= 1 && n_devices <= 16, BLAST_ERR_ARG, "blast_group_create: 1..16 device ids");
    auto* g = new blast_group();
    g->n = n_devices;
    g->slabs.resize(n_devices);
    for (uint32_t m = 0; m < n_devices; ++m) {
        blast_ctx* c = nullptr;
        if (int rc = blast_ctx_create(&c, device_ids[m])) { blast_group_destroy(g); return rc; }
        g->ctx.push_back(c);
        g->device.push_back(device_ids[m]);
    }
    // a device id may repeat (several members on one GPU: the multi-member protocol on a single-GPU box).  Members
    // wait for each other inside the render kernel, so all of their persistent CTAs must be resident at once.
    for (uint32_t m = 0; m < n_devices; ++m) {
        const int same = (int)std::count(g->device.begin(), g->device.end(), g->device[m]);
        if (same > 1) g->ctx[m]->render_ctas_per_sm = std::max(1, std::min(g->ctx[m]->render_ctas_per_sm, 3 / same));
        if (same > 3) { blast_group_destroy(g); return blast::set_error(BLAST_ERR_UNSUPPORTED, "at most 3 group members per GPU"); }
    }
    for (uint32_t a = 0; a < n_devices; ++a)
        for (uint32_t b = 0; b < n_devices; ++b) {
            if (g->device[a] == g->device[b]) continue;
            int can = 0;
            cudaSetDevice(g->device[a]);
            cudaDeviceCanAccessPeer(&can, g->device[a], g->device[b]);
            if (!can) { blast_group_destroy(g); return blast::set_error(BLAST_ERR_UNSUPPORTED, "GPU %d cannot map the memory of GPU %d (no peer access)", g->device[a], g->device[b]); }
            cudaError_t e = cudaDeviceEnablePeerAccess(g->device[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { blast_group_destroy(g); return blast::set_error(BLAST_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", g->device[a], g->device[b], cudaGetErrorString(e)); }
        }
    *out = g;
    return BLAST_OK;
}

void blast_group_destroy(blast_group* g) {
    if (!g) return;
    for (uint32_t m = 0; m < g->ctx.size(); ++m) blast_ctx_sync(g->ctx[m]);
    free_peer_buses(g);
    blast_group_free_tracks(g);
    for (auto* c : g->ctx) blast_ctx_destroy(c);
    delete g;
}

int blast_group_set_fused(blast_group* g, int fused) {
    BLAST_REQUIRE(g != nullptr, BLAST_ERR_ARG, "blast_group_set_fused: null group");
    g->fused = fused != 0;
    for (uint32_t m = 0; m < g->pb.size(); ++m)
        if (g->pb[m]) blast_peer_bus_set_fused(g->pb[m], g->fused ? 1 : 0);
    return BLAST_OK;
}

uint32_t blast_group_size(const blast_group* g) { return g ? g->n : 0; }
blast_ctx* blast_group_ctx(blast_group* g, uint32_t member) { return (g && member < g->n) ? g->ctx[member] : nullptr; }

int blast_group_free_tracks(blast_group* g) {
    BLAST_REQUIRE(g != nullptr, BLAST_ERR_ARG, "blast_group_free_tracks: null group");
    for (uint32_t m = 0; m < g->slabs.size(); ++m) {
        if (m < g->ctx.size()) { cudaSetDevice(g->device[m]); cudaStreamSynchronize(g->ctx[m]->stream); }
        for (void* p : g->slabs[m]) cudaFree(p);
        g->slabs[m].clear();
    }
    return BLAST_OK;
}

int blast_group_pcm_decode_batch(blast_group* g, uint32_t n, const uint8_t* const* files, const size_t* lens,
                                 const blast_pcm_desc* descs, int16_t* const* host_out, blast_track* tracks_out) {
    BLAST_REQUIRE(g && (n == 0 || (files && lens && descs)), BLAST_ERR_ARG, "blast_group_pcm_decode_batch: null argument");
    return for_members(g, [&](uint32_t m) -> int {
        blast_ctx* ctx = g->ctx[m];
        if (int rc = blast::bind(ctx)) return rc;
        std::vector<const uint8_t*> f;
        std::vector<size_t> l;
        std::vector<blast_pcm_desc> d;
        std::vector<int16_t*> ho, dv;
        std::vector<uint32_t> idx;
        size_t bytes = 0;
        for (uint32_t i = m; i < n; i += g->n) {
            idx.push_back(i);
            f.push_back(files[i]);
            l.push_back(lens[i]);
            d.push_back(descs[i]);
            ho.push_back(host_out ? host_out[i] : nullptr);
            bytes += (blast_pcm_out_len(&descs[i]) * sizeof(int16_t) + 255) & ~(size_t)255;
        }
        if (idx.empty()) return BLAST_OK;
        uint8_t* slab = nullptr;
        if (tracks_out) {                                   // one slab per member and batch; tracks are 256-byte aligned slices
            BLAST_CUDA_TRY(cudaMalloc(&slab, bytes + 256));
            g->slabs[m].push_back(slab);
            size_t off = 0;
            for (size_t k = 0; k < idx.size(); ++k) {
                dv.push_back(reinterpret_cast<int16_t*>(slab + off));
                off += (blast_pcm_out_len(&d[k]) * sizeof(int16_t) + 255) & ~(size_t)255;
            }
        }
        if (int rc = blast_pcm_decode_batch(ctx, (uint32_t)idx.size(), f.data(), l.data(), d.data(), host_out ? ho.data() : nullptr,
                                            tracks_out ? dv.data() : nullptr))
            return rc;
        if (tracks_out)
            for (size_t k = 0; k < idx.size(); ++k)
                tracks_out[idx[k]] = blast_track{dv[k], (uint64_t)blast_pcm_out_len(&d[k]), d[k].num_channels, d[k].sample_rate};
        return BLAST_OK;
    });
}

int blast_group_render(blast_group* g, const blast_track* tracks, uint32_t n_tracks, const blast_voice* voices, uint32_t n_voices,
                       uint32_t out_channels, uint64_t frames, int16_t* host_bus_out) {
    BLAST_REQUIRE(g && (tracks || n_tracks == 0) && (voices || n_voices == 0) && (host_bus_out || frames == 0), BLAST_ERR_ARG,
                  "blast_group_render: null argument");
    for (uint32_t v = 0; v < n_voices; ++v)
        if (voices[v].track >= n_tracks) return blast::set_error(BLAST_ERR_REF_PANIC, "voice %u: track index %u out of bounds (reference panics)", v, voices[v].track);
    const uint64_t slots = frames * out_channels;
    if (slots == 0) return BLAST_OK;
    if (int rc = ensure_peer_buses(g, slots)) return rc;
    // every member builds a scene of the voices whose track it holds (track t lives on member t mod n); the launches are
    // asynchronous, so one host thread feeds all GPUs
    std::vector<blast_scene*> sc(g->n, nullptr);
    int rc = BLAST_OK;
    for (uint32_t m = 0; m < g->n && rc == BLAST_OK; ++m) {
        std::vector<blast_voice> mine;
        for (uint32_t v = 0; v < n_voices; ++v)
            if (voices[v].track % g->n == m) mine.push_back(voices[v]);
        rc = blast_scene_create(g->ctx[m], tracks, n_tracks, mine.data(), (uint32_t)mine.size(), out_channels, &sc[m]);
        // everything the render allocates, now: members may share a GPU, and an allocation behind a kernel that waits for
        // a peer may never return
        if (rc == BLAST_OK) rc = blast_scene_reserve(g->ctx[m], sc[m], frames);
    }
    // from here on every member takes its step of the peer protocol, whatever happens to another one
    for (uint32_t m = 0; m < g->n; ++m) {
        int r = BLAST_OK;
        if (rc == BLAST_OK && sc[m]) r = blast_scene_render_reduce_dev(g->ctx[m], sc[m], frames, g->pb[m]);
        else r = blast_peer_bus_reduce_dev(g->ctx[m], g->pb[m], slots);
        if (r != BLAST_OK && rc == BLAST_OK) rc = r;
    }
    if (rc == BLAST_OK) rc = blast_peer_bus_wait_dev(g->ctx[0], g->pb[0]);
    if (rc == BLAST_OK) rc = blast_memcpy_d2h(g->ctx[0], host_bus_out, blast_peer_bus_bus(g->pb[0]), slots * sizeof(int16_t));
    for (uint32_t m = 0; m < g->n; ++m) {
        int r = blast_peer_bus_check(g->ctx[m], g->pb[m]);
        if (r == BLAST_OK && sc[m]) r = blast_scene_check(g->ctx[m], sc[m]);
        if (r != BLAST_OK && rc == BLAST_OK) rc = r;
    }
    const std::string keep = rc != BLAST_OK ? blast_last_error() : "";
    for (uint32_t m = 0; m < g->n; ++m)
        if (sc[m]) blast_scene_destroy(g->ctx[m], sc[m]);
    if (rc != BLAST_OK) return blast::set_error(rc, "%s", keep.c_str());
    return BLAST_OK;
}

int blast_group_conductor_create(blast_group* g, uint32_t out_channels, uint32_t sample_rate, const blast_track* tracks,
                                 uint32_t n_tracks, blast_group_conductor** out) {
    BLAST_REQUIRE(g && out, BLAST_ERR_ARG, "blast_group_conductor_create: null argument");
    *out = nullptr;
    auto* gc = new blast_group_conductor();
    gc->g = g;
    gc->out_channels = out_channels;
    for (uint32_t m = 0; m < g->n; ++m) {
        blast_conductor* c = nullptr;
        int rc = blast_conductor_create(g->ctx[m], out_channels, sample_rate, tracks, n_tracks, &c);
        if (rc == BLAST_OK) rc = blast_conductor_set_shard_by_track(c, m, g->n);
        if (rc != BLAST_OK) {
            if (c) blast_conductor_destroy(g->ctx[m], c);
            blast_group_conductor_destroy(gc);
            return rc;
        }
        gc->c.push_back(c);
    }
    *out = gc;
    return BLAST_OK;
}

void blast_group_conductor_destroy(blast_group_conductor* gc) {
    if (!gc) return;
    for (uint32_t m = 0; m < gc->c.size(); ++m) blast_conductor_destroy(gc->g->ctx[m], gc->c[m]);
    delete gc;
}

int blast_group_conductor_apply(blast_group_conductor* gc, const blast_command* cmd) {
    BLAST_REQUIRE(gc && cmd, BLAST_ERR_ARG, "blast_group_conductor_apply: null argument");
    // every member applies every command (host state machine, validated before it mutates: all members agree)
    int rc = BLAST_OK;
    for (uint32_t m = 0; m < gc->c.size(); ++m) {
        const int r = blast_conductor_apply(gc->g->ctx[m], gc->c[m], cmd);
        if (m == 0) rc = r;
        else if (r != rc) return blast::set_error(BLAST_ERR_CUDA, "group conductor: members disagree on a command (%d vs %d)", rc, r);
    }
    return rc;
}

blast_conductor* blast_group_conductor_member(blast_group_conductor* gc, uint32_t member) {
    return (gc && member < gc->c.size()) ? gc->c[member] : nullptr;
}

int blast_group_conductor_coordinate(blast_group_conductor* gc, uint64_t frames, int16_t* host_bus_out) {
    BLAST_REQUIRE(gc && (host_bus_out || frames == 0), BLAST_ERR_ARG, "blast_group_conductor_coordinate: null argument");
    blast_group* g = gc->g;
    const uint64_t slots = frames * gc->out_channels;
    if (slots == 0) return BLAST_OK;
    if (int rc = ensure_peer_buses(g, slots)) return rc;
    // Phase 1, one host thread per member: the span is rendered into the member's partial bus.  Everything that may
    // allocate or synchronise (the Conductor reads its state back) happens here, while no kernel of the group waits for a
    // peer — members may share a GPU, and a device-wide synchronisation behind a waiting kernel would never return.
    std::vector<int> rcs(g->n, BLAST_OK);
    std::vector<std::string> msgs(g->n);
    for_members(g, [&](uint32_t m) -> int {
        int rc = blast_peer_bus_begin_dev(g->ctx[m], g->pb[m]);
        if (rc == BLAST_OK) rc = blast_conductor_render_dev(g->ctx[m], gc->c[m], frames, blast_peer_bus_partial(g->pb[m]));
        rcs[m] = rc;
        if (rc != BLAST_OK) msgs[m] = blast_last_error();
        return BLAST_OK;
    });
    // Phase 2: every member takes its step of the exchange, also one that failed (capacity, allocation): the others must
    // not wait for it.  Asynchronous launches from this thread; nothing blocks between them.
    int rc = BLAST_OK;
    for (uint32_t m = 0; m < g->n; ++m) {
        const int r = blast_peer_bus_reduce_dev(g->ctx[m], g->pb[m], slots);
        if (r != BLAST_OK && rc == BLAST_OK) rc = r;
    }
    if (rc == BLAST_OK) rc = blast_peer_bus_wait_dev(g->ctx[0], g->pb[0]);
    if (rc == BLAST_OK) rc = blast_memcpy_d2h(g->ctx[0], host_bus_out, blast_peer_bus_bus(g->pb[0]), slots * sizeof(int16_t));
    for (uint32_t m = 0; m < g->n; ++m) {
        const int r = blast_peer_bus_check(g->ctx[m], g->pb[m]);
        if (r != BLAST_OK && rc == BLAST_OK) rc = r;
    }
    for (uint32_t m = 0; m < g->n; ++m)
        if (rcs[m] != BLAST_OK) return blast::set_error(rcs[m], "member %u (GPU %d): %s", m, g->device[m], msgs[m].c_str());
    return rc;
}

int blast_group_x128p_fill(blast_group* g, uint64_t seed, uint64_t stride, uint64_t n_streams, uint64_t draws_per_stream,
                           int64_t lower, int64_t upper, uint64_t* raw_out, int64_t* ranged_out, uint64_t* checks_out) {
    BLAST_REQUIRE(g != nullptr, BLAST_ERR_ARG, "blast_group_x128p_fill: null group");
    if (n_streams == 0) return BLAST_OK;
    blast_x128p base;
    blast_x128p_seed(seed, &base);
    return for_members(g, [&](uint32_t m) -> int {
        blast_ctx* ctx = g->ctx[m];
        if (int rc = blast::bind(ctx)) return rc;
        if (m >= n_streams) return BLAST_OK;
        const uint64_t mine = (n_streams - m + g->n - 1) / g->n;        // streams m, m + n, m + 2n, ...
        blast_x128p b = base;
        // stream s starts s * stride draws into the sequence: member m's first stream is advanced m times by `stride`
        for (uint32_t k = 0; k < m; ++k) {
            blast_x128p t;
            if (int rc = blast_x128p_advance(&b, stride, &t)) return rc;
            b = t;
        }
        const uint64_t row = draws_per_stream * sizeof(uint64_t);
        void *d_st = nullptr, *d_raw = nullptr, *d_rng = nullptr, *d_chk = nullptr;
        auto done = [&](int rc) {
            cudaStreamSynchronize(ctx->stream);
            if (d_st) cudaFree(d_st);
            if (d_raw) cudaFree(d_raw);
            if (d_rng) cudaFree(d_rng);
            if (d_chk) cudaFree(d_chk);
            return rc;
        };
        BLAST_CUDA_TRY(cudaMalloc(&d_st, mine * sizeof(blast_x128p)));
        if (raw_out && cudaMalloc(&d_raw, std::max<uint64_t>(mine * row, 16)) != cudaSuccess) return done(blast::set_error(BLAST_ERR_CUDA, "blast_group_x128p_fill: cudaMalloc failed"));
        if (ranged_out && cudaMalloc(&d_rng, std::max<uint64_t>(mine * row, 16)) != cudaSuccess) return done(blast::set_error(BLAST_ERR_CUDA, "blast_group_x128p_fill: cudaMalloc failed"));
        if (checks_out && cudaMalloc(&d_chk, mine * 32) != cudaSuccess) return done(blast::set_error(BLAST_ERR_CUDA, "blast_group_x128p_fill: cudaMalloc failed"));
        // (stride * n wraps mod 2^64 exactly like the sequence index does: the generator's period is 2^128 - 1 and the jump
        // by a 64-bit count is what blast_x128p_jump_dev takes)
        if (int rc = blast_x128p_jump_dev(ctx, &b, stride * g->n, mine, static_cast<blast_x128p*>(d_st))) return done(rc);
        if (int rc = blast_x128p_fill_dev(ctx, static_cast<blast_x128p*>(d_st), mine, draws_per_stream, lower, upper,
                                          static_cast<uint64_t*>(d_raw), static_cast<int64_t*>(d_rng), static_cast<uint64_t*>(d_chk)))
            return done(rc);
        const size_t pitch = (size_t)g->n * row;
        if (raw_out && row)
            if (cudaMemcpy2DAsync(raw_out + (size_t)m * draws_per_stream, pitch, d_raw, row, row, mine, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                return done(blast::set_error(BLAST_ERR_CUDA, "blast_group_x128p_fill: D2H failed"));
        if (ranged_out && row)
            if (cudaMemcpy2DAsync(ranged_out + (size_t)m * draws_per_stream, pitch, d_rng, row, row, mine, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                return done(blast::set_error(BLAST_ERR_CUDA, "blast_group_x128p_fill: D2H failed"));
        if (checks_out)
            if (cudaMemcpy2DAsync(checks_out + (size_t)m * 4, (size_t)g->n * 32, d_chk, 32, 32, mine, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                return done(blast::set_error(BLAST_ERR_CUDA, "blast_group_x128p_fill: D2H failed"));
        return done(BLAST_OK);
    });
}

int blast_group_mpeg_index(blast_group* g, const uint8_t* bytes, uint64_t len, int reference_compat, uint64_t* offsets_out,
                           uint64_t cap, uint64_t* n_offsets_out, uint32_t* ref_header_out, uint64_t* n_candidates_out) {
    BLAST_REQUIRE(g && (bytes || len == 0) && n_offsets_out, BLAST_ERR_ARG, "blast_group_mpeg_index: null argument");
    const uint32_t n = g->n;
    // contiguous byte ranges, all but the last a multiple of the scan's 32 KiB span, each followed by 16 halo bytes
    const uint64_t spans = (len + kMpegSpan - 1) / kMpegSpan, per = (spans + n - 1) / n;
    struct Range { uint64_t start, own, halo; };
    std::vector<Range> rg(n);
    for (uint32_t m = 0; m < n; ++m) {
        const uint64_t a = std::min(len, (uint64_t)m * per * kMpegSpan), b = std::min(len, (uint64_t)(m + 1) * per * kMpegSpan);
        rg[m] = Range{a, b - a, b < len ? std::min(kMpegHalo, len - b) : 0};
    }
    std::vector<void*> d_bytes(n, nullptr), d_pos(n, nullptr), d_hdr(n, nullptr), d_hist(n, nullptr), d_first(n, nullptr), d_off(n, nullptr);
    std::vector<blast_mpeg_shard_agg> agg(n);
    std::vector<uint64_t> count(n, 0), n_off(n, 0);
    std::vector<uint32_t> entry(n, 0);
    auto cleanup = [&](int rc) {
        const std::string keep = rc != BLAST_OK ? blast_last_error() : "";
        for (uint32_t m = 0; m < n; ++m) {
            cudaSetDevice(g->device[m]);
            cudaStreamSynchronize(g->ctx[m]->stream);
            for (void* p : {d_bytes[m], d_pos[m], d_hdr[m], d_hist[m], d_first[m], d_off[m]})
                if (p) cudaFree(p);
        }
        if (rc != BLAST_OK) return blast::set_error(rc, "%s", keep.c_str());
        return rc;
    };
    // phase 1: upload + walk: every range's action on the scan's 4-state machine
    int rc = for_members(g, [&](uint32_t m) -> int {
        blast_ctx* ctx = g->ctx[m];
        if (int r = blast::bind(ctx)) return r;
        for (int s = 0; s < 4; ++s) { agg[m].exit_state[s] = (uint32_t)s; agg[m].count[s] = 0; }
        if (rg[m].own == 0) return BLAST_OK;
        BLAST_CUDA_TRY(cudaMalloc(&d_bytes[m], rg[m].own + rg[m].halo + 256));
        BLAST_CUDA_TRY(cudaMemcpyAsync(d_bytes[m], bytes + rg[m].start, rg[m].own + rg[m].halo, cudaMemcpyHostToDevice, ctx->stream));
        return blast_mpeg_shard_walk_dev(ctx, static_cast<const uint8_t*>(d_bytes[m]), rg[m].own, rg[m].halo, &agg[m]);
    });
    if (rc != BLAST_OK) return cleanup(rc);
    // the exchange: fold the range aggregates in order (range 0 enters in state 0)
    uint32_t state = 0;
    uint64_t total = 0;
    for (uint32_t m = 0; m < n; ++m) {
        entry[m] = state;
        count[m] = agg[m].count[state];
        total += count[m];
        state = agg[m].exit_state[state];
    }
    if (n_candidates_out) *n_candidates_out = total;
    // phase 2: emit with global positions, histogram of the headers
    rc = for_members(g, [&](uint32_t m) -> int {
        blast_ctx* ctx = g->ctx[m];
        if (int r = blast::bind(ctx)) return r;
        BLAST_CUDA_TRY(cudaMalloc(&d_hist[m], (size_t)BLAST_MPEG_HDR_BINS * sizeof(uint32_t)));
        BLAST_CUDA_TRY(cudaMemsetAsync(d_hist[m], 0, (size_t)BLAST_MPEG_HDR_BINS * sizeof(uint32_t), ctx->stream));
        if (rg[m].own) {
            BLAST_CUDA_TRY(cudaMalloc(&d_pos[m], std::max<uint64_t>(count[m], 2) * sizeof(uint64_t)));
            BLAST_CUDA_TRY(cudaMalloc(&d_hdr[m], std::max<uint64_t>(count[m], 4) * sizeof(uint32_t)));
            uint64_t got = 0;
            if (int r = blast_mpeg_shard_emit_dev(ctx, static_cast<const uint8_t*>(d_bytes[m]), rg[m].own, rg[m].halo, entry[m], rg[m].start,
                                                  static_cast<uint64_t*>(d_pos[m]), static_cast<uint32_t*>(d_hdr[m]), count[m], &got))
                return r;
            if (got != count[m]) return blast::set_error(BLAST_ERR_CUDA, "range %u emitted %llu candidates, its aggregate said %llu", m,
                                                          (unsigned long long)got, (unsigned long long)count[m]);
            if (count[m])
                if (int r = blast_mpeg_hist_dev(ctx, static_cast<const uint32_t*>(d_hdr[m]), count[m], static_cast<uint32_t*>(d_hist[m]))) return r;
        }
        return blast_ctx_sync(ctx);
    });
    if (rc != BLAST_OK) return cleanup(rc);
    // header vote: member 0 sums the histograms through peer memory and picks the reference header
    blast_ctx* c0 = g->ctx[0];
    if ((rc = blast::bind(c0)) != BLAST_OK) return cleanup(rc);
    for (uint32_t m = 1; m < n; ++m) {
        add_u32_from<<<c0->sm_count * 4, 256, 0, c0->stream>>>(static_cast<uint32_t*>(d_hist[0]), static_cast<const uint32_t*>(d_hist[m]), BLAST_MPEG_HDR_BINS);
        c0->launches += 1;
    }
    uint32_t ref = 0;
    if ((rc = blast_mpeg_pick_ref_dev(c0, static_cast<const uint32_t*>(d_hist[0]), &ref)) != BLAST_OK) return cleanup(rc);
    if (ref_header_out) *ref_header_out = ref;
    // duplicate-first quirk (mpeg.rs:39): the first position of every header value, min-reduced onto member 0
    if (reference_compat) {
        rc = for_members(g, [&](uint32_t m) -> int {
            blast_ctx* ctx = g->ctx[m];
            if (int r = blast::bind(ctx)) return r;
            BLAST_CUDA_TRY(cudaMalloc(&d_first[m], (size_t)BLAST_MPEG_HDR_BINS * sizeof(uint64_t)));
            BLAST_CUDA_TRY(cudaMemsetAsync(d_first[m], 0xFF, (size_t)BLAST_MPEG_HDR_BINS * sizeof(uint64_t), ctx->stream));
            if (count[m])
                if (int r = blast_mpeg_first_pos_dev(ctx, static_cast<const uint64_t*>(d_pos[m]), static_cast<const uint32_t*>(d_hdr[m]), count[m], ref,
                                                     static_cast<uint64_t*>(d_first[m])))
                    return r;
            return blast_ctx_sync(ctx);
        });
        if (rc != BLAST_OK) return cleanup(rc);
        if ((rc = blast::bind(c0)) != BLAST_OK) return cleanup(rc);
        for (uint32_t m = 1; m < n; ++m) {
            min_u64_from<<<c0->sm_count * 4, 256, 0, c0->stream>>>(static_cast<unsigned long long*>(d_first[0]),
                                                                  static_cast<const unsigned long long*>(d_first[m]), BLAST_MPEG_HDR_BINS);
            c0->launches += 1;
        }
        if ((rc = blast_ctx_sync(c0)) != BLAST_OK) return cleanup(rc);
    }
    // frames of every range (the merged first-position table is read from member 0's memory), then the ordered index
    rc = for_members(g, [&](uint32_t m) -> int {
        blast_ctx* ctx = g->ctx[m];
        if (int r = blast::bind(ctx)) return r;
        if (count[m] == 0) return BLAST_OK;
        const uint64_t* first = reference_compat ? static_cast<const uint64_t*>(d_first[0]) : nullptr;
        uint64_t k = 0;
        int r = blast_mpeg_classify_dev(ctx, static_cast<const uint64_t*>(d_pos[m]), static_cast<const uint32_t*>(d_hdr[m]), count[m], ref, first, len,
                                        nullptr, 0, &k);
        if (r != BLAST_OK) return r;
        n_off[m] = k;
        if (k == 0 || !offsets_out) return BLAST_OK;
        BLAST_CUDA_TRY(cudaMalloc(&d_off[m], k * sizeof(uint64_t)));
        return blast_mpeg_classify_dev(ctx, static_cast<const uint64_t*>(d_pos[m]), static_cast<const uint32_t*>(d_hdr[m]), count[m], ref, first, len,
                                       static_cast<uint64_t*>(d_off[m]), k, &k);
    });
    if (rc != BLAST_OK) return cleanup(rc);
    uint64_t n_total = 0;
    for (uint32_t m = 0; m < n; ++m) n_total += n_off[m];
    *n_offsets_out = n_total;
    if (offsets_out) {
        if (n_total > cap) return cleanup(blast::set_error(BLAST_ERR_CAPACITY, "%llu frame offsets, room for %llu", (unsigned long long)n_total, (unsigned long long)cap));
        uint64_t at = 0;
        for (uint32_t m = 0; m < n; ++m) {
            if (n_off[m]) {
                cudaSetDevice(g->device[m]);
                if (cudaMemcpyAsync(offsets_out + at, d_off[m], n_off[m] * sizeof(uint64_t), cudaMemcpyDeviceToHost, g->ctx[m]->stream) != cudaSuccess)
                    return cleanup(blast::set_error(BLAST_ERR_CUDA, "blast_group_mpeg_index: D2H failed"));
            }
            at += n_off[m];
        }
    }
    return cleanup(BLAST_OK);
}

}  // extern "C"
