// render_internal.h — structures shared by render.cu (kernels + scene API) and conductor.cu (host state machine)
#pragma once
#include "blast_internal.h"

namespace blast_rdr {

constexpr int kMaxSeg = 160;                   // position segments per voice per render (plain voices: <= ~45 are ever needed)
constexpr int kMaxSegSeq = 2048;               // ... when Seq processes retrigger voices
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
    uint32_t adv;          // 0: every step is an advance event.  Else (voices with Seq processes) steps are
                           // calls: oc | lo << 8 | na << 16 = calls per frame, first advancing channel, count
    uint32_t first_seq;    // index of the voice's first SeqDev | n_seqs << 24 (0 seqs for plain voices)
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
    uint32_t meta;         // [15:0] steps this segment still covers (saturating), [31:16] segment index
};


constexpr int kMaxEvents = 256;        // retrigger events per voice per render call
// rows of a voice's piece-table pool: one header row per staged item + one row per piece; a tile that does not fit is
// left to K4's own (slower) cutting
constexpr uint32_t pool_rows_for(uint32_t seg_cap) { return 3u * seg_cap; }
constexpr uint32_t kTabTag = 0x7FC00000u;      // TileRec.scale of a tile with a table: this quiet-NaN pattern | item count

struct Split {             // a retrigger between the channels of one frame (C >= 2 voices): channels < k keep the old sample
    uint32_t frame;
    uint32_t k;
    float p_old;           // position of the trajectory the channels < k have read
    uint32_t pad;
};

struct SeqDev {            // one Seq process (processes.rs:52-99) flattened for the GPU
    uint32_t base;         // tempo.current as the Seq sees it at call 0 of this render
    uint32_t rate;         // ticks of its tempo per call
    float interval;        // TempoState.interval (samples)
    float period_f;        // period as f32
    uint32_t n_steps;
    uint32_t idx;          // SeqState.idx (in/out)
    const float* steps;    // device
    const float* chance;   // device
    unsigned long long s0, s1;   // X128P state (in/out)
};

struct RenderBuffers {     // device scratch of one render (owned by a scene or a conductor); grow-only
    VoiceDev* d_voices = nullptr;
    Seg* d_segs = nullptr;
    uint32_t* d_nsegs = nullptr;
    uint32_t* d_err = nullptr;         // [0], [1]: error words of alternate renders (bit 0: a trajectory needed more segments
                                       // than seg_cap, bit 1: > kMaxEvents retriggers); [2]: K4's work-item counter.  The
                                       // render with parity p reports in d_err[p]; its K3 clears d_err[p ^ 1] and the counter,
                                       // so no memset sits between the kernels of a render.
    uint32_t parity = 0;               // of the last launch_render: its error word is d_err[parity]
    uint32_t* err_word() const { return d_err + parity; }
    TileRec* d_recs = nullptr;
    size_t recs_cap = 0;               // in records
    uint4* d_pool = nullptr;           // piece tables of the tiles whose trajectory has several segments: kPoolRows(seg_cap)
    size_t pool_voices = 0;            // rows per voice, written by K3, bulk-copied into K4's shared memory
    uint32_t pool_rows = 0;
    size_t voices_cap = 0;
    uint32_t seg_cap = 0;              // segments per voice in d_segs
    SeqDev* d_seqs = nullptr;          // only with Seq processes
    size_t seqs_cap = 0;
    uint32_t* d_events = nullptr;      // [voices_cap][kMaxEvents] retrigger call indices
    uint32_t* d_nevents = nullptr;     // [voices_cap]
    Split* d_splits = nullptr;         // [voices_cap][kMaxEvents]
    uint32_t* d_nsplits = nullptr;     // [voices_cap]
};

// Where finished bus tiles go when the render feeds a peer bus (blast_peer_bus, peer_bus.cu): the render kernel itself
// publishes every completed tile to the rank that owns it (tile t -> rank t mod world), and takes "reduce items" from
// the same work queue: wait for the tile's ready flags of all ranks, sum the int32 partial tiles over peer memory
// (NVLink / NVSwitch loads), wrap to S16 and store into the root's bus.  world == 1 is the single-GPU finalize.
constexpr int kMaxPeers = 16;
struct BusSink {
    uint32_t world = 0;                    // 0: no sink, the render leaves int32 partial sums only
    uint32_t rank = 0;
    uint32_t step = 0;                     // value the flags of this render carry (counters compared with wrap-around)
    uint32_t lag = 1;                      // tile t is reduced by the work item that renders tile t + lag (>= 1)
    uint32_t max_tiles = 0;                // row length of the ready tables
    uint32_t n_tiles = 0;                  // tiles of this render
    uint32_t n_my_tiles = 0;               // ... of which this rank reduces
    uint32_t tile_slots = 0;               // bus slots per tile (kFT * out_channels inside the render kernel)
    uint64_t n_slots = 0;                  // bus slots of this render
    uint32_t n_done = 0;
    uint32_t timeout_ms = 0;
    const int32_t* part[kMaxPeers] = {};   // slot 0 of every rank's partial bus (part[rank] is local), mapped here
    uint32_t* ready_at[kMaxPeers] = {};    // ready_at[o] = this rank's row of owner o's ready table
    const uint32_t* ready_mine = nullptr;  // this rank's table [world][max_tiles], written by the peers
    int16_t* out = nullptr;                // slot 0 of the ROOT's S16 bus, mapped here
    uint32_t* tile_count = nullptr;        // local [max_tiles]: flushed voice groups per tile
    uint32_t* red_count = nullptr;         // local: tiles reduced so far in this step
    uint32_t* done[2 * kMaxPeers] = {};    // flags that receive `step` when this rank has reduced all its tiles
    const uint32_t* ack_mine = nullptr;    // this rank's ack flags [world]: rank p is done reading my partial bus
    uint32_t* err = nullptr;               // local: bit 2 = a flag wait timed out
};

int  reserve_buffers(blast_ctx* ctx, RenderBuffers& rb, uint32_t n_voices, uint32_t n_seqs);
int  reserve_frames(blast_ctx* ctx, RenderBuffers& rb, uint32_t n_voices, uint32_t out_channels, uint64_t frames);
int  preload_kernels(blast_ctx* ctx);
void free_buffers(RenderBuffers& rb);
// Seq event scan (when n_seqs > 0) + position scan + render/mix of the first n_voices records of rb.d_voices into
// the int32 partial bus (overwritten); async on ctx->stream.  Device-side capacity errors land in rb.d_err.
// sink != nullptr: the partial bus is sink->part[sink->rank] and finished tiles are reduced into sink->out (fused into the
// render kernel for out_channels <= 2, as two small kernels after it otherwise).
// rewind != nullptr: the voice rows are first copied from there (by the position scan itself).
int launch_render(blast_ctx* ctx, RenderBuffers& rb, uint32_t n_voices, uint32_t n_seqs, uint32_t out_channels,
                  uint64_t frames, int32_t* d_partial_bus, const BusSink* sink = nullptr, const VoiceDev* rewind = nullptr);
// the reduction on its own: publish every tile of this rank's partial bus (filled by earlier stream work), reduce the
// tiles this rank owns.  Async on ctx->stream.
int launch_bus_reduce(blast_ctx* ctx, const BusSink& sink, cudaStream_t stream = nullptr);   // nullptr: ctx->stream
// async: the stream waits (on the device, bounded) until the n flags have reached `value`
int launch_flag_wait(blast_ctx* ctx, const uint32_t* d_flags, uint32_t n, uint32_t value, uint32_t timeout_ms, uint32_t* d_err,
                     cudaStream_t stream = nullptr);
// peer_bus.cu: the exchange of the step `sink` describes, on the peer bus's own stream behind everything enqueued on
// ctx->stream so far — it overlaps whatever the caller enqueues next (the next batch's decode); blast_peer_bus_wait_dev
// joins it.
int peer_bus_exchange(blast_ctx* ctx, blast_peer_bus* pb, const BusSink& sink);
// peer_bus.cu: starts the next step of a peer bus for a render of `frames` frames on an out_channels bus and fills the
// sink the kernels take.  in_render: tiles are the render kernel's (kFT frames); else 4,096-slot tiles.
int peer_bus_next_step(blast_ctx* ctx, blast_peer_bus* pb, uint64_t frames, uint32_t out_channels, bool in_render, BusSink* out);
int32_t* peer_bus_partial(blast_peer_bus* pb);
bool peer_bus_fused(const blast_peer_bus* pb);
// VoiceDev routing fields (S, nch, adv) for a voice with C channels on an out_channels bus
void route_voice(VoiceDev& v, uint32_t out_channels, bool has_seq);

}  // namespace blast_rdr
