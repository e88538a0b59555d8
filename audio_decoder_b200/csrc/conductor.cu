// conductor.cu — the Command-driven control state of the render path, on top of render.cu's kernels.
//
// Replaces (paths relative to the reference root, blast/src/audio_processing/):
//   engine.rs:36-44     Conductor::prepare
//   engine.rs:83-248    Conductor::apply and its handlers (load/start/pause/resume/stop/unload/velocity/group/tc/seq)
//   engine.rs:252-275   tempo_from_repr
//   engine.rs:318-384   Voice::{start,pause,resume,stop}
//   engine.rs:477-528   Group::{start,pause,resume,stop}
//   engine.rs:46-81     Conductor::coordinate (the frame -> channel -> voices -> groups loop)
//   blast_time.rs:58-161 TempoState, convert_interval
//   processes.rs:52-99  Seq
//
// Design.  Between two commands nothing but time changes a voice's `active` flag, velocity, gain or the set of
// objects that tick a tempo, so one coordinate() span is a static scene whose only sequential pieces are
//   (1) tempo counters: `current += 1` by every ticker, once per CALL (frame x channel): a pure function of the
//       call index — current(c) = base + rate * c, with `base` including the ticks of earlier voices in the same
//       call (engine.rs:392-405 order) — so the host only carries (base, rate) per Seq and adds rate * calls
//       afterwards;
//   (2) Seq events: exact-equality hits of fmodf(current / interval, period) against steps[idx], each consuming
//       one xoroshiro128+ draw (K3a `seq_event_scan`, one thread per voice, bisection over the monotone tempo);
//   (3) the f32 position recurrence, restarted at every retrigger (K3 `voice_position_scan`).
// The host keeps the reference's object graph (voices, groups, shared tempi) and flattens it per span into the
// VoiceDev / SeqDev tables; samples, positions, Seq hits and RNG draws are all computed on the GPU (K3a, K3, K4).
// A span is cut into chunks when a voice would need more than kMaxSeg position segments or kMaxEvents retriggers:
// the chunk is re-run at half the length from the (host-authoritative) state, nothing is committed on overflow.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <numeric>
#include <unordered_map>
#include <vector>

#include "render_internal.h"

using namespace blast_rdr;

namespace {

struct Tempo {                 // blast_time.rs:58-65
    uint32_t mode = BLAST_TM_TBD;
    uint32_t unit = BLAST_TU_SAMPLES;
    float interval = 0.0f;
    bool active = false;
    uint32_t current = 0;
    // flatten()'s bookkeeping (not reference state): ticks per call booked in the flatten pass `flat_epoch`
    uint64_t flat_epoch = 0;
    uint32_t flat_ticks = 0;
    void reset() { current = 0; }                          // blast_time.rs:141-143
    void start() { reset(); active = true; }               // blast_time.rs:123-126
    void stop() { active = false; reset(); }               // blast_time.rs:136-139
};
using TempoRef = std::shared_ptr<Tempo>;

struct SeqH {                  // processes.rs:56-65
    bool active = true;
    TempoRef tempo;
    uint64_t period = 0;
    std::vector<float> steps, chance;
    blast_x128p rng{0, 0};
    uint64_t idx = 0;
};

struct VoiceH {                // engine.rs:279-295
    uint64_t uid = 0;          // load order number (sharding key)
    uint32_t track = 0;        // LoadArgs.track_idx (sharding key when the tracks themselves are sharded over GPUs)
    const int16_t* d_smp = nullptr;
    uint64_t end = 0;
    uint32_t C = 0;
    bool active = false;
    float pos = 0.0f, vel = 1.0f, gain = 1.0f;
    TempoRef tempo;
    std::vector<SeqH> procs;
    std::vector<TempoRef> proc_tempi;

    float home() const { return vel >= 0.0f ? 0.0f : (float)end; }       // engine.rs:340-343
    void start() {             // engine.rs:318-344
        active = true;
        for (auto& p : procs) p.idx = 0;
        if (tempo->mode == BLAST_TM_VOICE || tempo->mode == BLAST_TM_TBD) tempo->start();
        for (auto& t : proc_tempi) t->start();
        pos = home();
    }
    void pause() { active = false; }                       // engine.rs:346-348
    void resume() { active = true; }                       // engine.rs:350-359
    void stop() {              // engine.rs:361-384
        active = false;
        for (auto& p : procs) p.idx = 0;
        if (tempo->mode == BLAST_TM_VOICE) tempo->stop();
        for (auto& t : proc_tempi) { t->active = false; t->reset(); }
        pos = home();
    }
};

struct GroupH {                // engine.rs:451-461
    bool active = false;
    float gain = 1.0f;         // never read by the reference (engine.rs:453,467)
    TempoRef tempo;
    std::vector<VoiceH> voices;
    std::vector<SeqH> processes;   // stored, never run (engine.rs:530-542)
    void start() {             // engine.rs:477-497
        active = true;
        if (tempo->mode == BLAST_TM_GROUP) { tempo->active = true; tempo->reset(); }
        for (auto& v : voices) v.start();
    }
    void pause() { active = false; }                       // engine.rs:499-503
    void resume() { active = true; }                       // engine.rs:505-514
    void stop() {              // engine.rs:516-528
        active = false;
        for (auto& v : voices) v.active = false;
        if (tempo->mode == BLAST_TM_GROUP) { tempo->active = false; tempo->reset(); }
    }
};

float convert_interval(uint32_t sample_rate, uint32_t unit, float interval) {   // blast_time.rs:151-161
    float frac;
    switch (unit) {
        case BLAST_TU_MILLIS: frac = interval / 1000.0f; break;
        case BLAST_TU_BPM: frac = 60.0f / interval; break;
        default: return interval;
    }
    return (float)sample_rate * frac;
}

}  // namespace

struct blast_conductor {
    std::vector<VoiceH> voices;
    std::vector<GroupH> groups;
    std::vector<TempoRef> tempo_cons;
    uint32_t out_channels = 0;
    uint32_t sample_rate = 0;
    std::vector<blast_track> tracks;
    uint64_t clock = 0;
    uint64_t flat_epoch = 0;               // flatten passes so far (Tempo::flat_epoch)
    uint64_t next_uid = 0;
    uint32_t rank = 0, world = 1;
    bool shard_by_track = false;       // a voice is rendered where its track lives (track t on rank t mod world)
    RenderBuffers rb;
    float* d_fpool = nullptr;          // steps / chance of the live Seqs
    size_t fpool_cap = 0;
    // pinned staging (grow-only)
    void* h_pin = nullptr;
    size_t pin_cap = 0;

    TempoRef tempo_default() const {   // TempoState::new(None), blast_time.rs:84-97
        auto t = std::make_shared<Tempo>();
        t->interval = (float)sample_rate;
        return t;
    }
    // engine.rs:252-275; nullptr where the reference indexes out of bounds
    TempoRef tempo_from_repr(const blast_tempo_repr& tr) const {
        TempoRef t = tempo_default();
        if (tr.owned) {
            t->interval = convert_interval(sample_rate, tr.unit, tr.interval);    // TempoState::init, blast_time.rs:99-104
            t->mode = tr.mode;
            t->unit = tr.unit;
            return t;
        }
        switch (tr.mode) {
            case BLAST_TM_VOICE: return tr.idx < voices.size() ? voices[tr.idx].tempo : nullptr;
            case BLAST_TM_GROUP: return tr.idx < groups.size() ? groups[tr.idx].tempo : nullptr;
            case BLAST_TM_CONTEXT: return tr.idx < tempo_cons.size() ? tempo_cons[tr.idx] : nullptr;
            default: return t;
        }
    }
};

namespace {

const char* kOob = "index out of bounds (the reference panics)";

int apply_command(blast_conductor* c, const blast_command* cmd) {
    switch (cmd->kind) {
        case BLAST_CMD_LOAD: {                                  // engine.rs:103-107, 298-316
            if (cmd->idx >= c->tracks.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "load: track %s", kOob);
            if (cmd->tempo.mode > BLAST_TM_TBD || cmd->tempo.unit > BLAST_TU_BPM) return blast::set_error(BLAST_ERR_ARG, "load: bad tempo mode / unit");
            const blast_track& tr = c->tracks[cmd->idx];
            TempoRef tempo = c->tempo_from_repr(cmd->tempo);
            if (!tempo) return blast::set_error(BLAST_ERR_REF_PANIC, "load: tempo %s", kOob);
            if (tr.num_channels == 0) return blast::set_error(BLAST_ERR_REF_PANIC, "load: track has 0 channels (the reference divides by zero)");
            const uint64_t frames = tr.n_samples / tr.num_channels;
            if (frames == 0) return blast::set_error(BLAST_ERR_REF_PANIC, "load: empty track (usize underflow in Voice::new, engine.rs:302)");
            if (frames - 1 > 0x7FFFFFFFull) return blast::set_error(BLAST_ERR_CAPACITY, "load: track longer than 2^31 frames");
            if (((uintptr_t)tr.d_samples & 3) != 0) return blast::set_error(BLAST_ERR_ARG, "load: track samples must be 4-byte aligned");
            VoiceH v;
            v.uid = c->next_uid++;
            v.track = (uint32_t)cmd->idx;
            v.d_smp = tr.d_samples;
            v.end = frames - 1;
            v.C = tr.num_channels;
            v.tempo = tempo;
            c->voices.push_back(std::move(v));
            return BLAST_OK;
        }
        case BLAST_CMD_START: case BLAST_CMD_PAUSE: case BLAST_CMD_RESUME: case BLAST_CMD_STOP: {   // engine.rs:110-180
            const uint32_t k = cmd->kind;
            if (cmd->idx_kind == BLAST_IDX_VOICE) {
                if (cmd->idx >= c->voices.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "voice %s", kOob);
                VoiceH& v = c->voices[cmd->idx];
                if (k == BLAST_CMD_START) v.start(); else if (k == BLAST_CMD_PAUSE) v.pause();
                else if (k == BLAST_CMD_RESUME) v.resume(); else v.stop();
            } else if (cmd->idx_kind == BLAST_IDX_GROUP) {
                if (cmd->idx >= c->groups.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "group %s", kOob);
                GroupH& g = c->groups[cmd->idx];
                if (k == BLAST_CMD_START) g.start(); else if (k == BLAST_CMD_PAUSE) g.pause();
                else if (k == BLAST_CMD_RESUME) g.resume(); else g.stop();
            } else if (cmd->idx_kind == BLAST_IDX_TEMPO) {
                if (cmd->idx >= c->tempo_cons.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "tempo context %s", kOob);
                Tempo& t = *c->tempo_cons[cmd->idx];
                if (k == BLAST_CMD_START) t.start(); else if (k == BLAST_CMD_PAUSE) t.active = false;
                else if (k == BLAST_CMD_RESUME) t.active = true; else t.stop();
            }                                                    // Idx::Process: `_ => ()`
            return BLAST_OK;
        }
        case BLAST_CMD_UNLOAD:                                   // engine.rs:182-184
            if (cmd->idx >= c->voices.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "unload: voice %s", kOob);
            c->voices.erase(c->voices.begin() + (ptrdiff_t)cmd->idx);
            return BLAST_OK;
        case BLAST_CMD_VELOCITY:                                 // engine.rs:186-189
            if (cmd->idx >= c->voices.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "velocity: voice %s", kOob);
            c->voices[cmd->idx].vel = cmd->val;
            return BLAST_OK;
        case BLAST_CMD_GROUP: {                                  // engine.rs:191-212
            if (cmd->n_members && !cmd->member_voice) return blast::set_error(BLAST_ERR_ARG, "group: member_voice is null");
            if (cmd->tempo.mode > BLAST_TM_TBD || cmd->tempo.unit > BLAST_TU_BPM) return blast::set_error(BLAST_ERR_ARG, "group: bad tempo mode / unit");
            TempoRef tempo = c->tempo_from_repr(cmd->tempo);
            if (!tempo) return blast::set_error(BLAST_ERR_REF_PANIC, "group: tempo %s", kOob);
            // validate the whole removal sequence first: nothing changes unless every index is good
            std::vector<uint32_t> ids(c->voices.size());
            std::iota(ids.begin(), ids.end(), 0u);
            std::vector<uint32_t> picked;
            size_t pcur = 0;
            for (uint32_t m = 0; m < cmd->n_members; ++m) {
                const uint64_t vi = cmd->member_voice[m];
                if (vi >= ids.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "group: voice %s", kOob);
                const uint32_t orig = ids[vi];
                ids.erase(ids.begin() + (ptrdiff_t)vi);
                picked.push_back(orig);
                const uint32_t np = cmd->member_n_procs ? cmd->member_n_procs[m] : 0;
                if (np && !cmd->member_proc_ids) return blast::set_error(BLAST_ERR_ARG, "group: member_proc_ids is null");
                if (cmd->member_update_tempo && cmd->member_update_tempo[m])
                    for (uint32_t p = 0; p < np; ++p)
                        if (cmd->member_proc_ids[pcur + p] >= c->voices[orig].procs.size())
                            return blast::set_error(BLAST_ERR_REF_PANIC, "group: process %s", kOob);
                pcur += np;
            }
            GroupH g;
            g.tempo = tempo;
            pcur = 0;
            for (uint32_t m = 0; m < cmd->n_members; ++m) {
                VoiceH v = std::move(c->voices[picked[m]]);
                const uint32_t np = cmd->member_n_procs ? cmd->member_n_procs[m] : 0;
                if (cmd->member_update_tempo && cmd->member_update_tempo[m]) {
                    v.tempo = tempo;
                    for (uint32_t p = 0; p < np; ++p) v.procs[cmd->member_proc_ids[pcur + p]].tempo = tempo;
                }
                pcur += np;
                g.voices.push_back(std::move(v));
            }
            std::vector<VoiceH> rest;
            rest.reserve(ids.size());
            for (uint32_t orig : ids) rest.push_back(std::move(c->voices[orig]));
            c->voices = std::move(rest);
            c->groups.push_back(std::move(g));
            return BLAST_OK;
        }
        case BLAST_CMD_TC: {                                     // engine.rs:214-217
            if (cmd->tempo.mode > BLAST_TM_TBD || cmd->tempo.unit > BLAST_TU_BPM) return blast::set_error(BLAST_ERR_ARG, "tc: bad tempo mode / unit");
            TempoRef tempo = c->tempo_from_repr(cmd->tempo);
            if (!tempo) return blast::set_error(BLAST_ERR_REF_PANIC, "tc: tempo %s", kOob);
            c->tempo_cons.push_back(tempo);
            return BLAST_OK;
        }
        case BLAST_CMD_SEQ: {                                    // engine.rs:221-248
            if (cmd->n_steps && (!cmd->steps || !cmd->chance)) return blast::set_error(BLAST_ERR_ARG, "seq: steps / chance is null");
            if (cmd->tempo.mode > BLAST_TM_TBD || cmd->tempo.unit > BLAST_TU_BPM) return blast::set_error(BLAST_ERR_ARG, "seq: bad tempo mode / unit");
            TempoRef tempo = c->tempo_from_repr(cmd->tempo);
            if (!tempo) return blast::set_error(BLAST_ERR_REF_PANIC, "seq: tempo %s", kOob);
            SeqH s;
            s.tempo = tempo;
            s.period = cmd->period;
            s.steps.assign(cmd->steps, cmd->steps + cmd->n_steps);
            s.chance.assign(cmd->chance, cmd->chance + cmd->n_steps);
            s.rng = blast_x128p{cmd->rng_s0, cmd->rng_s1};
            if (cmd->idx_kind == BLAST_IDX_VOICE) {
                if (cmd->idx >= c->voices.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "seq: voice %s", kOob);
                VoiceH& v = c->voices[cmd->idx];
                if (v.procs.size() >= 255) return blast::set_error(BLAST_ERR_CAPACITY, "seq: at most 255 processes per voice");
                v.procs.push_back(std::move(s));
                if (cmd->tempo.mode == BLAST_TM_PROCESS) v.proc_tempi.push_back(tempo);
            } else if (cmd->idx_kind == BLAST_IDX_GROUP) {
                if (cmd->idx >= c->groups.size()) return blast::set_error(BLAST_ERR_REF_PANIC, "seq: group %s", kOob);
                c->groups[cmd->idx].processes.push_back(std::move(s));
            }                                                    // `_ => ()`
            return BLAST_OK;
        }
        case BLAST_CMD_QUIT: return BLAST_OK;                    // the reference raises SIGTERM (engine.rs:95-99)
        default: return blast::set_error(BLAST_ERR_ARG, "unknown command kind %u", cmd->kind);
    }
}

// One span of the flattened scene, in process() call order.
struct Flat {
    std::vector<VoiceH*> called;                  // every voice whose process() runs, in call order
    std::vector<VoiceDev> dev;                    // the ones this rank renders
    std::vector<VoiceH*> dev_owner;
    std::vector<SeqDev> seqs;                     // live Seqs of the rendered voices
    std::vector<SeqH*> seq_owner;
    std::vector<size_t> seq_pool_off;             // offset of steps[] in the float pool (chance follows)
    std::vector<float> pool;
    std::vector<Tempo*> ticked;                   // every tempo somebody updates; its ticks per call are in Tempo::flat_ticks
    uint64_t epoch = 0;                           // (a hash map here cost ~0.4 ms per span at 4,096 voices)
    uint32_t ticks_of(const Tempo* t) const { return t->flat_epoch == epoch ? t->flat_ticks : 0u; }
    void tick(Tempo* t) {
        if (t->flat_epoch != epoch) { t->flat_epoch = epoch; t->flat_ticks = 0; ticked.push_back(t); }
        t->flat_ticks += 1;
    }
};

int flatten(blast_conductor* c, Flat& f) {
    const uint32_t oc = c->out_channels;
    struct Pending { size_t seq; Tempo* tempo; };
    std::vector<Pending> pending;
    f.epoch = ++c->flat_epoch;
    {
        size_t nv = c->voices.size();
        for (auto& g : c->groups) nv += g.voices.size();
        f.called.reserve(nv); f.dev.reserve(nv); f.dev_owner.reserve(nv); f.ticked.reserve(nv);
        f.seqs.reserve(nv); f.seq_owner.reserve(nv); f.seq_pool_off.reserve(nv); pending.reserve(nv);
    }
    auto visit = [&](VoiceH& v) -> int {
        f.called.push_back(&v);
        const bool mine = ((c->shard_by_track ? (uint64_t)v.track : v.uid) % c->world) == c->rank;
        // processes run first and see the ticks of the voices before this one in the same call (engine.rs:392-394)
        uint32_t n_live = 0;
        const size_t first = f.seqs.size();
        for (auto& s : v.procs) {
            if (!s.active || !s.tempo->active) continue;                             // processes.rs:70-75
            if (s.idx >= s.steps.size())
                return blast::set_error(BLAST_ERR_REF_PANIC, "Seq with an empty step list is processed (index out of bounds, processes.rs:79)");
            if (!mine) continue;
            SeqDev q{};
            q.base = s.tempo->current + f.ticks_of(s.tempo.get());
            q.rate = 0;                                                              // filled once every ticker is known
            q.interval = s.tempo->interval;
            q.period_f = (float)s.period;
            q.n_steps = (uint32_t)s.steps.size();
            q.idx = (uint32_t)s.idx;
            q.s0 = s.rng.s0;
            q.s1 = s.rng.s1;
            f.seq_pool_off.push_back(f.pool.size());
            f.pool.insert(f.pool.end(), s.steps.begin(), s.steps.end());
            f.pool.insert(f.pool.end(), s.chance.begin(), s.chance.end());
            pending.push_back({f.seqs.size(), s.tempo.get()});
            f.seqs.push_back(q);
            f.seq_owner.push_back(&s);
            n_live += 1;
        }
        if (v.tempo->mode == BLAST_TM_VOICE || v.tempo->mode == BLAST_TM_TBD) f.tick(v.tempo.get());   // engine.rs:396-400
        for (auto& t : v.proc_tempi) f.tick(t.get());                                                  // engine.rs:402-405
        if (mine) {
            if (first + n_live > 0xFFFFFFull) return blast::set_error(BLAST_ERR_CAPACITY, "too many Seq processes in one span");
            VoiceDev d{};
            d.smp = v.d_smp;
            d.end = (uint32_t)v.end;
            d.C = v.C;
            d.pos = v.pos;
            d.vel = v.vel;
            d.gain = v.gain;
            d.active = 1;
            route_voice(d, oc, n_live > 0);
            d.first_seq = (uint32_t)first | (n_live << 24);
            f.dev.push_back(d);
            f.dev_owner.push_back(&v);
        }
        return BLAST_OK;
    };
    for (auto& v : c->voices)
        if (v.active)
            if (int rc = visit(v)) return rc;
    for (auto& g : c->groups) {
        if (!g.active) continue;
        for (auto& v : g.voices)
            if (v.active)
                if (int rc = visit(v)) return rc;
        if (g.tempo->mode == BLAST_TM_GROUP) f.tick(g.tempo.get());            // engine.rs:537-540
    }
    for (auto& p : pending) {
        f.seqs[p.seq].rate = f.ticks_of(p.tempo);
    }
    return BLAST_OK;
}

int ensure_pin(blast_conductor* c, size_t bytes) {
    if (bytes <= c->pin_cap) return BLAST_OK;
    if (c->h_pin) cudaFreeHost(c->h_pin);
    c->h_pin = nullptr;
    c->pin_cap = 0;
    const size_t cap = std::max<size_t>(bytes * 2, 1 << 16);
    BLAST_CUDA_TRY(cudaMallocHost(&c->h_pin, cap));
    c->pin_cap = cap;
    return BLAST_OK;
}

// renders `frames` frames from the current host state; commits the state only when the device reported no
// capacity overflow.  *overflow tells the caller to retry with a shorter chunk.
int render_chunk(blast_ctx* ctx, blast_conductor* c, Flat& f, uint64_t frames, int32_t* d_partial, bool* overflow) {
    *overflow = false;
    const uint32_t oc = c->out_channels;
    const uint32_t nv = (uint32_t)f.dev.size(), ns = (uint32_t)f.seqs.size();
    if (int rc = reserve_buffers(ctx, c->rb, nv, ns)) return rc;
    if (nv == 0) return launch_render(ctx, c->rb, 0, 0, oc, frames, d_partial);
    // pinned staging layout: [voices][seqs][pool] + readback [voices][seqs][err]
    const size_t vb = (size_t)nv * sizeof(VoiceDev), sb = (size_t)ns * sizeof(SeqDev), pb = f.pool.size() * sizeof(float);
    const size_t up = vb + sb + pb, total = up + vb + sb + 16;
    if (int rc = ensure_pin(c, total)) return rc;
    uint8_t* h = static_cast<uint8_t*>(c->h_pin);
    if (pb > c->fpool_cap * sizeof(float)) {
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (c->d_fpool) cudaFree(c->d_fpool);
        c->d_fpool = nullptr;
        c->fpool_cap = 0;
        BLAST_CUDA_TRY(cudaMalloc(&c->d_fpool, pb * 2));
        c->fpool_cap = f.pool.size() * 2;
    }
    for (uint32_t i = 0; i < ns; ++i) {
        f.seqs[i].steps = c->d_fpool + f.seq_pool_off[i];
        f.seqs[i].chance = f.seqs[i].steps + f.seqs[i].n_steps;
    }
    memcpy(h, f.dev.data(), vb);
    if (sb) memcpy(h + vb, f.seqs.data(), sb);
    if (pb) memcpy(h + vb + sb, f.pool.data(), pb);
    BLAST_CUDA_TRY(cudaMemcpyAsync(c->rb.d_voices, h, vb, cudaMemcpyHostToDevice, ctx->stream));
    if (sb) BLAST_CUDA_TRY(cudaMemcpyAsync(c->rb.d_seqs, h + vb, sb, cudaMemcpyHostToDevice, ctx->stream));
    if (pb) BLAST_CUDA_TRY(cudaMemcpyAsync(c->d_fpool, h + vb + sb, pb, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_render(ctx, c->rb, nv, ns, oc, frames, d_partial)) return rc;
    uint8_t* back = h + up;
    BLAST_CUDA_TRY(cudaMemcpyAsync(back, c->rb.d_voices, vb, cudaMemcpyDeviceToHost, ctx->stream));
    if (sb) BLAST_CUDA_TRY(cudaMemcpyAsync(back + vb, c->rb.d_seqs, sb, cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaMemcpyAsync(back + vb + sb, c->rb.err_word(), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint32_t err;
    memcpy(&err, back + vb + sb, sizeof(err));
    if (err) { *overflow = true; return BLAST_OK; }
    const VoiceDev* dv = reinterpret_cast<const VoiceDev*>(back);
    for (uint32_t i = 0; i < nv; ++i) f.dev_owner[i]->pos = dv[i].pos;
    const SeqDev* ds = reinterpret_cast<const SeqDev*>(back + vb);
    for (uint32_t i = 0; i < ns; ++i) {
        f.seq_owner[i]->idx = ds[i].idx;
        f.seq_owner[i]->rng = blast_x128p{ds[i].s0, ds[i].s1};
    }
    return BLAST_OK;
}

int render_span(blast_ctx* ctx, blast_conductor* c, uint64_t frames, int32_t* d_partial) {
    const uint32_t oc = c->out_channels;
    uint64_t done = 0;
    uint64_t chunk = frames;
    static const bool debug = getenv("BLAST_CONDUCTOR_DEBUG") != nullptr;
    uint32_t n_ok = 0, n_retry = 0;
    while (done < frames) {
        Flat f;
        if (int rc = flatten(c, f)) return rc;
        // bounds: per-tile records (16 B per tile per voice) <= 256 MiB; calls fit in 31 bits when Seqs are live
        const uint64_t nv = std::max<uint64_t>(f.dev.size(), 1);
        uint64_t cap = std::max<uint64_t>(1, (1ull << 24) / nv) * (uint64_t)kFT;
        cap = std::min<uint64_t>(cap, 1ull << 30);
        if (!f.seqs.empty()) cap = std::min<uint64_t>(cap, (1ull << 22));
        uint64_t n = std::min<uint64_t>({chunk, cap, frames - done});
        bool overflow = false;
        if (int rc = render_chunk(ctx, c, f, n, d_partial + done * oc, &overflow)) return rc;
        if (overflow) {
            n_retry += 1;
            if (n == 1) return blast::set_error(BLAST_ERR_CAPACITY, "a single frame needs more than %d position segments / %d retriggers", kMaxSeg, kMaxEvents);
            chunk = std::max<uint64_t>(1, n / 2);
            continue;
        }
        // commit the tempo counters: every ticker adds 1 per call (blast_time.rs:113-115, u32 wrapping)
        const uint64_t calls = n * oc;
        for (Tempo* t : f.ticked) t->current += (uint32_t)((uint64_t)t->flat_ticks * calls);
        c->clock += n;                                                               // clock::advance(1) per frame
        done += n;
        n_ok += 1;
    }
    if (debug) fprintf(stderr, "[blast conductor] span of %llu frames: %u chunk(s), %u overflow retr%s, last chunk %llu frames\n",
                       (unsigned long long)frames, n_ok, n_retry, n_retry == 1 ? "y" : "ies", (unsigned long long)chunk);
    return BLAST_OK;
}

VoiceH* find_voice(blast_conductor* c, int group, uint32_t idx) {
    std::vector<VoiceH>* vs = nullptr;
    if (group < 0) vs = &c->voices;
    else if ((size_t)group < c->groups.size()) vs = &c->groups[group].voices;
    if (!vs || idx >= vs->size()) return nullptr;
    return &(*vs)[idx];
}

}  // namespace

extern "C" {

float blast_convert_interval(uint32_t sample_rate, uint32_t unit, float interval) {
    return convert_interval(sample_rate, unit, interval);
}

int blast_conductor_create(blast_ctx* ctx, uint32_t out_channels, uint32_t sample_rate, const blast_track* tracks,
                           uint32_t n_tracks, blast_conductor** out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(out != nullptr, BLAST_ERR_ARG, "blast_conductor_create: out is null");
    BLAST_REQUIRE(tracks || n_tracks == 0, BLAST_ERR_ARG, "blast_conductor_create: tracks is null");
    if (out_channels < 1 || out_channels > (uint32_t)kMaxOut)
        return blast::set_error(BLAST_ERR_UNSUPPORTED, "out_channels must be 1..%d (got %u)", kMaxOut, out_channels);
    auto* c = new blast_conductor();
    c->out_channels = out_channels;
    c->sample_rate = sample_rate;
    c->tracks.assign(tracks, tracks + n_tracks);
    *out = c;
    return BLAST_OK;
}

void blast_conductor_destroy(blast_ctx* ctx, blast_conductor* c) {
    if (!c) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    free_buffers(c->rb);
    if (c->d_fpool) cudaFree(c->d_fpool);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    delete c;
}

int blast_conductor_apply(blast_ctx* ctx, blast_conductor* c, const blast_command* cmd) {
    (void)ctx;
    BLAST_REQUIRE(c && cmd, BLAST_ERR_ARG, "blast_conductor_apply: null argument");
    return apply_command(c, cmd);
}

int blast_conductor_set_shard(blast_conductor* c, uint32_t rank, uint32_t world) {
    BLAST_REQUIRE(c != nullptr, BLAST_ERR_ARG, "blast_conductor_set_shard: null conductor");
    BLAST_REQUIRE(world >= 1 && rank < world, BLAST_ERR_ARG, "blast_conductor_set_shard: need rank < world");
    c->rank = rank;
    c->world = world;
    c->shard_by_track = false;
    return BLAST_OK;
}

int blast_conductor_set_shard_by_track(blast_conductor* c, uint32_t rank, uint32_t world) {
    BLAST_REQUIRE(c != nullptr, BLAST_ERR_ARG, "blast_conductor_set_shard_by_track: null conductor");
    BLAST_REQUIRE(world >= 1 && rank < world, BLAST_ERR_ARG, "blast_conductor_set_shard_by_track: need rank < world");
    c->rank = rank;
    c->world = world;
    c->shard_by_track = true;
    return BLAST_OK;
}

int blast_conductor_reserve(blast_ctx* ctx, blast_conductor* c, uint64_t frames) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(c != nullptr, BLAST_ERR_ARG, "blast_conductor_reserve: null conductor");
    Flat f;
    if (int rc = flatten(c, f)) return rc;
    const uint32_t nv = (uint32_t)f.dev.size(), ns = (uint32_t)f.seqs.size();
    if (int rc = reserve_buffers(ctx, c->rb, nv, ns)) return rc;
    if (int rc = reserve_frames(ctx, c->rb, nv, c->out_channels, frames)) return rc;
    const size_t vb = (size_t)nv * sizeof(VoiceDev), sb = (size_t)ns * sizeof(SeqDev), pb = f.pool.size() * sizeof(float);
    if (int rc = ensure_pin(c, 2 * (vb + sb) + pb + 16)) return rc;
    if (pb > c->fpool_cap * sizeof(float)) {
        BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (c->d_fpool) cudaFree(c->d_fpool);
        c->d_fpool = nullptr;
        c->fpool_cap = 0;
        BLAST_CUDA_TRY(cudaMalloc(&c->d_fpool, pb * 2));
        c->fpool_cap = f.pool.size() * 2;
    }
    return BLAST_OK;
}

int blast_conductor_render_dev(blast_ctx* ctx, blast_conductor* c, uint64_t frames, int32_t* d_partial_bus) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(c && (d_partial_bus || frames == 0), BLAST_ERR_ARG, "blast_conductor_render_dev: null argument");
    if (frames == 0) return BLAST_OK;
    return render_span(ctx, c, frames, d_partial_bus);
}

int blast_conductor_render_timeline_dev(blast_ctx* ctx, blast_conductor* c, const blast_timed_command* events,
                                        uint32_t n_events, uint64_t total_frames, int32_t* d_partial_bus) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(c && (events || n_events == 0) && (d_partial_bus || total_frames == 0), BLAST_ERR_ARG,
                  "blast_conductor_render_timeline_dev: null argument");
    uint64_t prev = 0;
    for (uint32_t i = 0; i < n_events; ++i) {
        if (events[i].frame < prev || events[i].frame > total_frames)
            return blast::set_error(BLAST_ERR_ARG, "timeline event %u: frames must be sorted and <= total_frames", i);
        prev = events[i].frame;
    }
    uint64_t cur = 0;
    for (uint32_t i = 0; i < n_events; ++i) {
        if (events[i].frame > cur) {
            if (int rc = render_span(ctx, c, events[i].frame - cur, d_partial_bus + cur * c->out_channels)) return rc;
            cur = events[i].frame;
        }
        if (int rc = apply_command(c, &events[i].cmd)) return rc;
    }
    if (total_frames > cur)
        if (int rc = render_span(ctx, c, total_frames - cur, d_partial_bus + cur * c->out_channels)) return rc;
    return BLAST_OK;
}

static int with_host_bus(blast_ctx* ctx, blast_conductor* c, uint64_t frames, int16_t* host_bus_out,
                         const blast_timed_command* events, uint32_t n_events, bool timeline) {
    const size_t slots = (size_t)frames * c->out_channels;
    if (slots == 0 && !timeline) return BLAST_OK;
    int32_t* d_partial = static_cast<int32_t*>(blast::scratch(ctx, 10, std::max<size_t>(slots, 1) * sizeof(int32_t)));
    int16_t* d_bus = static_cast<int16_t*>(blast::scratch(ctx, 11, std::max<size_t>(slots, 1) * sizeof(int16_t) + 16));
    if (!d_partial || !d_bus) return BLAST_ERR_CUDA;
    int rc = timeline ? blast_conductor_render_timeline_dev(ctx, c, events, n_events, frames, d_partial)
                      : render_span(ctx, c, frames, d_partial);
    if (rc != BLAST_OK) return rc;
    if (slots == 0) return BLAST_OK;
    if ((rc = blast_bus_finalize_dev(ctx, d_partial, d_bus, slots)) != BLAST_OK) return rc;
    BLAST_CUDA_TRY(cudaMemcpyAsync(host_bus_out, d_bus, slots * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    BLAST_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BLAST_OK;
}

int blast_conductor_coordinate(blast_ctx* ctx, blast_conductor* c, uint64_t frames, int16_t* host_bus_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(c && (host_bus_out || frames == 0), BLAST_ERR_ARG, "blast_conductor_coordinate: null argument");
    return with_host_bus(ctx, c, frames, host_bus_out, nullptr, 0, false);
}

int blast_conductor_render_timeline(blast_ctx* ctx, blast_conductor* c, const blast_timed_command* events,
                                    uint32_t n_events, uint64_t total_frames, int16_t* host_bus_out) {
    if (int rc = blast::bind(ctx)) return rc;
    BLAST_REQUIRE(c && (events || n_events == 0) && (host_bus_out || total_frames == 0), BLAST_ERR_ARG,
                  "blast_conductor_render_timeline: null argument");
    return with_host_bus(ctx, c, total_frames, host_bus_out, events, n_events, true);
}

int blast_conductor_n_voices(const blast_conductor* c, int group) {
    if (!c) return -1;
    if (group < 0) return (int)c->voices.size();
    if ((size_t)group >= c->groups.size()) return -1;
    return (int)c->groups[group].voices.size();
}

int blast_conductor_n_groups(const blast_conductor* c) { return c ? (int)c->groups.size() : -1; }

int blast_conductor_get_voice(const blast_conductor* c, int group, uint32_t idx, blast_voice_state* out) {
    BLAST_REQUIRE(c && out, BLAST_ERR_ARG, "blast_conductor_get_voice: null argument");
    const VoiceH* v = find_voice(const_cast<blast_conductor*>(c), group, idx);
    if (!v) return blast::set_error(BLAST_ERR_ARG, "no such voice");
    out->active = v->active;
    out->position = v->pos;
    out->velocity = v->vel;
    out->gain = v->gain;
    out->end = v->end;
    out->channels = v->C;
    out->tempo_current = v->tempo->current;
    out->tempo_active = v->tempo->active;
    out->n_processes = (uint32_t)v->procs.size();
    return BLAST_OK;
}

int blast_conductor_set_voice(blast_conductor* c, int group, uint32_t idx, const float* position, const float* velocity,
                              const float* gain, const int* active) {
    BLAST_REQUIRE(c != nullptr, BLAST_ERR_ARG, "blast_conductor_set_voice: null conductor");
    VoiceH* v = find_voice(c, group, idx);
    if (!v) return blast::set_error(BLAST_ERR_ARG, "no such voice");
    if (position) v->pos = *position;
    if (velocity) v->vel = *velocity;
    if (gain) v->gain = *gain;
    if (active) v->active = *active != 0;
    return BLAST_OK;
}

uint64_t blast_conductor_clock(const blast_conductor* c) { return c ? c->clock : 0; }

}  // extern "C"
