// group_harness.cpp — the single-process multi-GPU surface (blast_group) driven from plain C++: no Python, no torch, no
// launcher.  What a host like the reference's main() (blast/src/main.rs:13-128) would do: decode an asset set over the
// group, take the consensus, run a Conductor over all GPUs, read one bus.  Checked here against arithmetic the harness
// can do itself (velocity 1.0 / gain 1.0 voices: the mix is the wrapping i16 sum of the decoded samples, engine.rs:441).
//   g++ -std=c++17 -I include tests/checks/group_harness.cpp -L audio_decoder_b200 -lblast_cuda -o _group_harness
//   ./_group_harness 0 1 2 3        (device ids; an id may repeat)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "blast_cuda.h"

#define CHECK(call)                                                                               \
    do {                                                                                          \
        int _rc = (call);                                                                         \
        if (_rc != BLAST_OK) {                                                                    \
            std::fprintf(stderr, "%s failed: status %d: %s\n", #call, _rc, blast_last_error());   \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

static uint64_t lcg(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return s >> 11; }

static std::vector<uint8_t> wav_image(uint64_t seed, uint32_t data_len) {
    std::vector<uint8_t> f(44 + data_len);
    auto u32 = [&](size_t at, uint32_t v) { std::memcpy(&f[at], &v, 4); };
    auto u16 = [&](size_t at, uint16_t v) { std::memcpy(&f[at], &v, 2); };
    std::memcpy(&f[0], "RIFF", 4); u32(4, 36 + data_len); std::memcpy(&f[8], "WAVEfmt ", 8); u32(16, 16);
    u16(20, 1); u16(22, 2); u32(24, 48000); u32(28, 48000 * 4); u16(32, 4); u16(34, 16);
    std::memcpy(&f[36], "data", 4); u32(40, data_len);
    for (uint32_t i = 0; i < data_len; ++i) f[44 + i] = (uint8_t)lcg(seed);
    return f;
}

int main(int argc, char** argv) {
    std::vector<int> devs;
    for (int i = 1; i < argc; ++i) devs.push_back(std::atoi(argv[i]));
    if (devs.empty()) devs.push_back(0);
    blast_group* g = nullptr;
    CHECK(blast_group_create(&g, devs.data(), (uint32_t)devs.size()));
    const uint32_t n_files = 11, frames = 30011;

    // ---- main.rs:18-89: decode every asset (file i -> member i mod n), consensus of rate / channels
    std::vector<std::vector<uint8_t>> images;
    std::vector<const uint8_t*> files;
    std::vector<size_t> lens;
    std::vector<blast_pcm_desc> descs(n_files);
    // clips longer than the render: a voice is silent from the step whose trunc(position) reaches end = frames - 1
    // (engine.rs:407-410), so the last frame of a clip exactly as long as the render would not play
    for (uint32_t i = 0; i < n_files; ++i) images.push_back(wav_image(1000 + i, (frames + 7 * (i + 1)) * 4));
    for (uint32_t i = 0; i < n_files; ++i) {
        files.push_back(images[i].data());
        lens.push_back(images[i].size());
        CHECK(blast_wav_probe(files[i], lens[i], &descs[i]));
    }
    uint32_t rate = 0, chans = 0;
    CHECK(blast_asset_consensus(descs.data(), n_files, &rate, &chans));
    if (rate != 48000 || chans != 2) { std::fprintf(stderr, "consensus %u / %u\n", rate, chans); return 1; }
    std::vector<std::vector<int16_t>> samples(n_files);
    std::vector<int16_t*> host_out;
    for (uint32_t i = 0; i < n_files; ++i) { samples[i].resize(blast_pcm_out_len(&descs[i])); host_out.push_back(samples[i].data()); }
    std::vector<blast_track> tracks(n_files);
    CHECK(blast_group_pcm_decode_batch(g, n_files, files.data(), lens.data(), descs.data(), host_out.data(), tracks.data()));
    for (uint32_t i = 0; i < n_files; ++i)                           // WAV is little-endian: the payload as it lies
        if (std::memcmp(samples[i].data(), images[i].data() + 44, samples[i].size() * 2) != 0) { std::fprintf(stderr, "decode %u differs\n", i); return 1; }

    // ---- static scene over the group: every track twice, velocity 1, gain 1 -> wrapping i16 sum
    std::vector<blast_voice> voices;
    for (uint32_t k = 0; k < 2 * n_files; ++k) voices.push_back(blast_voice{k % n_files, 1u, 0.0f, 1.0f, 1.0f, 0u});
    std::vector<int16_t> bus((size_t)frames * 2), want((size_t)frames * 2);
    for (size_t s = 0; s < want.size(); ++s) {
        int32_t acc = 0;
        for (const blast_voice& v : voices) acc += samples[v.track][s];
        want[s] = (int16_t)acc;
    }
    for (int rep = 0; rep < 3; ++rep) {
        std::fill(bus.begin(), bus.end(), 0);
        CHECK(blast_group_render(g, tracks.data(), n_files, voices.data(), (uint32_t)voices.size(), 2, frames, bus.data()));
        if (bus != want) { std::fprintf(stderr, "group render differs (rep %d)\n", rep); return 1; }
    }

    // ---- the Conductor over the group: load + start every track, two spans with a Stop between them
    blast_group_conductor* gc = nullptr;
    CHECK(blast_group_conductor_create(g, 2, rate, tracks.data(), n_files, &gc));
    for (uint32_t t = 0; t < n_files; ++t) {
        blast_command c{};
        c.kind = BLAST_CMD_LOAD; c.idx = t;
        c.tempo = blast_tempo_repr{0, 1, BLAST_TM_TBD, BLAST_TU_SAMPLES, 0.0f};
        CHECK(blast_group_conductor_apply(gc, &c));
        blast_command s{};
        s.kind = BLAST_CMD_START; s.idx_kind = BLAST_IDX_VOICE; s.idx = t;
        CHECK(blast_group_conductor_apply(gc, &s));
    }
    const uint32_t span1 = 12345, span2 = 9000;
    std::vector<int16_t> b1((size_t)span1 * 2), b2((size_t)span2 * 2);
    CHECK(blast_group_conductor_coordinate(gc, span1, b1.data()));
    blast_command stop{};
    stop.kind = BLAST_CMD_STOP; stop.idx_kind = BLAST_IDX_VOICE; stop.idx = 3;
    CHECK(blast_group_conductor_apply(gc, &stop));
    CHECK(blast_group_conductor_coordinate(gc, span2, b2.data()));
    for (size_t s = 0; s < b1.size(); ++s) {
        int32_t acc = 0;
        for (uint32_t t = 0; t < n_files; ++t) acc += samples[t][s];
        if (b1[s] != (int16_t)acc) { std::fprintf(stderr, "conductor span 1 differs at %zu\n", s); return 1; }
    }
    for (size_t s = 0; s < b2.size(); ++s) {
        int32_t acc = 0;
        for (uint32_t t = 0; t < n_files; ++t) if (t != 3) acc += samples[t][(size_t)span1 * 2 + s];
        if (b2[s] != (int16_t)acc) { std::fprintf(stderr, "conductor span 2 differs at %zu\n", s); return 1; }
    }
    // an out-of-range index is the reference's panic, reported by every member alike
    blast_command bad{};
    bad.kind = BLAST_CMD_VELOCITY; bad.idx = 999; bad.val = 2.0f;
    if (blast_group_conductor_apply(gc, &bad) != BLAST_ERR_REF_PANIC) { std::fprintf(stderr, "bad index not reported\n"); return 1; }
    blast_group_conductor_destroy(gc);

    // ---- RNG streams over the group against X128P::new + sequential draws done here (blast_rand.rs:10-39)
    {
        const uint64_t n_streams = 10, draws = 33, stride = 50;
        std::vector<uint64_t> raw(n_streams * draws);
        CHECK(blast_group_x128p_fill(g, 42, stride, n_streams, draws, 0, 100, raw.data(), nullptr, nullptr));
        blast_x128p st;
        blast_x128p_seed(42, &st);
        auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
        std::vector<uint64_t> seq(stride * n_streams + draws);
        for (auto& r : seq) {
            r = st.s0 + st.s1;
            const uint64_t t = st.s1 ^ st.s0;
            st.s0 = rotl(st.s0, 55) ^ t ^ (t << 14);
            st.s1 = rotl(t, 36);
        }
        for (uint64_t s = 0; s < n_streams; ++s)
            for (uint64_t j = 0; j < draws; ++j)
                if (raw[s * draws + j] != seq[s * stride + j]) { std::fprintf(stderr, "rng stream %llu draw %llu differs\n", (unsigned long long)s, (unsigned long long)j); return 1; }
    }
    blast_group_destroy(g);
    std::printf("group harness OK: %zu member(s), decode + render + conductor + rng bit-exact\n", devs.size());
    return 0;
}
