// probe.cu — host-side container walks for WAV / AIFF and the file-name rule.
//
// These are the ~20 scalar field reads per file that precede the sample loop in the
// reference; they stay on the host (they are not data-parallel) and produce the
// blast_pcm_desc that drives the decode kernel.  Behaviour follows
//   blast/src/file_parsing/wav.rs:30-138   (print_id / parse_bytes LE / parse header walk)
//   blast/src/file_parsing/aiff.rs:6-154   (print_id / parse_bytes BE / parse_ieee_extended)
// including the quirks listed in SURVEY.md Appendix B (#1, #2, #6, #7).
#include <cmath>
#include <cstring>

#include "blast_internal.h"

namespace {

// A forward-only reader over the file image.  `ok` goes false on the first read past the
// end and stays false (the reference returns UnexpectedEof from the first failing get()).
struct FieldReader {
    const uint8_t* base;
    size_t size;
    size_t at = 0;
    bool ok = true;

    void advance(size_t n) {
        // every skipped byte is bounds-checked in the reference (print_id loops over get())
        if (!ok) return;
        if (n > size || at > size - n) { ok = false; at = size; return; }
        at += n;
    }
    // unchecked cursor move: the extensible-fmt skip moves the cursor without reading
    void jump(size_t n) { if (ok) at += n; }

    template <bool kBigEndian>
    uint32_t uint(int width) {
        if (!ok) return 0;
        if (at >= size || (size_t)width > size - at) { ok = false; at = size; return 0; }
        uint32_t v = 0;
        for (int i = 0; i < width; ++i) {
            uint32_t byte = base[at + i];
            v |= kBigEndian ? byte << (8 * (width - 1 - i)) : byte << (8 * i);
        }
        at += width;
        return v;
    }
};

// Sample-loop bounds rule (wav.rs:143-151 / aiff.rs:159-167): byte pairs (i, i+1) for
// i = off, off+2, ... < off+len must exist.
bool payload_complete(size_t file_size, uint64_t off, uint64_t len) {
    if (len == 0) return true;
    uint64_t pairs = (len + 1) / 2;
    return off + 2 * pairs <= (uint64_t)file_size;
}

// Rust `f64 as u32`: saturating, truncating, NaN -> 0 (aiff.rs:182)
uint32_t saturating_u32(double x) {
    if (std::isnan(x) || x <= 0.0) return 0;
    if (x >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)x;
}

// f64::powi(2.0, e) as compiler-rt evaluates it (square-and-multiply, reciprocal for e < 0)
double pow2i(int e) {
    bool neg = e < 0;
    double base = 2.0, acc = 1.0;
    for (int k = e;;) {
        if (k & 1) acc *= base;
        k /= 2;
        if (k == 0) break;
        base *= base;
    }
    return neg ? 1.0 / acc : acc;
}

// aiff.rs:51-94
double extended_to_f64(const uint8_t b[10]) {
    const bool negative = (b[0] & 0x80) != 0;
    const int exponent = ((b[0] & 0x7F) << 8) | b[1];
    uint64_t mantissa = 0;
    for (int i = 2; i < 10; ++i) mantissa = (mantissa << 8) | b[i];
    if (exponent == 0 && mantissa == 0) return 0.0;
    if (exponent == 0x7FFF) return mantissa ? NAN : (negative ? -INFINITY : INFINITY);
    double v = (double)mantissa * pow2i(exponent - 16383 - 63);
    return negative ? -v : v;
}

}  // namespace

extern "C" {

int blast_wav_probe(const uint8_t* file, size_t len, blast_pcm_desc* out) {
    BLAST_REQUIRE(out != nullptr && (file != nullptr || len == 0), BLAST_ERR_ARG, "blast_wav_probe: null argument");
    FieldReader r{file, len};
    r.advance(4);                                   // "RIFF" (never compared, wav.rs:30-44)
    (void)r.uint<false>(4);                         // riff size
    r.advance(4);                                   // "WAVE"
    r.advance(4);                                   // "fmt "
    const uint32_t fmt_size = r.uint<false>(4);
    const uint32_t tag = r.uint<false>(2);
    if (!r.ok) return blast::set_error(BLAST_ERR_UNEXPECTED_EOF, "UnexpectedEof in WAV header");
    switch (tag) {                                  // FormatCode::from_u16, wav.rs:17-28
        case 0x0001: case 0x0003: case 0x0006: case 0x0007: case 0xFFFE: break;
        default: return blast::set_error(BLAST_ERR_UNSUPPORTED_FORMAT, "Unrecognized format tag");
    }
    const uint32_t channels = r.uint<false>(2);
    const uint32_t rate = r.uint<false>(4);
    (void)r.uint<false>(4);                         // data rate
    (void)r.uint<false>(2);                         // block size
    const uint32_t bits = r.uint<false>(2);
    if (fmt_size >= 18) {                           // wav.rs:112-130
        const uint32_t cb = r.uint<false>(2);
        if (r.ok && cb > 0) {
            (void)r.uint<false>(2);                 // valid bits
            (void)r.uint<false>(4);                 // channel mask
            (void)r.uint<false>(2);                 // old format
            r.jump(91);                             // `for i in 0..14 { end += i }` == 0+1+...+13
        }
    }
    r.advance(4);                                   // "data" (never compared)
    const uint32_t data_size = r.uint<false>(4);
    if (!r.ok) return blast::set_error(BLAST_ERR_UNEXPECTED_EOF, "UnexpectedEof in WAV header");
    out->sample_rate = rate;
    out->num_channels = channels;
    out->bits_per_sample = bits;
    out->big_endian = 0;
    out->data_off = r.at;
    out->data_len = data_size;
    if (!payload_complete(len, out->data_off, out->data_len))
        return blast::set_error(BLAST_ERR_UNEXPECTED_EOF, "UnexpectedEof in WAV data chunk");
    return BLAST_OK;
}

int blast_aiff_probe(const uint8_t* file, size_t len, blast_pcm_desc* out) {
    BLAST_REQUIRE(out != nullptr && (file != nullptr || len == 0), BLAST_ERR_ARG, "blast_aiff_probe: null argument");
    FieldReader r{file, len};
    r.advance(4);                                   // "FORM"
    (void)r.uint<true>(4);                          // form size
    r.advance(4);                                   // "AIFF"
    r.advance(4);                                   // "COMM"
    const uint32_t comm_size = r.uint<true>(4);
    if (!r.ok) return blast::set_error(BLAST_ERR_UNEXPECTED_EOF, "UnexpectedEof in AIFF header");
    if (comm_size != 18) return blast::set_error(BLAST_ERR_INVALID_DATA, "Comm size should be 18");   // aiff.rs:121-126
    const uint32_t channels = r.uint<true>(2);
    (void)r.uint<true>(4);                          // num sample frames (unused by the reference)
    const uint32_t sample_size = r.uint<true>(2);
    uint8_t ext[10] = {0};
    if (r.ok && r.size - r.at >= 10) {
        std::memcpy(ext, r.base + r.at, 10);
        r.at += 10;
    } else {
        r.ok = false;
    }
    if (!r.ok) return blast::set_error(BLAST_ERR_UNEXPECTED_EOF, "UnexpectedEof in AIFF COMM chunk");
    const double rate = extended_to_f64(ext);
    r.advance(4);                                   // "SSND"
    const uint32_t chunk = r.uint<true>(4);
    const uint32_t ssnd_size = chunk - 8u;          // aiff.rs:146, release-build wrapping
    (void)r.uint<true>(4);                          // offset, ignored
    (void)r.uint<true>(4);                          // block size, ignored
    if (!r.ok) return blast::set_error(BLAST_ERR_UNEXPECTED_EOF, "UnexpectedEof in AIFF SSND chunk");
    out->sample_rate = saturating_u32(rate);
    out->num_channels = channels;
    out->bits_per_sample = sample_size;
    out->big_endian = 1;
    out->data_off = r.at;
    out->data_len = ssnd_size;
    if (!payload_complete(len, out->data_off, out->data_len))
        return blast::set_error(BLAST_ERR_UNEXPECTED_EOF, "UnexpectedEof in AIFF sound data");
    return BLAST_OK;
}

size_t blast_pcm_out_len(const blast_pcm_desc* desc) { return desc ? (size_t)((desc->data_len + 1) / 2) : 0; }

// wav.rs:156-164 / aiff.rs:172-180: name = between the last '/' and the last '.'
int blast_file_name(const char* path, char* out, size_t cap) {
    BLAST_REQUIRE(path && out, BLAST_ERR_ARG, "blast_file_name: null argument");
    const char* dot = std::strrchr(path, '.');
    if (!dot || dot == path || dot[1] == '\0') return blast::set_error(BLAST_ERR_INVALID_DATA, "File has no name");
    const char* slash = nullptr;
    for (const char* p = path; p < dot; ++p)
        if (*p == '/') slash = p;
    if (!slash) return blast::set_error(BLAST_ERR_INVALID_DATA, "File is not nested");
    size_t n = (size_t)(dot - slash - 1);
    if (n + 1 > cap) return blast::set_error(BLAST_ERR_CAPACITY, "file name buffer too small");
    std::memcpy(out, slash + 1, n);
    out[n] = '\0';
    return BLAST_OK;
}

// main.rs:79-120: the engine runs at the most frequent sample rate of the decoded assets and with the largest
// channel count.  Ties between rates follow HashMap iteration order in the reference (nondeterministic); here the
// smallest rate wins.  No assets: 44100 Hz / 2 channels, as the reference's fall-backs.
int blast_asset_consensus(const blast_pcm_desc* descs, uint32_t n, uint32_t* sample_rate_out, uint32_t* num_channels_out) {
    BLAST_REQUIRE((descs || n == 0) && sample_rate_out && num_channels_out, BLAST_ERR_ARG, "blast_asset_consensus: null argument");
    uint32_t best_rate = 44100, best_count = 0, channels = n ? 0 : 2;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t count = 0;
        for (uint32_t k = 0; k < n; ++k) count += descs[k].sample_rate == descs[i].sample_rate;
        if (count > best_count || (count == best_count && descs[i].sample_rate < best_rate)) {
            best_count = count;
            best_rate = descs[i].sample_rate;
        }
        channels = descs[i].num_channels > channels ? descs[i].num_channels : channels;
    }
    *sample_rate_out = best_rate;
    *num_channels_out = channels;
    return BLAST_OK;
}

}  // extern "C"
