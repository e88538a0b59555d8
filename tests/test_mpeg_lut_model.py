"""CPU model of K8's candidate validity rule (audio_decoder_b200/csrc/mpeg_scan.cu: cand_valid_lut) against the oracle's
parse_header / match_ref / compute_frame_len (mpeg.rs:367-496, 194-204, 207-234): a candidate that matches the reference
header can only differ from it in the bitrate index and the padding bit, so validity is a mask compare plus a 28-entry
table built from the reference header."""
import numpy as np

import oracle

REF_MASK = 0x00170CC0          # kRefMask: version / protection, layer, sample-rate and channel-mode bits


def _lut(ref: int):
    lut = {}
    for e in range(16):
        for pad in (0, 1):
            h = (ref & ~0x0000F200) | (e << 12) | (pad << 9)
            o = oracle.mpeg_parse_header(h)
            lut[(e, pad)] = (o.payload_len, o.skip) if (o.ok and o.frame_len_ok) else None
    return lut


def _model(h: int, ref: int, lut):
    if (h ^ ref) & REF_MASK:
        return None
    return lut[((h >> 12) & 0xF, (h >> 9) & 1)]


def _oracle(h: int, ref_hdr):
    o = oracle.mpeg_parse_header(h)
    if not o.ok or not oracle.mpeg_match_ref(ref_hdr, o) or not o.frame_len_ok:
        return None
    return (o.payload_len, o.skip)


def test_mask_and_table_equal_parse_match_frame_len():
    rng = np.random.default_rng(8)
    refs = [0xFFFB9064, 0xFFFA9064, 0xFFF3A044, 0xFFE3508C, 0xFFFD8804, 0xFFF51000 | 0x00, 0xFFFFE0C0 & 0xFFFEEFFF]
    refs = [r for r in refs if oracle.mpeg_parse_header(r).ok]
    assert len(refs) >= 4
    for ref in refs:
        ref_hdr = oracle.mpeg_parse_header(ref)
        lut = _lut(ref)
        cands = [ref, ref ^ 0x200, ref ^ 0x1000, ref ^ 0x100, ref ^ 0x3F, ref ^ 0x10000, ref ^ 0x100000, ref ^ 0x400, ref ^ 0x40]
        # every bitrate index x padding of the reference, each also with one masked bit flipped
        for e in range(16):
            for pad in (0, 1):
                h = (ref & ~0x0000F200) | (e << 12) | (pad << 9)
                cands += [h, h ^ 0x20000, h ^ 0x800, h ^ 0x80]
        cands += [0xFFE00000 | int(x) for x in rng.integers(0, 1 << 21, size=3000)]
        # random values in the free bits only (private / copyright / emphasis ... do not matter)
        cands += [ref ^ (int(x) & ~REF_MASK & 0x001FFFFF) for x in rng.integers(0, 1 << 21, size=3000)]
        for h in cands:
            assert _model(h, ref, lut) == _oracle(h, ref_hdr), (hex(ref), hex(h))


# ---------------------------------------------------------------------------------------------------------------------
# CPU model of the fused pass of the single-GPU index (mpeg_scan.cu: mpeg_first_count, mpeg_dup_blocks, mpeg_scan_blocks,
# the emit pass of mpeg_classify) against the oracle's mpeg::parse with the duplicate-first quirk (mpeg.rs:39, 77-116):
# the outputs of a block of candidates are its valid candidates plus one per header value whose FIRST position lies in the
# block, found by a bisection over the blocks' first positions; the vote histogram may leave out headers that do not parse.
import synth

HDR_BINS = 1 << 21


def _valid_mask(hdr: np.ndarray, ref: int) -> np.ndarray:
    lut = _lut(ref)
    return np.array([_model(int(h), ref, lut) is not None for h in hdr], dtype=bool)


def _fused_index_model(pos: np.ndarray, hdr: np.ndarray, block: int):
    # vote over the headers that parse only (mpeg_hist); ties -> smallest header value (mpeg_pick_ref)
    keys, counts = np.unique(hdr & np.uint32(HDR_BINS - 1), return_counts=True)
    best = None
    for k, c in zip(keys, counts):
        if oracle.mpeg_parse_header(0xFFE00000 | int(k)).ok and (best is None or c > best[1]):
            best = (int(k), int(c))
    ref = 0xFFE00000 | best[0]
    valid = _valid_mask(hdr, ref)
    n = len(pos)
    n_blocks = (n + block - 1) // block
    # mpeg_first_count: per-block valid counts + the first position of every valid header value
    counts_b = np.array([int(valid[b * block:(b + 1) * block].sum()) for b in range(n_blocks)], dtype=np.int64)
    first = {}
    for p, h, v in zip(pos, hdr, valid):
        if v:
            k = int(h) & (HDR_BINS - 1)
            first[k] = min(first.get(k, 1 << 63), int(p))
    # mpeg_dup_blocks: the largest block whose first candidate is not behind the position
    starts = pos[::block]
    for p in first.values():
        lo, hi = 0, n_blocks
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if int(starts[mid]) <= p:
                lo = mid
            else:
                hi = mid
        counts_b[lo] += 1
    base = np.concatenate([[0], np.cumsum(counts_b)])
    # emit pass: every block writes its outputs from its own base
    out = np.zeros(int(base[-1]), dtype=np.uint64)
    for b in range(n_blocks):
        o = int(base[b])
        for p, h, v in zip(pos[b * block:(b + 1) * block], hdr[b * block:(b + 1) * block], valid[b * block:(b + 1) * block]):
            if v:
                reps = 1 + (first[int(h) & (HDR_BINS - 1)] == int(p))
                out[o:o + reps] = p
                o += reps
        assert o == int(base[b + 1]), "a block's count differs from what it emits"
    return ref, out


def test_fused_first_position_and_counts_equal_the_oracle_index():
    for seed, frames, block in [(5, 400, 64), (6, 900, 128), (7, 300, 2048), (8, 1200, 32)]:
        stream = synth.mp3_like(seed, frames)
        pos, hdr = oracle.mpeg_sync_scan(stream)
        exp = oracle.mpeg_parse(stream, reference_compat=True, want_payload=False)
        ref, out = _fused_index_model(pos, hdr, block)
        assert ref == exp["ref_header"]
        assert np.array_equal(out, exp["offsets"])
