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
