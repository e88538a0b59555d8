"""GPU parity: PCM decode (K1) and the 24-bit extension (K2) through the C ABI vs the CPU oracle."""
import os
import tempfile

import numpy as np
import pytest

import audio_decoder_b200 as blast
from audio_decoder_b200 import _lib, file_parsing as fp
import oracle
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = blast.Context(0)
    yield c
    c.close()


def _oracle_words(payload: np.ndarray, n_words: int, be: bool) -> np.ndarray:
    return payload[:2 * n_words].view(">i2" if be else "<i2").astype(np.int16)


def test_decode_batch_matches_oracle_parse(ctx):
    """whole-file path: probe + blast_pcm_decode_batch == faithful wav::parse / aiff::parse"""
    rng = np.random.default_rng(0)
    images, kinds = [], []
    for i in range(24):
        n = int(rng.choice([0, 2, 14, 16, 18, 4000, 4001 if False else 4002, 70000, 300001 * 2]))
        if i % 2:
            images.append(synth.wav_image(100 + i, n)); kinds.append("wav")
        else:
            images.append(synth.aiff_image(100 + i, n)); kinds.append("aiff")
    descs = [fp.probe(k, im) for k, im in zip(kinds, images)]
    outs, dev = fp.decode_batch(ctx, images, descs, to_host=True, keep_on_device=True)
    for k, im, o, d, db in zip(kinds, images, outs, descs, dev):
        _, exp = (oracle.wav_parse if k == "wav" else oracle.aiff_parse)(im)
        assert np.array_equal(o, exp), k
        assert np.array_equal(db.download(np.int16, len(exp)), exp) if len(exp) else True


def test_odd_payload_reads_one_byte_past_chunk(ctx):
    img = np.concatenate([synth.wav_image(7, 4001), np.array([0x5A], dtype=np.uint8)])
    d = fp.probe("wav", img)
    outs, _ = fp.decode_batch(ctx, [img], [d])
    _, exp = oracle.wav_parse(img)
    assert len(exp) == 2001 and np.array_equal(outs[0], exp)
    with pytest.raises(blast.UnexpectedEof):
        fp.probe("wav", img[:-1])
    # a descriptor that lies about the payload is refused before any GPU work
    d.data_len = 10_000
    with pytest.raises(blast.UnexpectedEof):
        fp.decode_batch(ctx, [img], [d])


@pytest.mark.parametrize("be", [False, True])
def test_all_source_and_destination_alignments(ctx, be):
    """device-resident jobs: every src byte alignment 0..15 x dst word alignment 0..7 x ragged lengths"""
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, size=1 << 16, dtype=np.uint8)
    d_src = ctx.to_device(raw)
    d_dst = ctx.alloc(1 << 18)
    jobs, expect = [], []
    out_off = 0
    for sa in range(16):
        for da in range(8):
            n = int(rng.integers(0, 1200))
            s0 = 64 * int(rng.integers(0, 100)) + sa
            out_off = (out_off + 15) // 16 * 16 + 2 * da
            jobs.append((d_src.ptr + s0, d_dst.ptr + out_off, n, be))
            expect.append((out_off, _oracle_words(raw[s0:], n, be)))
            out_off += 2 * n
    d_dst.zero()
    fp.decode_jobs_dev(ctx, jobs)
    got = d_dst.download(np.uint8, 1 << 18)
    mask = np.zeros(1 << 18, dtype=bool)
    for off, exp in expect:
        assert np.array_equal(got[off:off + 2 * len(exp)].view(np.int16), exp)
        mask[off:off + 2 * len(exp)] = True
    assert not got[~mask].any(), "decode wrote outside its destination ranges"


def test_large_ragged_batch_property(ctx):
    """C2-shaped (scaled down) batch, ragged +-10 %: checksum against numpy byteswap + spot files vs oracle"""
    rng = np.random.default_rng(2)
    n_files = 64
    lens = (synth.C2_DATA_LEN // 8 * (0.9 + 0.2 * rng.random(n_files))).astype(np.int64) // 2 * 2
    images = [synth.aiff_image(1000 + i, int(n)) for i, n in enumerate(lens)]
    descs = [fp.probe("aiff", im) for im in images]
    outs, _ = fp.decode_batch(ctx, images, descs)
    for i in (0, 17, 63):
        _, exp = oracle.aiff_parse(images[i])
        assert np.array_equal(outs[i], exp)
    for im, o in zip(images, outs):
        assert np.array_equal(o, im[54:].view(">i2").astype(np.int16))
    # involution: decoding the BE-decoded words as BE again gives the LE reading of the payload
    i = 5
    again_img = np.concatenate([images[i][:54], outs[i].view(np.uint8)])
    outs2, _ = fp.decode_batch(ctx, [again_img], [descs[i]])
    assert np.array_equal(outs2[0], images[i][54:].view("<i2"))


def test_full_size_c1_roundtrip(ctx):
    """BASELINE config 1 at full size: one 105.84 MB WAV; LE decode is the identity on the payload"""
    img = synth.wav_image(0xC1, synth.C1_DATA_LEN)
    d = fp.probe("wav", img)
    assert (d.data_off, d.data_len) == (44, synth.C1_DATA_LEN)
    outs, _ = fp.decode_batch(ctx, [img], [d])
    assert outs[0].size == 52_920_000
    assert np.array_equal(outs[0].view(np.uint8), img[44:])


def test_plan_rerun_and_launch_count(ctx):
    raw = synth.payload(3, 1 << 20)
    d_src = ctx.to_device(raw)
    d_dst = ctx.alloc(1 << 20)
    plan = fp.PcmPlan(ctx, [(d_src.ptr + 6, d_dst.ptr, (1 << 19) - 8, True)])
    before = ctx.launch_count
    for _ in range(3):
        plan.run()
    ctx.sync()
    assert ctx.launch_count - before == 3
    assert plan.words == (1 << 19) - 8
    got = d_dst.download(np.int16, (1 << 19) - 8)
    assert np.array_equal(got, _oracle_words(raw[6:], (1 << 19) - 8, True))
    plan.close()


def test_parse_path_drop_in(ctx):
    """file_parsing::{wav,aiff}::parse(path) -> AudioFile, incl. the name rule applied after decoding"""
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "assets"))
        p = os.path.join(td, "assets", "fairies.wav")
        img = synth.wav_image(11, 9000)
        img.tofile(p)
        af = fp.wav.parse(p, ctx)
        _, exp = oracle.wav_parse(img)
        assert (af.file_name, af.format, af.sample_rate, af.num_channels, af.bits_per_sample) == \
            ("fairies", "wav", 44100, 2, 16)
        assert np.array_equal(af.samples, exp)
        p = os.path.join(td, "assets", "winterly.aif")
        img = synth.aiff_image(12, 9000)
        img.tofile(p)
        af = fp.aiff.parse(p, ctx)
        _, exp = oracle.aiff_parse(img)
        assert (af.file_name, af.format, af.sample_rate, af.bits_per_sample) == ("winterly", "aiff", 48000, 24)
        assert np.array_equal(af.samples, exp)
        bad = os.path.join(td, "assets", "noext")
        img.tofile(bad)
        with pytest.raises(blast.InvalidData):
            fp.aiff.parse(bad, ctx)


@pytest.mark.parametrize("be", [False, True])
@pytest.mark.parametrize("kind", [0, 1])
def test_pcm24_extension(ctx, be, kind):
    rng = np.random.default_rng(4)
    raw = rng.integers(0, 256, size=3 * 50_000 + 64, dtype=np.uint8)
    d_src = ctx.to_device(raw)
    jobs, expect = [], []
    bufs = []
    for sa, n in [(0, 50_000), (1, 1023), (5, 1025), (7, 3), (16, 0), (9, 4096)]:
        dst = ctx.alloc(max(16, n * 4))
        bufs.append(dst)
        jobs.append((d_src.ptr + sa, dst.ptr, n, be, kind))
        ref = oracle.pcm24_unpack(raw[sa:sa + 3 * n], be)
        expect.append(ref if kind == 0 else (ref >> 8).astype(np.int16))
    fp.pcm24_unpack_dev(ctx, jobs)
    for b, exp in zip(bufs, expect):
        got = b.download(np.int32 if kind == 0 else np.int16, len(exp))
        assert np.array_equal(got, exp)


def test_pcm24_job_table_is_reused_only_when_identical(ctx):
    """the job table of a batch that is unpacked again is not uploaded again (pcm_decode.cu: tab8_copy); a batch of
    the same size with other buffers, a 16-bit decode in between (it shares the scratch slot) and the first batch
    again must all come out right"""
    rng = np.random.default_rng(41)
    raws = [rng.integers(0, 256, size=3 * 9000 + 16, dtype=np.uint8) for _ in range(2)]
    d_src = [ctx.to_device(r) for r in raws]
    shapes = [(0, 5000), (3, 4000)]
    sets = []
    for which in (0, 1):
        jobs, expect, bufs = [], [], []
        for sa, n in shapes:
            dst = ctx.alloc(n * 4)
            bufs.append(dst)
            jobs.append((d_src[which].ptr + sa, dst.ptr, n, True, 0))
            expect.append(oracle.pcm24_unpack(raws[which][sa:sa + 3 * n], True))
        sets.append((jobs, expect, bufs))

    def run_and_check(k):
        jobs, expect, bufs = sets[k]
        for b, e in zip(bufs, expect):
            b.upload(np.zeros(len(e), np.int32))
        fp.pcm24_unpack_dev(ctx, jobs)
        for b, e in zip(bufs, expect):
            assert np.array_equal(b.download(np.int32, len(e)), e)

    run_and_check(0)
    run_and_check(0)                                      # identical table: no upload
    run_and_check(1)                                      # same size, other buffers
    img = synth.aiff_image(77, 3000)
    fp.decode_batch(ctx, [img], [fp.probe("aiff", img)])  # the 16-bit path between two unpacks
    wav = synth.wav_image(78, 2000)
    d_img = ctx.to_device(wav)
    d_out = ctx.alloc(4000)
    desc = fp.probe("wav", wav)
    fp.decode_jobs_dev(ctx, [(d_img.ptr + desc.data_off, d_out.ptr, desc.data_len // 2, False)])   # writes scratch slot 8
    run_and_check(1)
    run_and_check(0)


def test_back_to_back_images_are_coalesced_and_still_exact(ctx):
    """an asset directory read into ONE host buffer: consecutive images are adjacent, so their copies travel
    coalesced (payload + the next file's header bytes in one cudaMemcpyAsync); separate buffers in between"""
    rng = np.random.default_rng(99)
    kinds, images = [], []
    for k in range(9):
        kind = "wav" if k % 3 == 0 else "aiff"
        n = int(rng.choice([0, 2, 777, 4096, 100001, 300000]))
        img = synth.wav_image(500 + k, n) if kind == "wav" else synth.aiff_image(500 + k, n)
        if n % 2:
            img = np.concatenate([img, np.array([0x3C], np.uint8)])          # the byte the odd payload reads past its chunk
        kinds.append(kind)
        images.append(img)
    slab = np.concatenate(images[:6])                                        # files 0..5 back to back
    views, off = [], 0
    for img in images[:6]:
        views.append(slab[off:off + img.size])
        off += img.size
    views += images[6:]                                                      # files 6..8 in buffers of their own
    descs = [fp.probe(k, im) for k, im in zip(kinds, views)]
    outs, tracks = fp.decode_batch(ctx, views, descs, keep_on_device=True)
    for k, im, o, t in zip(kinds, views, outs, tracks):
        _, exp = (oracle.wav_parse if k == "wav" else oracle.aiff_parse)(im)
        assert np.array_equal(o, exp), k
        assert np.array_equal(t.download(np.int16, exp.size) if exp.size else exp, exp), k


def test_full_size_c2_batch_properties(ctx):
    """BASELINE config 2 at full size — 1,024 x 2,880,000-byte 24-bit BE AIFF payloads, 1,474,560,000 i16 words — in
    four sub-batches of 256 files: every word equals numpy's big-endian reading of the payload, and a checksum of
    per-file checksums ties the four sub-batches together"""
    n_files, sub = synth.C2_FILES, 256
    hdr = np.frombuffer(synth.aiff_header(synth.C2_DATA_LEN), dtype=np.uint8)
    image_len = hdr.size + synth.C2_DATA_LEN
    total_words, digest = 0, np.uint64(0)
    for b0 in range(0, n_files, sub):
        rng = np.random.default_rng(0xC20000 + b0)
        slab = np.empty(sub * image_len, dtype=np.uint8)              # the sub-batch lies back to back in one buffer
        view = slab.reshape(sub, image_len)
        view[:, :hdr.size] = hdr
        view[:, hdr.size:] = rng.integers(0, 256, size=(sub, synth.C2_DATA_LEN), dtype=np.uint8)
        images = [view[i] for i in range(sub)]
        d = fp.probe("aiff", images[0])
        assert (d.data_off, d.data_len, d.bits_per_sample, d.sample_rate) == (54, synth.C2_DATA_LEN, 24, 48000)
        outs, _ = fp.decode_batch(ctx, images, [d] * sub)
        for i in range(sub):
            exp = view[i, 54:].view(">i2")
            assert outs[i].size == synth.C2_DATA_LEN // 2
            assert np.array_equal(outs[i], exp), (b0, i)
            total_words += outs[i].size
            digest ^= np.uint64(int(outs[i].view(np.uint16).astype(np.uint64).sum()) * (b0 + i + 1) & 0xFFFFFFFFFFFFFFFF)
    assert total_words == 1_474_560_000
    assert int(digest) != 0
