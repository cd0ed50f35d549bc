"""GPU: the CUDA hot path through the C ABI against (a) the reference's own pictures (golden
fixtures) and (b) the oracle on seeded synthetic batches.  Bit-exact."""
import glob
import os

import numpy as np
import pytest

from losslessh264_b200.records import load_fixture, n_slots_needed
from losslessh264_b200 import synth

pytestmark = pytest.mark.gpu

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_cuda_matches_reference_pictures(ctx, path):
    pics = load_fixture(path)
    mb_w, mb_h = pics[0].mb_w, pics[0].mb_h
    s = ctx.open_stream(mb_w, mb_h, n_slots_needed(pics))
    try:
        for i, p in enumerate(pics):
            ctx.recon([p], [s])
            out = ctx.read_picture(s, p.cur_slot, mb_w, mb_h)
            ref = p.final_yuv()
            if not np.array_equal(out, ref):
                d = np.nonzero(out != ref)[0]
                raise AssertionError(f"{os.path.basename(path)} picture {i}: {len(d)} samples differ, first at {d[0]}")
    finally:
        ctx.close_stream(s)


@pytest.mark.parametrize("mb_w,mb_h,seed", [(11, 9, 1), (20, 15, 2), (120, 68, 3)])
def test_cuda_matches_oracle_on_synthetic_batches(ctx, mb_w, mb_h, seed):
    from oracle import oracle

    n_streams = 3 if mb_w < 100 else 2
    seqs = [synth.synthetic_sequence(mb_w, mb_h, n_pictures=3, seed=seed * 100 + k) for k in range(n_streams)]
    ora = [oracle.OracleStream(mb_w, mb_h, 3) for _ in range(n_streams)]
    streams = [ctx.open_stream(mb_w, mb_h, 3) for _ in range(n_streams)]
    try:
        for t in range(3):
            batch = [seq[t] for seq in seqs]
            ctx.recon(batch, streams)
            for k in range(n_streams):
                want = ora[k].recon(batch[k])
                got = ctx.read_picture(streams[k], batch[k].cur_slot, mb_w, mb_h)
                assert np.array_equal(got, want), f"stream {k} picture {t} differs from the oracle"
    finally:
        for s in streams:
            ctx.close_stream(s)


def test_kernel_idct_mc_expand_match_oracle(ctx):
    from oracle import oracle

    rng = np.random.default_rng(7)
    pix = rng.integers(0, 256, (1000, 16), dtype=np.uint8)
    coef = rng.integers(-2048, 2048, (1000, 16)).astype(np.int16)  # 12-bit, as DecUT_IdctResAddPred.cpp
    assert np.array_equal(ctx.k_idct4x4_add(pix, coef), oracle.idct4x4_add(pix, coef))
    coef[:10] = rng.integers(-32768, 32768, (10, 16)).astype(np.int16)  # int16 wrap of the row pass
    assert np.array_equal(ctx.k_idct4x4_add(pix, coef), oracle.idct4x4_add(pix, coef))
    pix8 = rng.integers(0, 256, (300, 64), dtype=np.uint8)
    coef8 = rng.integers(-2048, 2048, (300, 64)).astype(np.int16)
    assert np.array_equal(ctx.k_idct8x8_add(pix8, coef8), oracle.idct8x8_add(pix8, coef8))

    ref = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    blocks = []
    for _ in range(400):
        bw, bh = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)][rng.integers(7)]
        blocks.append([rng.integers(-8, 60), rng.integers(-8, 44), rng.integers(-80, 80), rng.integers(-80, 80), bw, bh])
    blocks = np.array(blocks, np.int32)
    got = ctx.k_mc(ref, blocks, luma=True)
    for b, g in zip(blocks, got):
        want = oracle.mc_luma(ref, *[int(v) for v in b])
        assert np.array_equal(g[: b[5], : b[4]], want)
    cblocks = blocks.copy()
    cblocks[:, 4:] //= 2
    got = ctx.k_mc(ref, cblocks, luma=False)
    for b, g in zip(cblocks, got):
        want = oracle.mc_chroma(ref, *[int(v) for v in b])
        assert np.array_equal(g[: b[5], : b[4]], want)
    assert np.array_equal(ctx.k_expand_plane(ref, 32), oracle.expand_plane(ref, 32))
    assert np.array_equal(ctx.k_expand_plane(ref[:24, :32], 16), oracle.expand_plane(ref[:24, :32], 16))


def test_prepared_descriptor_arrays_submit_the_same_work(ctx):
    """Context.recon_descs / recon_prepared (descriptor arrays built once, submitted many times — what bench.py's host-buffer loop
    does) must give the pictures of the plain recon() call."""
    mb_w, mb_h = 20, 15
    seq = synth.synthetic_sequence(mb_w, mb_h, n_pictures=3, seed=77)
    a, b = ctx.open_stream(mb_w, mb_h, 3), ctx.open_stream(mb_w, mb_h, 3)
    try:
        prepared = [ctx.recon_descs([p], [b]) for p in seq]
        for rep in range(2):                              # the same arrays again: a second pass over the sequence
            for p, d in zip(seq, prepared):
                ctx.recon([p], [a])
                ctx.recon_prepared(d)
                assert np.array_equal(ctx.read_picture(a, p.cur_slot, mb_w, mb_h), ctx.read_picture(b, p.cur_slot, mb_w, mb_h))
    finally:
        ctx.close_stream(a)
        ctx.close_stream(b)
