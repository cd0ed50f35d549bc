"""GPU: the CUDA recompression model (lh264_model_* through the C ABI) against the `.pip.<tag>` side streams of
the compiled reference (golden SHA1s) and against the sequential oracle.  Byte-exact."""
import os

import pytest

from test_model_oracle import GOLD, check_tags, model_pictures
from streams import decodable_streams

pytestmark = pytest.mark.gpu

ALL = [s for s in decodable_streams() if s in GOLD and "tags" in GOLD[s]]


def run_cuda(ctx, pics, with_8x8):
    ms = ctx.model_open(max(p.mb_w * p.mb_h for p in pics), with_8x8)
    try:
        for p in pics:
            ctx.model_submit([p], [ms])
        return ctx.model_finish(ms), ctx.model_decisions(ms)
    finally:
        ctx.model_close(ms)


@pytest.mark.parametrize("name", ALL)
def test_cuda_model_matches_reference_pip_tags(ctx, name, tmp_path):
    pics = model_pictures(name, str(tmp_path))
    got, _ = run_cuda(ctx, pics, "8x8" in name)
    check_tags(name, got)


def test_cuda_model_batches_independent_streams(ctx, tmp_path):
    """Several streams of different length and size coded together, one picture per stream per submit."""
    from oracle import oracle

    names = ["BA_MW_D.264", "tibby.264", "SVA_FM1_E.264", "black.264", "CVFC1_Sony_C.jsv"]
    seqs = [model_pictures(n, str(tmp_path)) for n in names]
    handles = [ctx.model_open(max(p.mb_w * p.mb_h for p in s)) for s in seqs]
    try:
        for t in range(max(len(s) for s in seqs)):
            live = [k for k in range(len(seqs)) if t < len(seqs[k])]
            ctx.model_submit([seqs[k][t] for k in live], [handles[k] for k in live])
        for k, n in enumerate(names):
            check_tags(n, ctx.model_finish(handles[k]))
            mo = oracle.ModelOracle()
            for p in seqs[k]:
                mo.picture(p)
            mo.finish()
            assert mo.decisions() == ctx.model_decisions(handles[k]), f"{n}: decision count differs from the oracle"
            mo.close()
    finally:
        for h in handles:
            ctx.model_close(h)


def test_pipelined_submits_survive_a_batch_that_does_not_fit(tmp_path):
    """Two batches in flight, no host wait between submits: the first picture overflows the initial decision list, so it skips
    its coding on the device and the batch behind it skips entirely; the library grows and runs both again in order."""
    from losslessh264_b200.api import Context

    name = "syn720p_8f.264"
    pics = model_pictures(name, str(tmp_path))
    c = Context(0)                                          # a fresh context: its decision list starts at the first guess
    try:
        ms = c.model_open(max(p.mb_w * p.mb_h for p in pics))
        for p in pics:
            c.model_submit([p], [ms], sync=False)
        check_tags(name, c.model_finish(ms))
        c.model_close(ms)
    finally:
        c.close()
