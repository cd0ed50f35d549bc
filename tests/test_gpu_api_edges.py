"""GPU: error behaviour and edge cases of the C ABI (include/lh264_b200.h): every misuse returns a negative LH264_E_* code with
a message, nothing crashes, and valid work after an error still produces exact results."""
import ctypes as C

import numpy as np
import pytest

from losslessh264_b200 import api, synth
from losslessh264_b200.records import MODEL_MB, MODEL_SLICE, ModelPicture

pytestmark = pytest.mark.gpu


def test_empty_batches_are_noops(ctx):
    assert ctx.lib.lh264_recon_submit(ctx.h, None, 0) == 0
    assert ctx.lib.lh264_model_submit(ctx.h, None, 0) == 0
    ctx.sync()


def test_bad_arguments_fail_loudly(ctx):
    lib = ctx.lib
    assert lib.lh264_stream_open(ctx.h, 0, 10, 3) < 0
    assert lib.lh264_stream_open(ctx.h, 10, 10, 0) < 0
    assert lib.lh264_stream_close(ctx.h, 12345) < 0
    assert lib.lh264_model_open(ctx.h, 0, 0) < 0
    assert lib.lh264_model_close(ctx.h, 999) < 0
    assert lib.lh264_model_finish(ctx.h, 999) < 0
    assert lib.lh264_model_read(ctx.h, 999, 1, None, 0, 0) < 0
    assert lib.lh264_recon_submit(ctx.h, None, 3) < 0
    assert b"" != lib.lh264_last_error()
    h = C.c_void_p()
    assert lib.lh264_ctx_create(C.byref(h), 9999, None) < 0 and not h.value


def test_closed_stream_and_duplicate_stream_in_batch_are_rejected(ctx):
    mb_w, mb_h = 11, 9
    seq = synth.synthetic_sequence(mb_w, mb_h, n_pictures=2, seed=5)
    s = ctx.open_stream(mb_w, mb_h, 3)
    with pytest.raises(api.Lh264Error):
        ctx.recon([seq[0], seq[0]], [s, s])              # two pictures of one stream in one batch
    ctx.close_stream(s)
    with pytest.raises(api.Lh264Error):
        ctx.recon([seq[0]], [s])                          # closed stream
    # the context still works
    from oracle import oracle

    s2 = ctx.open_stream(mb_w, mb_h, 3)
    ora = oracle.OracleStream(mb_w, mb_h, 3)
    for p in seq:
        ctx.recon([p], [s2])
        assert np.array_equal(ctx.read_picture(s2, p.cur_slot, mb_w, mb_h), ora.recon(p))
    ctx.close_stream(s2)


def test_model_rejects_malformed_pictures_and_oversized_ones(ctx):
    ms = ctx.model_open(99)
    sl = np.zeros(1, MODEL_SLICE)
    sl["n_recs"] = 5                                       # slice claims more records than the picture holds
    rc = np.zeros(2, MODEL_MB)
    with pytest.raises(api.Lh264Error):
        ctx.model_submit([ModelPicture(11, 9, sl, rc, np.zeros((0, 16), np.int16))], [ms])
    sl["n_recs"] = 2
    with pytest.raises(api.Lh264Error):
        ctx.model_submit([ModelPicture(20, 15, sl, rc, np.zeros((0, 16), np.int16))], [ms])   # 300 MBs > 99
    sl["slice_type"] = 1                                   # B slices are not modelled
    with pytest.raises(api.Lh264Error):
        ctx.model_submit([ModelPicture(11, 9, sl, rc, np.zeros((0, 16), np.int16))], [ms])
    ctx.model_close(ms)


def test_model_stream_that_never_coded_anything_has_no_side_streams(ctx):
    ms = ctx.model_open(99)
    assert ctx.model_finish(ms) == {}
    assert ctx.model_decisions(ms) == 0
    ctx.model_close(ms)


def test_all_skipped_picture_and_single_macroblock_picture(ctx):
    """Ragged inputs: a 1x1-MB picture, and a P picture that is one long skip run — CUDA vs oracle, byte for byte."""
    from oracle import oracle

    def run(pics):
        ms = ctx.model_open(max(p.mb_w * p.mb_h for p in pics))
        mo = oracle.ModelOracle()
        for p in pics:
            ctx.model_submit([p], [ms])
            mo.picture(p)
        got, want = ctx.model_finish(ms), mo.finish()
        ctx.model_close(ms)
        mo.close()
        assert got == want

    # one I16x16 macroblock with nothing coded, then the same position skipped
    sl = np.zeros(1, MODEL_SLICE); sl["n_recs"] = 1; sl["slice_type"] = 2; sl["num_ref_idx_l0"] = 1; sl["chroma_format_idc"] = 1
    r = np.zeros(1, MODEL_MB); r["flags"] = 1 | 4 | 8 | (1 << 6); r["type_code"] = 1; r["luma_qp"] = 26; r["qp_delta"] = 26
    p0 = ModelPicture(1, 1, sl.copy(), r.copy(), np.zeros((0, 16), np.int16))
    sl2 = sl.copy(); sl2["slice_type"] = 0; sl2["frame_num"] = 1
    r2 = np.zeros(1, MODEL_MB); r2["flags"] = 1 | 2 | 8 | (1 << 6); r2["type_code"] = 11; r2["skip_run"] = 1
    p1 = ModelPicture(1, 1, sl2, r2, np.zeros((0, 16), np.int16))
    run([p0, p1])
    # 11x9 picture: first an intra picture of empty I16x16 MBs, then one skip run over the whole picture
    n = 99
    sl = np.zeros(1, MODEL_SLICE); sl["n_recs"] = n; sl["slice_type"] = 2; sl["num_ref_idx_l0"] = 1; sl["chroma_format_idc"] = 1
    r = np.zeros(n, MODEL_MB); r["mb_addr"] = np.arange(n); r["mb_in_slice"] = np.arange(n); r["flags"] = 1 | 4 | (1 << 6); r["type_code"] = 1; r["luma_qp"] = 30
    r["flags"][-1] |= 8; r["qp_delta"][0] = 30
    a = ModelPicture(11, 9, sl, r, np.zeros((0, 16), np.int16))
    sl2 = sl.copy(); sl2["slice_type"] = 0; sl2["frame_num"] = 1
    r2 = np.zeros(n, MODEL_MB); r2["mb_addr"] = np.arange(n); r2["mb_in_slice"] = np.arange(n); r2["type_code"] = 11; r2["flags"] = 1 << 6
    r2["flags"][0] |= 1; r2["skip_run"][0] = n; r2["flags"][-1] |= 2 | 8
    b = ModelPicture(11, 9, sl2, r2, np.zeros((0, 16), np.int16))
    run([a, b, b])
