"""GPU: decompress direction through the C ABI — lh264_model_decode (k_model_decode: one warp per stream walks
the side streams back through the adaptive priors and re-serialises CAVLC) between the host container code.
Checks: (1) the full CLI round trip on the GPU, both directions (h264dec_lh264 in.264 x.pip ; h264dec_lh264 x.pip out.264)
gives back the original file byte for byte on every CAVLC stream of the corpus; (2) many streams in ONE call give the same
slice bits as the CPU build of the same core (oracle/expand_cpu) and the original files; (3) error behaviour."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle
from streams import LIB_DIR, STREAM_DIR
from test_pip_expand import CPU_SET
from test_pip_expand import write_pip_cpu

pytestmark = pytest.mark.gpu
CLI = os.path.join(LIB_DIR, "h264dec_lh264")
ALL = sorted(CPU_SET + ["LS_SVA_D.264", "BA1_FT_C.264", "CI1_FT_B.264"])
# the two-process CLI round trip costs seconds of context set-up per file: a representative subset; every stream of the corpus goes
# through the same kernel in test_every_stream_in_one_call below
CLI_SET = ["tibby.264", "tibby8x8cavlc.264", "BA1_FT_C.264", "LS_SVA_D.264", "black.264", "BAMQ2_JVC_C.264", "MR2_TANDBERG_E.264", "SVA_FM1_E.264",
           "syn_qcif_4slices_8f.264", "syn4k_8slices_ltr_4f.264", "CVFC1_Sony_C.jsv", "MR1_BT_A.h264", "tibbycabac.264", "syn_cif_cabac_4slices_ltr_6f.264"]


@pytest.mark.parametrize("name", CLI_SET)
def test_cli_round_trip_on_the_gpu(name, tmp_path):
    src = os.path.join(STREAM_DIR, name)
    pip = os.path.join(str(tmp_path), "o.pip")
    r = subprocess.run([CLI, src, pip], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    back = os.path.join(str(tmp_path), "back.264")
    r = subprocess.run([CLI, pip, back], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    assert open(back, "rb").read() == open(src, "rb").read(), f"{name}: .264 -> .pip -> .264 on the GPU is not the identity"


def test_every_stream_in_one_call(ctx, tmp_path):
    from losslessh264_b200 import pip as pipmod

    names = ALL
    sets = []
    for i, n in enumerate(names):
        d = os.path.join(str(tmp_path), str(i))
        os.makedirs(d)
        sets.append(pipmod.read_pip(write_pip_cpu(n, d)))
    got = pipmod.expand(ctx, sets)
    for n, g in zip(names, got):
        assert g == open(os.path.join(STREAM_DIR, n), "rb").read(), f"{n}: expanded file differs from the original"
    assert ctx.lib.lh264_kernel_ms(ctx.h, 6) >= 0.0


def test_same_stream_many_times_is_deterministic(ctx, tmp_path):
    from losslessh264_b200 import pip as pipmod

    name = "tibby8x8cavlc.264"
    fs = pipmod.read_pip(write_pip_cpu(name, str(tmp_path)))
    got = pipmod.expand(ctx, [fs] * 40)
    want = open(os.path.join(STREAM_DIR, name), "rb").read()
    assert all(g == want for g in got)


def test_output_buffer_too_small_is_reported_and_retried(ctx, tmp_path):
    from losslessh264_b200 import api, pip as pipmod

    name = "syn720p_8f.264"
    d, tags = pipmod.read_pip(write_pip_cpu(name, str(tmp_path)))
    plan = pipmod.Plan(d)
    # retry path of Context.model_decode: starts at 4 KB, doubles until the job fits
    (bits, outs), = ctx.model_decode([plan.job(tags)], out_capacity=4096)
    assert plan.assemble(bits, outs) == open(os.path.join(STREAM_DIR, name), "rb").read()
    plan.close()


def test_pcm_stream_is_refused_like_the_reference(tmp_path):
    src = os.path.join(STREAM_DIR, "CVPCMNL1_SVA_C.264")
    pip = os.path.join(str(tmp_path), "o.pip")
    assert subprocess.run([CLI, src, pip], capture_output=True).returncode == 0
    r = subprocess.run([CLI, pip, os.path.join(str(tmp_path), "b.264")], capture_output=True, text=True)
    assert r.returncode != 0 and "21" in r.stderr          # I_PCM samples are not in the model (reference: same limitation)


@pytest.mark.parametrize("name", ["SVA_BA1_B.264", "test_qcif_cabac.264"])
def test_corrupt_side_stream_does_not_hang_or_pass(name, ctx, tmp_path):
    from losslessh264_b200 import api, pip as pipmod

    d, tags = pipmod.read_pip(write_pip_cpu(name, str(tmp_path)))
    big = max(tags, key=lambda t: len(tags[t]))
    rng = np.random.default_rng(5)
    tags[big] = rng.integers(0, 256, len(tags[big]), dtype=np.uint8).tobytes()
    plan = pipmod.Plan(d)
    try:
        (bits, outs), = ctx.model_decode([plan.job(tags)])
        out = plan.assemble(bits, outs)
    except api.Lh264Error:
        out = None
    assert out != open(os.path.join(STREAM_DIR, name), "rb").read()
    plan.close()


def test_cli_single_file_container_round_trip(tmp_path):
    """`.pipx` in place of `.pip`: compress writes ONE file, decompress reads it back to the original."""
    src = os.path.join(STREAM_DIR, "tibby.264")
    arc = os.path.join(str(tmp_path), "o.pipx")
    back = os.path.join(str(tmp_path), "b.264")
    r = subprocess.run([CLI, src, arc], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    assert os.listdir(str(tmp_path)) == ["o.pipx"]
    r = subprocess.run([CLI, arc, back], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    assert open(back, "rb").read() == open(src, "rb").read()
