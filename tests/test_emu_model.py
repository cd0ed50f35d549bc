"""CPU: the SOURCES of the recompression-model kernels (losslessh264_b200/csrc/model_kernels.cuh) compiled with g++ behind
tests/emu/cuda_emu.h and launched as model_host.inl launches them — generation per picture (k_model_frame / count / scan / plan /
emit / commit), then ONE deferred k_model_code launch per LH264_CODE_STEPS pictures — on whole streams; the `.pip.<tag>` side
streams must equal, byte for byte, those of the compiled reference (tests/golden/pip_tag_sha1.json).  Test infrastructure: the
product has no CPU path."""
import hashlib
import json
import os
import subprocess

import pytest

from streams import STREAM_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "losslessh264_b200", "host")
EMU = os.path.join(ROOT, "tests", "emu")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pip_tag_sha1.json")))
CORE = ["decoder.cpp", "h264_syntax.cpp", "cavlc.cpp", "cavlc_writer.cpp", "pip_container.cpp", "pip_expand.cpp"]
# (stream, pictures coded per k_model_code launch): 1 = a launch per picture, as before coding was deferred
CASES = [("tibby.264", 1), ("tibby.264", 7), ("SVA_BA1_B.264", 3), ("syn_qcif_4slices_8f.264", 16), ("tibby8x8cavlc.264", 5), ("test_qcif_cabac.264", 4),
         ("MR1_BT_A.h264", 6), ("BA_MW_D_IDR_LOST.264", 2), ("Error_I_P.264", 4), ("QCIF_2P_I_allIPCM.264", 3)]


@pytest.fixture(scope="module")
def model_emu(tmp_path_factory):
    exe = os.path.join(str(tmp_path_factory.mktemp("emu")), "model_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", exe, os.path.join(EMU, "model_emu.cpp"), os.path.join(EMU, "cuda_emu.cpp")]
                          + [os.path.join(HOST, f) for f in CORE])
    return exe


@pytest.mark.parametrize("name,steps", CASES)
def test_emulated_model_kernels_give_the_reference_side_streams(name, steps, model_emu, tmp_path):
    out = os.path.join(str(tmp_path), "o.pip")
    r = subprocess.run([model_emu, os.path.join(STREAM_DIR, name), out], capture_output=True, text=True, timeout=900, env=dict(os.environ, LH264_CODE_STEPS=str(steps)))
    assert r.returncode == 0, r.stderr[-400:]
    g = GOLD[name]
    assert hashlib.sha1(open(out, "rb").read()).hexdigest() == g["default"]["sha1"], f"{name}: default stream differs"
    tags = sorted(f[len("o.pip") + 1:] for f in os.listdir(str(tmp_path)) if f.startswith("o.pip."))
    assert tags == sorted(g["tags"]), f"{name}: tag set {tags}"
    for t in tags:
        assert hashlib.sha1(open(f"{out}.{t}", "rb").read()).hexdigest() == g["tags"][t]["sha1"], f"{name}: .pip.{t} differs (code steps {steps})"
