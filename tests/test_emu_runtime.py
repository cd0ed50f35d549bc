"""CPU: the multi-stream runtime (losslessh264_b200/host/runtime.cpp, lh264_rt_* of include/lh264_runtime.h) linked against
tests/emu/mock_abi.cpp — a stand-in for the C ABI that runs the product's kernel SOURCES under the CUDA emulator — instead of
liblh264_b200.so.  What is checked here without a GPU: the parse-thread / submit-thread pipeline (staging buffers, steps parsed
ahead, pictures completed late by damaged streams, geometry changes, looping, model-only mode) delivers, for every stream of a
mixed batch, the reference's YUV SHA1 and the reference's `.pip` file set, and terminates.  The GPU suite (test_gpu_batch.py,
test_gpu_runtime_api.py) runs the same through the real library."""
import hashlib
import json
import os
import re
import subprocess

import pytest

from streams import SHA1, STREAM_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "losslessh264_b200", "host")
EMU = os.path.join(ROOT, "tests", "emu")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pip_tag_sha1.json")))
SRC = ["main_batch.cpp", "runtime.cpp", "decoder.cpp", "h264_syntax.cpp", "cavlc.cpp", "cavlc_writer.cpp", "pip_container.cpp", "pip_expand.cpp"]


@pytest.fixture(scope="module")
def batch_emu(tmp_path_factory):
    exe = os.path.join(str(tmp_path_factory.mktemp("emu")), "batch_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", exe] + [os.path.join(HOST, f) for f in SRC]
                          + [os.path.join(EMU, "mock_abi.cpp"), os.path.join(EMU, "cuda_emu.cpp"), "-lpthread"])
    return exe


def run(exe, names, copies, threads, tmp_path, extra=()):
    d = os.path.join(str(tmp_path), "pip")
    os.makedirs(d, exist_ok=True)
    # LH264_RT_VERIFY_COMPACT: every compact record the runtime hands over is expanded again on the spot and compared with its source
    r = subprocess.run([exe, "--copies", str(copies), "--threads", str(threads), "--pip-dir", d, *extra] + [os.path.join(STREAM_DIR, n) for n in names],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, LH264_RT_VERIFY_COMPACT="1"))
    assert r.returncode == 0, r.stdout[-400:] + r.stderr[-400:]
    lines = r.stdout.strip().splitlines()
    per = {}
    for ln in lines[:-1]:
        m = re.match(r"(\S+) pictures (\d+) yuv_sha1 (\S+) pip_sha1 (\S+)", ln)
        per[os.path.basename(m.group(1))] = (int(m.group(2)), m.group(3), m.group(4))
    return json.loads(lines[-1]), per, d


def check_pip_files(d, idx, name):
    g = GOLD[name]
    pre = f"s{idx}.pip"
    assert hashlib.sha1(open(os.path.join(d, pre), "rb").read()).hexdigest() == g["default"]["sha1"], f"{name}: default stream differs"
    tags = sorted(f[len(pre) + 1:] for f in os.listdir(d) if f.startswith(pre + "."))
    assert tags == sorted(g["tags"]), f"{name}: tag set {tags}"
    for t in tags:
        assert hashlib.sha1(open(os.path.join(d, f"{pre}.{t}"), "rb").read()).hexdigest() == g["tags"][t]["sha1"], f"{name}: .pip.{t} differs"


# damaged streams (pictures completed late / held back), a geometry change mid-stream, multi-slice, CABAC, 8x8 transform
MIX = ["SVA_BA1_B.264", "BA_MW_D_IDR_LOST.264", "BA_MW_D_P_LOST.264", "Error_I_P.264", "syn_qcif_4slices_8f.264", "syn_qcif_cabac_qp38_8f.264", "SVA_CL1_E.264"]


@pytest.mark.parametrize("threads,groups,depth", [(1, 1, 2), (3, 1, 3), (5, 2, 4)])
def test_mixed_batch_matches_reference(batch_emu, tmp_path, threads, groups, depth):
    copies = 2
    summary, per, d = run(batch_emu, MIX, copies, threads, tmp_path, ("--groups", str(groups), "--depth", str(depth)))
    assert summary["copy_mismatches"] == 0 and summary["streams"] == copies * len(MIX)
    for i, n in enumerate(MIX):
        assert per[n][1] == SHA1[n]["sha1"], f"{n}: YUV SHA1 differs from the reference"
        for c in range(copies):
            check_pip_files(d, c * len(MIX) + i, n)


def test_model_only_mode_writes_the_same_pip(batch_emu, tmp_path):
    """`h264dec in.264 out.pip` without a YUV sink: no reconstruction is submitted, the `.pip` bytes are the same."""
    names = ["SVA_BA1_B.264", "syn_qcif_4slices_8f.264"]
    summary, per, d = run(batch_emu, names, 1, 2, tmp_path, ("--no-yuv",))
    assert summary["d2h_bytes"] == 0
    for i, n in enumerate(names):
        assert per[n][1] == "-"
        check_pip_files(d, i, n)


def test_loops_and_picture_limit(batch_emu, tmp_path):
    n = "syn_qcif_4slices_8f.264"
    summary, per, _ = run(batch_emu, [n], 3, 2, tmp_path, ("--loops", "3", "--no-model"))
    assert summary["pictures"] == 3 * 3 * 8 and per[n][0] == 24
    summary, per, _ = run(batch_emu, [n], 2, 2, tmp_path, ("--max-pictures", "5", "--no-model"))
    assert summary["pictures"] == 10 and per[n][0] == 5


def test_no_device_is_a_loud_failure(batch_emu, tmp_path):
    env = dict(os.environ, LH264_MOCK_NO_DEVICE="1")
    r = subprocess.run([batch_emu, os.path.join(STREAM_DIR, "SVA_BA1_B.264")], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
