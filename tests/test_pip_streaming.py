"""CPU: streaming flush of the single-file container — `h264dec in.264 out.pipx` appends chunks to the archive
WHILE it compresses (the reference keeps everything until exit, FlushOnClose h264dec.cpp:105-121).  The CLI and the API layer
(host/main_h264dec.cpp, host/openh264_api.cpp) are linked here against the CPU stand-in of the C ABI (tests/emu/mock_abi.cpp,
the kernel sources under the emulation), so the check runs in the build container: whatever the flush interval, the archive
unpacks to the very file set the exit-time path writes — whose bytes are pinned to the reference's by tests/golden/pip_tag_sha1.json
— and every chunk that left early is a prefix that never changed afterwards (a bool coder's pending carry is the risk)."""
import hashlib
import json
import os
import struct
import subprocess

import pytest

from losslessh264_b200 import pip as pipmod
from streams import ROOT, STREAM_DIR

HOST = os.path.join(ROOT, "losslessh264_b200", "host")
EMU = os.path.join(ROOT, "tests", "emu")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "pip_tag_sha1.json")))


@pytest.fixture(scope="module")
def cli_emu(tmp_path_factory):
    exe = os.path.join(str(tmp_path_factory.mktemp("cli")), "h264dec_emu")
    src = ["main_h264dec.cpp", "openh264_api.cpp", "sinks_gpu.cpp", "sinks_file.cpp", "decoder.cpp", "h264_syntax.cpp", "cavlc.cpp", "cavlc_writer.cpp",
           "pip_container.cpp", "pip_expand.cpp"]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-I" + os.path.join(ROOT, "include"), "-o", exe] + [os.path.join(HOST, f) for f in src]
                          + [os.path.join(EMU, "mock_abi.cpp"), os.path.join(EMU, "cuda_emu.cpp"), "-lpthread",
                             "-Wl,--unresolved-symbols=ignore-all"])      # the CLI's decompress leg (lh264_model_decode) is not in the stand-in and not used here
    return exe


def chunks(path):
    b = open(path, "rb").read()
    assert b[:4] == b"LHPX" and struct.unpack_from("<I", b, 4)[0] == 2
    at, out = 8, []
    while True:
        tag, n = struct.unpack_from("<II", b, at)
        at += 8
        if tag == 0xFFFFFFFE:
            assert n == 0 and at == len(b)
            return out
        out.append((tag, b[at:at + n]))
        at += n


@pytest.mark.parametrize("name,units", [("SVA_BA1_B.264", 3), ("SVA_BA1_B.264", 0), ("syn_qcif_cabac_qp38_8f.264", 2), ("sl_ba1_pps_wide.264", 5), ("MR1_BT_A.h264", 7)])
def test_streamed_archive_unpacks_to_the_exit_time_file_set(name, units, cli_emu, tmp_path):
    d = str(tmp_path)
    src = os.path.join(STREAM_DIR, name)
    r = subprocess.run([cli_emu, src, os.path.join(d, "a.pip")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-400:]
    r = subprocess.run([cli_emu, src, os.path.join(d, "b.pipx")], capture_output=True, text=True, timeout=900, env=dict(os.environ, LH264_PIP_FLUSH_UNITS=str(units)))
    assert r.returncode == 0, r.stderr[-400:]
    assert sorted(f for f in os.listdir(d) if f.startswith("b.")) == ["b.pipx"]
    ch = chunks(os.path.join(d, "b.pipx"))
    if units:
        assert sum(1 for t, _ in ch if t == 0xFFFFFFFF) > 2, "the default stream left in one piece: nothing was flushed early"
        assert max(sum(1 for t, _ in ch if t == k) for k in range(64)) > 1, "no side stream left in more than one piece"
    pipmod.unpack(os.path.join(d, "b.pipx"), os.path.join(d, "c.pip"))
    a = {f[1:]: open(os.path.join(d, f), "rb").read() for f in os.listdir(d) if f.startswith("a.pip")}
    c = {f[1:]: open(os.path.join(d, f), "rb").read() for f in os.listdir(d) if f.startswith("c.pip")}
    assert a.keys() == c.keys() and all(a[k] == c[k] for k in a), f"{name}: the streamed archive does not hold the exit-time file set"
    # ... which is the reference's, byte for byte
    want = GOLD[name]
    assert hashlib.sha1(a[".pip"]).hexdigest() == want["default"]["sha1"]
    for tag, sha in want["tags"].items():
        assert hashlib.sha1(a[f".pip.{tag}"]).hexdigest() == sha["sha1"]
    assert len(a) == 1 + len(want["tags"])


def test_archive_that_was_cut_off_is_refused(cli_emu, tmp_path):
    d = str(tmp_path)
    r = subprocess.run([cli_emu, os.path.join(STREAM_DIR, "SVA_BA1_B.264"), os.path.join(d, "b.pipx")], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, LH264_PIP_FLUSH_UNITS="4"))
    assert r.returncode == 0
    b = open(os.path.join(d, "b.pipx"), "rb").read()
    open(os.path.join(d, "cut.pipx"), "wb").write(b[:-8])          # the closing chunk is missing
    with pytest.raises(Exception):
        pipmod.unpack(os.path.join(d, "cut.pipx"), os.path.join(d, "z.pip"))
