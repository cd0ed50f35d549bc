"""CPU: the compact transfer form of the hand-off records (lh264_cmb / lh264_cmm, include/lh264_b200.h; mapping in
losslessh264_b200/csrc/compact_records.h).  The runtime compacts every picture's MB records and model records before they cross
host memory and PCIe, the device expands them (k_expand_mb / k_expand_mm) in front of the same kernels: expansion must give back
the parser's records BYTE FOR BYTE.  With LH264_RT_VERIFY_COMPACT the runtime expands every record it has just compacted with the
same functions the kernels run and aborts on a difference; LH264_RT_PARSE_ONLY leaves the (emulated) kernels out, so the whole
corpus goes through in seconds.  test_emu_runtime.py runs the kernels behind the expansion on a mixed batch."""
import ctypes as C
import json
import os
import subprocess

import pytest

from streams import STREAM_DIR, decodable_streams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "losslessh264_b200", "host")
EMU = os.path.join(ROOT, "tests", "emu")
SRC = ["main_batch.cpp", "runtime.cpp", "decoder.cpp", "h264_syntax.cpp", "cavlc.cpp", "cavlc_writer.cpp", "pip_container.cpp", "pip_expand.cpp"]


@pytest.fixture(scope="module")
def batch_emu(tmp_path_factory):
    exe = os.path.join(str(tmp_path_factory.mktemp("emu")), "batch_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", exe] + [os.path.join(HOST, f) for f in SRC]
                          + [os.path.join(EMU, "mock_abi.cpp"), os.path.join(EMU, "cuda_emu.cpp"), "-lpthread"])
    return exe


def test_every_record_of_the_corpus_survives_compaction(batch_emu):
    names = decodable_streams()
    env = dict(os.environ, LH264_RT_VERIFY_COMPACT="1", LH264_RT_PARSE_ONLY="1")
    r = subprocess.run([batch_emu, "--threads", "8", "--no-hash"] + [os.path.join(STREAM_DIR, n) for n in names], capture_output=True, text=True, timeout=1200, env=env)
    assert r.returncode == 0, r.stderr[-600:]
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    assert summary["streams"] == len(names) and summary["pictures"] > 3000
    # what the compact form saves on this corpus (mixed content: many intra and 8x8-partition macroblocks)
    assert summary["h2d_bytes"] > 0


def test_struct_sizes():
    from losslessh264_b200.build import build_library

    lib = C.CDLL(build_library())
    assert lib.lh264_abi_version() == 3
    src = open(os.path.join(ROOT, "include", "lh264_b200.h")).read()
    assert "typedef struct lh264_cmb" in src and "typedef struct lh264_cmm" in src
    # sizes are part of the ABI: checked where the structs are compiled
    code = r'''
#include "%s"
static_assert (sizeof (lh264_cmb) == 32, "lh264_cmb");
static_assert (sizeof (lh264_cmm) == 48, "lh264_cmm");
static_assert (sizeof (lh264_mb_rec) == 128 && sizeof (lh264_model_mb) == 128, "records");
int main () { return 0; }
''' % os.path.join(ROOT, "include", "lh264_b200.h")
    r = subprocess.run(["g++", "-std=c++17", "-x", "c++", "-fsyntax-only", "-"], input=code, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr
