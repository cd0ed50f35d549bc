"""CPU: the SOURCES of the reconstruction + deblocking kernel (losslessh264_b200/csrc/decode_rows.cuh and everything it includes)
compiled with g++ behind tests/emu/cuda_emu.h — a fiber-per-CUDA-thread emulation of warps, shuffles, ballots, shared memory and
the acquire / release row protocol — and run on whole streams: host parse -> hand-off records -> emulated k_decode_rows -> cropped
I420, whose SHA1 must equal the reference's (tests/golden/stream_sha1.json; 36 of them are the constants of the reference's
test/api/decoder_test.cpp:90-128).

This is test infrastructure (the product has no CPU path, and nothing under losslessh264_b200/ knows about it): it lets the
build container, which has no GPU, check the kernel logic bit for bit before the GPU suite does; memory-model races, alignment
and bank conflicts remain the GPU suite's business."""
import hashlib
import os
import subprocess

import pytest

from streams import SHA1, STREAM_DIR, decodable_streams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "losslessh264_b200", "host")
EMU = os.path.join(ROOT, "tests", "emu")
CORE = ["decoder.cpp", "h264_syntax.cpp", "cavlc.cpp", "cavlc_writer.cpp", "pip_container.cpp", "pip_expand.cpp"]

# the long streams stay with the GPU suite (the emulation runs ~0.1-1 M macroblocks/s)
SKIP = {"LS_SVA_D.264", "BA1_FT_C.264", "CI1_FT_B.264", "test_cif_I_CABAC_slice.264", "test_cif_P_CABAC_slice.264", "syn4k_8slices_ltr_4f.264",
        "MR2_MW_A.264", "MR2_TANDBERG_E.264", "black.264", "tibbycabac.264"}
NAMES = [s for s in decodable_streams() if s not in SKIP]


@pytest.fixture(scope="module")
def recon_emu(tmp_path_factory):
    exe = os.path.join(str(tmp_path_factory.mktemp("emu")), "recon_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", exe, os.path.join(EMU, "recon_emu.cpp"), os.path.join(EMU, "cuda_emu.cpp")]
                          + [os.path.join(HOST, f) for f in CORE])
    return exe


@pytest.mark.parametrize("name", NAMES)
def test_emulated_kernel_sources_give_the_reference_sha1(name, recon_emu, tmp_path):
    out = os.path.join(str(tmp_path), "o.yuv")
    # 3 warps in the (single) emulated CTA: rows of one picture are in flight together, so the inter-row waits are exercised
    r = subprocess.run([recon_emu, os.path.join(STREAM_DIR, name), out, "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    assert "record validation failed" not in r.stderr
    sha = hashlib.sha1(open(out, "rb").read()).hexdigest()
    assert sha == SHA1[name]["sha1"], f"{name}: emulated k_decode_rows gives {sha}, the reference {SHA1[name]['sha1']}"


def test_malformed_records_are_flagged_not_dereferenced(recon_emu, tmp_path):
    """A record whose slice_idx / coef_off point outside the picture's tables must be skipped and reported (the C ABI documents
    LH264_E_INVALID for malformed records); the harness corrupts one record per picture when asked to."""
    out = os.path.join(str(tmp_path), "o.yuv")
    env = dict(os.environ, LH264_EMU_CORRUPT="1")
    r = subprocess.run([recon_emu, os.path.join(STREAM_DIR, "SVA_BA1_B.264"), out, "2"], capture_output=True, text=True, timeout=600, env=env)
    assert "record validation failed" in r.stderr
