"""The OpenH264 decoder API boundary (include/lh264_codec_api.h, libopenh264_lh264.so) exercised the way the reference's own API
tests do it (tests/api/api_driver.cpp, tests/api/c_vtable_check.c):
  CPU  the C view of ISVCDecoder (function table) compiles as C and agrees with the C++ virtual order; struct sizes and `bool`
       agree between the two languages (test/api/c_interface_test.c, cpp_interface_test.cpp); WelsGetDecoderCapability returns
       the reference's constants (test/api/decoder_test.cpp:20-33)
  GPU  BaseDecoderTest's procedure (test/api/BaseDecoderTest.cpp:8-133): units cut at 4-byte start codes -> DecodeFrame2, end of
       stream + DecodeFrame2(NULL, 0), SHA-1 over the cropped rows == the constants of test/api/decoder_test.cpp:90-128."""
import os
import re
import subprocess

import pytest

from streams import LIB_DIR, SHA1, STREAM_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
API = os.path.join(ROOT, "tests", "api")


@pytest.fixture(scope="module")
def api_driver(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("api"))
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "losslessh264_b200", "host")])
    obj, exe = os.path.join(d, "c_vtable_check.o"), os.path.join(d, "api_driver")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-c", "-o", obj, os.path.join(API, "c_vtable_check.c")])      # the header as C
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(API, "api_driver.cpp"), obj, "-L" + LIB_DIR, "-lopenh264_lh264", "-llh264_b200",
                           "-Wl,-rpath," + LIB_DIR])
    return exe


def test_c_and_cpp_views_of_the_decoder_interface_agree(api_driver):
    r = subprocess.run([api_driver, "vtable"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr


def test_initialize_without_a_device_fails_loudly(api_driver):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([api_driver, "cinit"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "Initialize from C: 0" not in r.stdout


@pytest.mark.gpu
def test_decoder_driven_from_c(api_driver):
    r = subprocess.run([api_driver, "cinit"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.fixture(scope="module")
def api_driver_emu(tmp_path_factory):
    """The same driver over the API layer linked against the CPU stand-in of the C ABI (tests/emu/mock_abi.cpp: the kernel sources
    under the CUDA emulator), so that the DecodeFrame2 procedure is checked in the build container too."""
    d = str(tmp_path_factory.mktemp("api_emu"))
    host, emu = os.path.join(ROOT, "losslessh264_b200", "host"), os.path.join(ROOT, "tests", "emu")
    obj, exe = os.path.join(d, "c_vtable_check.o"), os.path.join(d, "api_driver_emu")
    subprocess.check_call(["gcc", "-std=c99", "-c", "-o", obj, os.path.join(API, "c_vtable_check.c")])
    src = ["openh264_api.cpp", "sinks_gpu.cpp", "sinks_file.cpp", "decoder.cpp", "h264_syntax.cpp", "cavlc.cpp", "cavlc_writer.cpp", "pip_container.cpp", "pip_expand.cpp"]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-I" + os.path.join(ROOT, "include"), "-o", exe, os.path.join(API, "api_driver.cpp"), obj]
                          + [os.path.join(host, f) for f in src] + [os.path.join(emu, "mock_abi.cpp"), os.path.join(emu, "cuda_emu.cpp"), "-lpthread"])
    return exe


# NRF_MW_E: pictures that are no reference lose their slot to the next picture - DecodeFrame2 hands them out one call late
@pytest.mark.parametrize("name", ["NRF_MW_E.264", "SVA_BA1_B.264", "test_vd_1d.264", "MIDR_MW_D.264"])
def test_base_decoder_test_procedure_cpu(name, api_driver_emu):
    r = subprocess.run([api_driver_emu, "decode", os.path.join(STREAM_DIR, name)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-400:]
    m = re.search(r"frames (\d+) sha1 (\w+) state_or (\d+)", r.stdout)
    assert m and m.group(2) == SHA1[name]["sha1"] and int(m.group(1)) == SHA1[name]["frames"] and int(m.group(3)) == 0, r.stdout


# streams of the reference's DecoderOutputTest (decoder_test.cpp:90-128); error-free ones: BaseDecoderTest asserts dsErrorFree per call
API_STREAMS = ["BA_MW_D.264", "SVA_BA1_B.264", "MR1_BT_A.h264", "CI1_FT_B.264", "BAMQ1_JVC_C.264", "NRF_MW_E.264", "MIDR_MW_D.264", "test_vd_1d.264",
               "CVPCMNL1_SVA_C.264", "test_qcif_cabac.264"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", API_STREAMS)
def test_base_decoder_test_procedure_gives_the_reference_sha1(name, api_driver):
    assert SHA1[name]["in_decoder_test"]
    r = subprocess.run([api_driver, "decode", os.path.join(STREAM_DIR, name)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-400:]
    m = re.search(r"frames (\d+) sha1 (\w+) state_or (\d+)", r.stdout)
    assert m, r.stdout
    assert m.group(2) == SHA1[name]["sha1"], f"{name}: DecodeFrame2 path gives {m.group(2)}"
    assert int(m.group(3)) == 0, "a DecodeFrame2 call reported an error state on an error-free stream"
    if "frames" in SHA1[name]:
        assert int(m.group(1)) == SHA1[name]["frames"]
