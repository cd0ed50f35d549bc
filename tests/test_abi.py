"""CPU: the C-ABI library loads and exports every symbol include/lh264_b200.h declares (no compute calls), the
Python record layouts match the C structs, and the product fails loudly without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from losslessh264_b200 import api
from losslessh264_b200.build import build_library
from losslessh264_b200.records import MB_REC, MODEL_MB, MODEL_SLICE, SLICE_REC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lh264_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lh264_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(build_library())
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/lh264_b200.h but not exported"
    assert sorted(api.EXPORTS) == names, "api.EXPORTS is out of date with the header"


def test_openh264_api_library_exports_decoder_symbols():
    import subprocess

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "losslessh264_b200", "host")])
    lib = C.CDLL(os.path.join(ROOT, "losslessh264_b200", "_lib", "libopenh264_lh264.so"))
    for n in ("WelsCreateDecoder", "WelsDestroyDecoder", "WelsGetDecoderCapability", "WelsGetCodecVersion", "WelsGetCodecVersionEx"):
        assert hasattr(lib, n)
    # ... and every lh264_pip_* entry point include/lh264_codec_api.h declares (container code of both directions)
    src = open(os.path.join(ROOT, "include", "lh264_codec_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(lh264_pip_[a-z0-9_]+)\s*\(", src)))
    assert names == ["lh264_pip_assemble", "lh264_pip_pack", "lh264_pip_plan", "lh264_pip_plan_free", "lh264_pip_plan_slices", "lh264_pip_unpack"]
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/lh264_codec_api.h but not exported"


def test_runtime_entry_points_are_exported():
    """include/lh264_runtime.h: the multi-stream runtime lives behind the library boundary (libopenh264_lh264.so)."""
    import subprocess

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "losslessh264_b200", "host")])
    lib = C.CDLL(os.path.join(ROOT, "losslessh264_b200", "_lib", "libopenh264_lh264.so"))
    src = open(os.path.join(ROOT, "include", "lh264_runtime.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(lh264_rt_[a-z0-9_]+)\s*\(", src)) - {"lh264_rt_picture_fn"})
    assert names == ["lh264_rt_add_stream", "lh264_rt_create", "lh264_rt_destroy", "lh264_rt_last_error", "lh264_rt_pip_read", "lh264_rt_run",
                     "lh264_rt_stats_get", "lh264_rt_stream_result_get", "lh264_rt_trim"]
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/lh264_runtime.h but not exported"
    # the ctypes mirror of the structs has the C layout
    from losslessh264_b200.runtime import RtConfig, RtStats, RtStreamResult

    assert C.sizeof(RtConfig) == 64 and C.sizeof(RtStreamResult) == 2 * 8 + 82 + 6 + 8 + 8 and C.sizeof(RtStats) == 18 * 8


def test_record_layouts():
    assert MB_REC.itemsize == 128 and MODEL_MB.itemsize == 128 and MODEL_SLICE.itemsize == 20
    assert SLICE_REC.itemsize == 16 + 2 * 16 * 7 + 240          # + scaling lists (flags, 6 x 16, 2 x 64)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.Lh264Error, match="no CUDA device|CPU fallback|-2"):
        api.Context(0)
