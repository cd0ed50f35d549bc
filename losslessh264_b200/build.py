"""In-tree build of the CUDA shared library for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc")
_OUT = os.path.join(_HERE, "_lib")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def library_path() -> str:
    return os.path.join(_OUT, "liblh264_b200.so")


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build_library(force: bool = False, verbose: bool = False) -> str:
    so = library_path()
    srcs = [os.path.join(_SRC, f) for f in os.listdir(_SRC)] + [os.path.join(os.path.dirname(_HERE), "include", "lh264_b200.h")]
    srcs += [os.path.join(_HERE, "host", f) for f in ("cavlc_tables.h", "cabac_tables.h")]      # tables the decompress kernel includes
    if not force and os.path.exists(so) and os.path.getmtime(so) >= _newest(srcs):
        return so
    os.makedirs(_OUT, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", so, os.path.join(_SRC, "lh264_shim.cu")]
    subprocess.check_call(cmd)
    return so
