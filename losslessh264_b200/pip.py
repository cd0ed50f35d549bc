"""`.pip` file sets in the decompress direction: host-side container code (libopenh264_lh264.so: lh264_pip_plan /
lh264_pip_assemble, include/lh264_codec_api.h) around Context.model_decode (the GPU).  Mirrors what the reference's h264dec
does when its input name contains ".pip" (codec/console/dec/src/h264dec.cpp:170-173), for many files per call."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .api import DecodeSlice, Lh264Error
from .records import NUM_TAGS

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libopenh264_lh264.so")
        if not os.path.exists(path):
            raise Lh264Error(f"{path} is not built (make -C losslessh264_b200/host)")
        lib = C.CDLL(path)
        lib.lh264_pip_plan.restype = C.c_void_p
        lib.lh264_pip_plan.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.lh264_pip_plan_free.argtypes = [C.c_void_p]
        lib.lh264_pip_plan_slices.restype = C.POINTER(DecodeSlice)
        lib.lh264_pip_plan_slices.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.lh264_pip_assemble.restype = C.c_longlong
        lib.lh264_pip_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _LIB = lib
    return _LIB


def read_pip(path):
    """(default stream bytes, {tag id: bytes}) of the file set `path`, `path.<tag>`."""
    tags = {}
    for t in range(NUM_TAGS):
        p = f"{path}.{t}"
        if os.path.exists(p):
            tags[t] = open(p, "rb").read()
    return open(path, "rb").read(), tags


class Plan:
    def __init__(self, default_stream: bytes):
        lib = _lib()
        self.default = np.frombuffer(default_stream, np.uint8)
        err = C.create_string_buffer(256)
        self.h = lib.lh264_pip_plan(self.default.ctypes.data, len(default_stream), err, 256)
        if not self.h:
            raise Lh264Error("cannot expand: " + err.value.decode())
        n, mbs, w8 = C.c_int(), C.c_int(), C.c_int()
        self.slices = lib.lh264_pip_plan_slices(self.h, C.byref(n), C.byref(mbs), C.byref(w8))
        self.n_slices, self.max_mbs, self.with_8x8 = n.value, max(mbs.value, 1), bool(w8.value)

    def job(self, tags):
        return {"tags": tags, "slices": self.slices, "n_slices": self.n_slices, "max_mbs": self.max_mbs, "with_8x8": self.with_8x8}

    def assemble(self, bits, outs) -> bytes:
        lib = _lib()
        n = lib.lh264_pip_assemble(self.h, self.default.ctypes.data, bits.ctypes.data, C.cast(outs, C.c_void_p), None, 0)
        if n < 0:
            raise Lh264Error("assembly failed")
        dst = np.empty(max(n, 1), np.uint8)
        lib.lh264_pip_assemble(self.h, self.default.ctypes.data, bits.ctypes.data, C.cast(outs, C.c_void_p), dst.ctypes.data, n)
        return dst[:n].tobytes()

    def close(self):
        if self.h:
            _lib().lh264_pip_plan_free(self.h)
            self.h = None


def expand(ctx, file_sets):
    """file_sets: list of (default stream bytes, {tag: bytes}); returns the original Annex-B files, all streams decoded in ONE
    lh264_model_decode call."""
    plans = [Plan(d) for d, _ in file_sets]
    try:
        res = ctx.model_decode([p.job(t) for p, (_, t) in zip(plans, file_sets)])
        return [p.assemble(bits, outs) for p, (bits, outs) in zip(plans, res)]
    finally:
        for p in plans:
            p.close()


# ---- single-file container `.pipx` (host/pip_container.h) -------------------------------------------------------
def pack(pip_path, archive_path):
    """x.pip + x.pip.<tag> -> one `.pipx` file (lh264_pip_pack)."""
    lib = _lib()
    lib.lh264_pip_pack.argtypes = [C.c_char_p, C.c_char_p]
    rc = lib.lh264_pip_pack(os.fsencode(pip_path), os.fsencode(archive_path))
    if rc:
        raise OSError(f"lh264_pip_pack({pip_path}) failed: {rc}")


def unpack(archive_path, pip_path):
    """`.pipx` -> x.pip + x.pip.<tag> (lh264_pip_unpack); a missing, truncated or corrupt archive raises."""
    lib = _lib()
    lib.lh264_pip_unpack.argtypes = [C.c_char_p, C.c_char_p]
    rc = lib.lh264_pip_unpack(os.fsencode(archive_path), os.fsencode(pip_path))
    if rc:
        raise OSError(f"lh264_pip_unpack({archive_path}) failed: {rc}")


def read_archive(archive_path):
    """(default stream bytes, {tag: bytes}) of a `.pipx` file, parsed here (the format is documented in host/pip_container.h)."""
    import struct

    b = open(archive_path, "rb").read()
    if b[:4] != b"LHPX" or struct.unpack_from("<I", b, 4)[0] != 1:
        raise ValueError("not a .pipx archive")
    n = struct.unpack_from("<I", b, 8)[0]
    at = 12 + 12 * n
    default, tags = None, {}
    for i in range(n):
        tag, length = struct.unpack_from("<IQ", b, 12 + 12 * i)
        payload = b[at : at + length]
        if len(payload) != length:
            raise ValueError("truncated .pipx archive")
        if tag == 0xFFFFFFFF:
            default = payload
        else:
            tags[tag] = payload
        at += length
    if default is None or at != len(b):
        raise ValueError("corrupt .pipx archive")
    return default, tags
