"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product package (losslessh264_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, "recon_oracle.c"), os.path.join(_HERE, "model_oracle.c")]
    dep = max(os.path.getmtime(p) for p in srcs + [os.path.join(os.path.dirname(_HERE), "include", "lh264_b200.h")])
    if force or not os.path.exists(so) or os.path.getmtime(so) < dep:
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", so] + srcs)
    return so


def build_expand_checker(force=False):
    """oracle/expand_cpu: CPU build of the decompress core the CUDA kernel runs (oracle/model_decode_cpu.cpp)."""
    root = os.path.dirname(_HERE)
    host = os.path.join(root, "losslessh264_b200", "host")
    exe = os.path.join(_HERE, "expand_cpu")
    srcs = [os.path.join(_HERE, "model_decode_cpu.cpp")] + [os.path.join(host, f) for f in ("pip_expand.cpp", "pip_container.cpp", "h264_syntax.cpp")]
    deps = srcs + [os.path.join(root, "include", "lh264_b200.h"), os.path.join(host, "cavlc_tables.h"), os.path.join(host, "pip_expand.h"),
                   os.path.join(root, "losslessh264_b200", "csrc", "model_decode_core.inl"), os.path.join(root, "losslessh264_b200", "csrc", "model_layout.h")]
    dep = max(os.path.getmtime(p) for p in deps)
    if force or not os.path.exists(exe) or os.path.getmtime(exe) < dep:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe] + srcs)
    return exe


class _Pic(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("w", C.c_int), ("h", C.c_int)]


class _Desc(C.Structure):
    _fields_ = [
        ("stream", C.c_int32),
        ("cur_slot", C.c_int32),
        ("n_slices", C.c_int32),
        ("n_coef_blocks", C.c_int32),
        ("slices", C.c_void_p),
        ("mbs", C.c_void_p),
        ("coefs", C.c_void_p),
    ]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ora_recon_picture.restype = C.c_int
        _LIB.mo_open.restype = C.c_void_p
        _LIB.mo_close.argtypes = [C.c_void_p]
        _LIB.mo_picture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.mo_finish.argtypes = [C.c_void_p]
        _LIB.mo_tag_bytes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        _LIB.mo_tag_bytes.restype = C.c_long
        _LIB.mo_decisions.argtypes = [C.c_void_p]
        _LIB.mo_decisions.restype = C.c_ulonglong
    return _LIB


class OracleStream:
    """Picture slots of one stream on the CPU; mirrors lh264_stream_open()."""

    def __init__(self, mb_w, mb_h, n_slots):
        self.mb_w, self.mb_h, self.n_slots = mb_w, mb_h, n_slots
        self.buf = [np.full(mb_w * mb_h * 384, 128, np.uint8) for _ in range(n_slots)]
        self._pics = (_Pic * n_slots)()
        ysz = mb_w * mb_h * 256
        for i, b in enumerate(self.buf):
            base = b.ctypes.data
            self._pics[i].plane[0] = base
            self._pics[i].plane[1] = base + ysz
            self._pics[i].plane[2] = base + ysz + ysz // 4
            self._pics[i].w = mb_w * 16
            self._pics[i].h = mb_h * 16

    def recon(self, pic):
        """pic: losslessh264_b200.records.Picture.  Returns the reconstructed I420 (flat uint8)."""
        slices = np.ascontiguousarray(pic.slices)
        mbs = np.ascontiguousarray(pic.mbs)
        coefs = np.ascontiguousarray(pic.coefs)
        d = _Desc(0, pic.cur_slot, len(slices), len(coefs), slices.ctypes.data, mbs.ctypes.data, coefs.ctypes.data)
        rc = lib().ora_recon_picture(C.byref(d), self.mb_w, self.mb_h, self._pics, self.n_slots)
        if rc != 0:
            raise RuntimeError(f"ora_recon_picture failed: {rc}")
        return self.buf[pic.cur_slot]


def mc_luma(ref, x, y, mvx, mvy, bw, bh):
    ref = np.ascontiguousarray(ref, np.uint8)
    h, w = ref.shape
    out = np.zeros((bh, bw), np.uint8)
    lib().ora_mc_luma(ref.ctypes.data_as(C.c_void_p), w, h, x, y, mvx, mvy, bw, bh, out.ctypes.data_as(C.c_void_p), bw)
    return out


def mc_chroma(ref, x, y, mvx, mvy, bw, bh):
    ref = np.ascontiguousarray(ref, np.uint8)
    h, w = ref.shape
    out = np.zeros((bh, bw), np.uint8)
    lib().ora_mc_chroma(ref.ctypes.data_as(C.c_void_p), w, h, x, y, mvx, mvy, bw, bh, out.ctypes.data_as(C.c_void_p), bw)
    return out


def idct4x4_add(pix, coef):
    pix = np.ascontiguousarray(pix, np.uint8).copy()
    coef = np.ascontiguousarray(coef, np.int16)
    for i in range(len(pix)):
        lib().ora_idct4x4_add(C.c_void_p(pix[i].ctypes.data), 4, C.c_void_p(coef[i].ctypes.data))
    return pix


def idct8x8_add(pix, coef):
    pix = np.ascontiguousarray(pix, np.uint8).copy()
    coef = np.ascontiguousarray(coef, np.int16)
    for i in range(len(pix)):
        lib().ora_idct8x8_add(C.c_void_p(pix[i].ctypes.data), 8, C.c_void_p(coef[i].ctypes.data))
    return pix


def expand_plane(plane, pad):
    plane = np.ascontiguousarray(plane, np.uint8)
    h, w = plane.shape
    out = np.zeros((h + 2 * pad, w + 2 * pad), np.uint8)
    lib().ora_expand_plane(C.c_void_p(plane.ctypes.data), w, h, pad, C.c_void_p(out.ctypes.data))
    return out


class ModelOracle:
    """Sequential CPU restatement of the recompression model for one stream (oracle/model_oracle.c)."""

    def __init__(self):
        self._h = lib().mo_open()

    def picture(self, mp):
        """mp: losslessh264_b200.records.ModelPicture."""
        sl = np.ascontiguousarray(mp.slices)
        rc = np.ascontiguousarray(mp.recs)
        co = np.ascontiguousarray(mp.coefs)
        r = lib().mo_picture(self._h, mp.mb_w, mp.mb_h, len(sl), sl.ctypes.data, rc.ctypes.data, co.ctypes.data)
        if r != 0:
            raise RuntimeError("mo_picture: index outside the reference's prior tables (the reference asserts here)")

    def finish(self):
        """Ends every tag stream; returns {tag: bytes} for the tags that were used."""
        lib().mo_finish(self._h)
        out = {}
        for t in range(70):
            p = C.c_void_p()
            n = lib().mo_tag_bytes(self._h, t, C.byref(p))
            if n >= 0:
                out[t] = C.string_at(p, n)
        return out

    def decisions(self):
        return int(lib().mo_decisions(self._h))

    def close(self):
        if self._h:
            lib().mo_close(self._h)
            self._h = None

    def __del__(self):
        self.close()
