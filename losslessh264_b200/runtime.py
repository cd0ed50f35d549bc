"""ctypes binding of the multi-stream runtime (include/lh264_runtime.h, host/runtime.cpp in libopenh264_lh264.so): Annex-B
bytes of many independent streams in, decoded I420 pictures + `.pip` file sets out — S concurrent `h264dec in.264 out.pip
out.yuv` runs (codec/console/dec/src/h264dec.cpp:246-364) behind one call.  There is no CPU fallback: run() raises when the
library is missing or no CUDA device is usable."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NUM_TAGS = 70


class RtConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("threads", C.c_int32), ("pipelines", C.c_int32), ("want_yuv", C.c_int32), ("want_hash", C.c_int32),
                ("want_pip", C.c_int32), ("loops", C.c_int32), ("max_pictures", C.c_int32), ("depth", C.c_int32), ("reserved", C.c_int32 * 7)]


class RtStreamResult(C.Structure):
    _fields_ = [("pictures", C.c_long), ("pictures_output", C.c_long), ("yuv_sha1", C.c_char * 41), ("pip_sha1", C.c_char * 41),
                ("pip_bytes", C.c_size_t), ("error", C.c_int32)]


class RtStats(C.Structure):
    _fields_ = [("streams", C.c_long), ("pictures", C.c_long), ("seconds", C.c_double), ("seconds_total", C.c_double),
                ("input_bytes", C.c_size_t), ("pip_bytes", C.c_size_t), ("submit_wait_parse_s", C.c_double), ("submit_wait_gpu_s", C.c_double),
                ("parse_busy_s", C.c_double), ("parse_wait_s", C.c_double), ("kernel_launches", C.c_ulonglong),
                ("h2d_bytes", C.c_size_t), ("d2h_bytes", C.c_size_t), ("setup_s", C.c_double), ("setup_pictures", C.c_long), ("pinned_alloc_s", C.c_double), ("open_streams_s", C.c_double),
                ("threads", C.c_int32), ("pipelines", C.c_int32)]


PICTURE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_int, C.POINTER(C.c_uint8))


class RuntimeError_(RuntimeError):
    pass


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "_lib", "libopenh264_lh264.so")
    if not os.path.exists(path):
        raise RuntimeError_(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.lh264_rt_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(RtConfig)]
    lib.lh264_rt_destroy.argtypes = [C.c_void_p]
    lib.lh264_rt_add_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    lib.lh264_rt_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lh264_rt_stream_result_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(RtStreamResult)]
    lib.lh264_rt_pip_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.lh264_rt_pip_read.restype = C.c_longlong
    lib.lh264_rt_stats_get.argtypes = [C.c_void_p, C.POINTER(RtStats)]
    lib.lh264_rt_last_error.argtypes = [C.c_void_p]
    lib.lh264_rt_last_error.restype = C.c_char_p
    lib.lh264_rt_trim.argtypes = []
    _LIB = lib
    return lib


def trim():
    """Gives the pinned staging memory kept from finished runs back to the system."""
    load_library().lh264_rt_trim()


class Runtime:
    """rt = Runtime(threads=16); rt.add(bytes) ...; rt.run(); rt.result(i); rt.pip(i)"""

    def __init__(self, device=0, threads=0, pipelines=1, want_yuv=True, want_hash=True, want_pip=True, loops=1, max_pictures=0, depth=0):
        self.lib = load_library()
        cfg = RtConfig(device=device, threads=threads, pipelines=pipelines, want_yuv=int(want_yuv), want_hash=int(want_hash), want_pip=int(want_pip),
                       loops=loops, max_pictures=max_pictures, depth=depth)
        self.h = C.c_void_p()
        rc = self.lib.lh264_rt_create(C.byref(self.h), C.byref(cfg))
        if rc:
            raise RuntimeError_(f"lh264_rt_create: {rc}")
        self._keep = []
        self.n = 0

    def add(self, data: bytes) -> int:
        """One Annex-B stream; the runtime borrows the bytes (kept alive here) and wants the 00 00 00 01 sentinel behind them."""
        buf = C.create_string_buffer(bytes(data) + b"\x00\x00\x00\x01", len(data) + 4)
        self._keep.append(buf)
        i = self.lib.lh264_rt_add_stream(self.h, buf, len(data), 1)
        if i < 0:
            raise RuntimeError_(f"lh264_rt_add_stream: {i}")
        self.n = i + 1
        return i

    def add_shared(self, buf, length) -> int:
        """Another stream over a buffer already added (copies of one input share its bytes)."""
        i = self.lib.lh264_rt_add_stream(self.h, buf, length, 1)
        if i < 0:
            raise RuntimeError_(f"lh264_rt_add_stream: {i}")
        self.n = i + 1
        return i

    def run(self, on_picture=None):
        cb = None
        if on_picture is not None:
            def tramp(_user, stream, index, w, h, ptr):
                n = w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)
                on_picture(stream, index, w, h, C.string_at(ptr, n))
            cb = PICTURE_FN(tramp)
        rc = self.lib.lh264_rt_run(self.h, C.cast(cb, C.c_void_p) if cb else None, None)
        if rc:
            raise RuntimeError_(f"lh264_rt_run: {rc}: {self.lib.lh264_rt_last_error(self.h).decode()}")

    def result(self, i):
        r = RtStreamResult()
        if self.lib.lh264_rt_stream_result_get(self.h, i, C.byref(r)):
            raise RuntimeError_("bad stream index")
        return {"pictures": r.pictures, "pictures_output": r.pictures_output, "yuv_sha1": r.yuv_sha1.decode(), "pip_sha1": r.pip_sha1.decode(),
                "pip_bytes": r.pip_bytes, "error": r.error}

    def pip(self, i):
        """{tag: bytes} of stream i's `.pip` file set; tag 0 = the default stream."""
        out = {}
        for tag in range(NUM_TAGS):
            n = self.lib.lh264_rt_pip_read(self.h, i, tag, None, 0)
            if n < 0:
                raise RuntimeError_(f"lh264_rt_pip_read: {n}")
            if n == 0 and tag != 0:
                continue
            b = C.create_string_buffer(max(int(n), 1))
            self.lib.lh264_rt_pip_read(self.h, i, tag, b, int(n))
            out[tag] = b.raw[:int(n)]
        return out

    def stats(self):
        s = RtStats()
        self.lib.lh264_rt_stats_get(self.h, C.byref(s))
        return {f: getattr(s, f) for f, _ in RtStats._fields_}

    def close(self):
        if self.h:
            self.lib.lh264_rt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
