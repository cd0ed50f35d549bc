"""ctypes host layer over the C ABI (include/lh264_b200.h).

Mirrors the reference's decoder-side operator interface for the hot path: a `Context` owns a
CUDA stream, `open_stream` allocates the picture pool of one decoder instance (the reference's
PPicture pool, pic_queue.cpp:60-107), `recon` runs what WelsTargetSliceConstruction
(decode_slice.cpp:110-206) does for every picture of a batch.  No CPU fallback: if the CUDA
library or a device is missing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import library_path
from .records import MB_REC, MODEL_MB, MODEL_SLICE, NUM_TAGS, SLICE_REC, Picture

_LIB = None

EXPORTS = [
    "lh264_abi_version", "lh264_last_error", "lh264_ctx_create", "lh264_ctx_destroy",
    "lh264_stream_open", "lh264_stream_close", "lh264_recon_submit", "lh264_batch_upload",
    "lh264_batch_recon", "lh264_batch_free", "lh264_kernel_launches", "lh264_kernel_ms",
    "lh264_picture_read", "lh264_picture_write", "lh264_picture_device_ptrs", "lh264_sync", "lh264_join", "lh264_inputs_consumed",
    "lh264_pictures_read_packed", "lh264_output_wait", "lh264_host_alloc", "lh264_host_alloc_wc", "lh264_host_free", "lh264_thread_bind",
    "lh264_k_idct4x4_add", "lh264_k_idct8x8_add", "lh264_k_mc_luma", "lh264_k_mc_chroma",
    "lh264_k_expand_plane",
    "lh264_model_open", "lh264_model_close", "lh264_model_submit", "lh264_model_finish", "lh264_model_read",
    "lh264_model_decisions", "lh264_model_batch_upload", "lh264_model_batch_run", "lh264_model_batch_free",
    "lh264_model_decode", "lh264_submit_compact", "lh264_read_ticket", "lh264_output_wait_ticket",
]


class DecodeSlice(C.Structure):       # lh264_decode_slice
    _fields_ = [("first_mb", C.c_uint32), ("frame_num", C.c_int32), ("mb_w", C.c_uint16), ("mb_h", C.c_uint16), ("slice_type", C.c_uint8),
                ("num_ref_idx_l0", C.c_uint8), ("slice_qp", C.c_uint8), ("flags", C.c_uint8), ("header_bits", C.c_uint32)]


class DecodeSliceOut(C.Structure):    # lh264_decode_slice_out
    _fields_ = [("byte_offset", C.c_uint64), ("n_bits", C.c_uint32), ("n_mbs", C.c_uint32), ("n_pad_bits", C.c_uint8), ("pad_bits", C.c_uint8),
                ("reserved", C.c_uint16), ("reserved2", C.c_uint32)]


class DecodeJob(C.Structure):         # lh264_decode_job
    _fields_ = [("max_mbs", C.c_int32), ("with_8x8", C.c_int32), ("tag_data", C.c_void_p * NUM_TAGS), ("tag_len", C.c_uint32 * NUM_TAGS),
                ("slices", C.c_void_p), ("n_slices", C.c_int32), ("error", C.c_int32), ("out", C.c_void_p), ("out_capacity", C.c_uint64),
                ("slice_out", C.c_void_p)]


class PictureDesc(C.Structure):
    _fields_ = [
        ("stream", C.c_int32),
        ("cur_slot", C.c_int32),
        ("n_slices", C.c_int32),
        ("n_coef_blocks", C.c_int32),
        ("slices", C.c_void_p),
        ("mbs", C.c_void_p),
        ("coefs", C.c_void_p),
    ]


class ModelPictureDesc(C.Structure):
    _fields_ = [
        ("model_stream", C.c_int32),
        ("mb_w", C.c_int32),
        ("mb_h", C.c_int32),
        ("n_slices", C.c_int32),
        ("n_recs", C.c_int32),
        ("n_coef_blocks", C.c_int32),
        ("slices", C.c_void_p),
        ("recs", C.c_void_p),
        ("coefs", C.c_void_p),
    ]


class LibraryMissing(RuntimeError):
    pass


def load_library():
    """Loads liblh264_b200.so (built in-tree by losslessh264_b200.build).  Never falls back."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("LH264_B200_LIB") or library_path()      # (the override loads a kernel experiment build: development only)
    if not os.path.exists(path):
        raise LibraryMissing(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.lh264_last_error.restype = C.c_char_p
    lib.lh264_kernel_launches.restype = C.c_uint64
    lib.lh264_kernel_ms.restype = C.c_float
    lib.lh264_host_alloc.restype = C.c_void_p
    lib.lh264_host_alloc.argtypes = [C.c_size_t]
    lib.lh264_host_free.argtypes = [C.c_void_p]
    lib.lh264_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    lib.lh264_ctx_destroy.argtypes = [C.c_void_p]
    lib.lh264_stream_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.lh264_stream_close.argtypes = [C.c_void_p, C.c_int]
    lib.lh264_recon_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.lh264_batch_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.lh264_batch_recon.argtypes = [C.c_void_p, C.c_void_p]
    lib.lh264_batch_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.lh264_kernel_launches.argtypes = [C.c_void_p]
    lib.lh264_kernel_ms.argtypes = [C.c_void_p, C.c_int]
    lib.lh264_sync.argtypes = [C.c_void_p]
    lib.lh264_model_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.lh264_join.argtypes = [C.c_void_p]
    lib.lh264_inputs_consumed.argtypes = [C.c_void_p]
    lib.lh264_output_wait.argtypes = [C.c_void_p]
    lib.lh264_picture_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.lh264_picture_write.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.lh264_pictures_read_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
    lib.lh264_k_idct4x4_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.lh264_k_idct8x8_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.lh264_k_mc_luma.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.lh264_k_mc_chroma.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.lh264_k_expand_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.lh264_model_open.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.lh264_model_close.argtypes = [C.c_void_p, C.c_int]
    lib.lh264_model_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.lh264_model_finish.argtypes = [C.c_void_p, C.c_int]
    lib.lh264_model_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.lh264_model_read.restype = C.c_longlong
    lib.lh264_model_batch_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.lh264_model_batch_run.argtypes = [C.c_void_p, C.c_void_p]
    lib.lh264_model_batch_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.lh264_model_decisions.argtypes = [C.c_void_p, C.c_int]
    lib.lh264_model_decisions.restype = C.c_ulonglong
    _LIB = lib
    return lib


class Lh264Error(RuntimeError):
    pass


def _check(lib, rc):
    if rc < 0:
        raise Lh264Error(f"lh264 error {rc}: {lib.lh264_last_error().decode()}")
    return rc


class PinnedArray:
    """numpy view over pinned host memory from lh264_host_alloc."""

    def __init__(self, nbytes):
        self.lib = load_library()
        self.ptr = self.lib.lh264_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("lh264_host_alloc failed")
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.lh264_host_free(self.ptr)
            self.ptr = None


class Context:
    def __init__(self, device=0, cuda_stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib, self.lib.lh264_ctx_create(C.byref(h), device, cuda_stream))
        self.h = h
        self._keep = []

    def close(self):
        if self.h:
            self.lib.lh264_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- streams / pictures -------------------------------------------------------------------
    def open_stream(self, mb_w, mb_h, n_slots):
        return _check(self.lib, self.lib.lh264_stream_open(self.h, mb_w, mb_h, n_slots))

    def close_stream(self, s):
        _check(self.lib, self.lib.lh264_stream_close(self.h, s))

    def _descs(self, pics, streams):
        arr = (PictureDesc * len(pics))()
        keep = []
        for i, (p, s) in enumerate(zip(pics, streams)):
            slices = np.ascontiguousarray(p.slices, SLICE_REC)
            mbs = np.ascontiguousarray(p.mbs, MB_REC)
            coefs = np.ascontiguousarray(p.coefs, np.int16)
            keep += [slices, mbs, coefs]
            arr[i] = PictureDesc(s, p.cur_slot, len(slices), len(coefs), slices.ctypes.data, mbs.ctypes.data, coefs.ctypes.data if len(coefs) else None)
        return arr, keep

    def recon(self, pics, streams, sync=True):
        """Reconstructs `pics` (records.Picture) into their streams' cur_slot.  Host buffers in."""
        arr, keep = self._descs(pics, streams)
        self._keep = keep
        _check(self.lib, self.lib.lh264_recon_submit(self.h, arr, len(pics)))
        if sync:
            self.sync()

    def recon_descs(self, pics, streams):
        """The lh264_picture_desc array of a submit, built once (plain structs of pointers into the caller's host buffers - what a
        C host has at hand anyway); pass it to recon_prepared() as often as needed.  The buffers must stay alive and unchanged."""
        arr, keep = self._descs(pics, streams)
        return (arr, len(pics), keep)

    def recon_prepared(self, prepared, sync=True):
        arr, n, _ = prepared
        _check(self.lib, self.lib.lh264_recon_submit(self.h, arr, n))
        if sync:
            self.sync()

    def upload_batch(self, pics, streams):
        arr, keep = self._descs(pics, streams)
        h = C.c_void_p()
        _check(self.lib, self.lib.lh264_batch_upload(self.h, arr, len(pics), C.byref(h)))
        return h

    def recon_resident(self, batch):
        _check(self.lib, self.lib.lh264_batch_recon(self.h, batch))

    def free_batch(self, batch):
        self.lib.lh264_batch_free(self.h, batch)

    def sync(self):
        _check(self.lib, self.lib.lh264_sync(self.h))
        self._keep = []

    def join(self):
        """Orders the context's CUDA stream after the model work submitted so far."""
        _check(self.lib, self.lib.lh264_join(self.h))

    def read_picture(self, stream, slot, mb_w, mb_h, crop=None):
        """Returns the (cropped) I420 picture as a flat uint8 array Y|U|V."""
        x, y, w, h = crop if crop else (0, 0, mb_w * 16, mb_h * 16)
        cw, ch = (w + 1) // 2, (h + 1) // 2
        out = np.empty(w * h + 2 * cw * ch, np.uint8)
        base = out.ctypes.data
        _check(self.lib, self.lib.lh264_picture_read(self.h, stream, slot, x, y, w, h, base, w, base + w * h, base + w * h + cw * ch, cw, 1))
        return out

    def write_picture(self, stream, slot, mb_w, mb_h, yuv):
        yuv = np.ascontiguousarray(yuv, np.uint8)
        w, h = mb_w * 16, mb_h * 16
        base = yuv.ctypes.data
        _check(self.lib, self.lib.lh264_picture_write(self.h, stream, slot, base, w, base + w * h, base + w * h + w * h // 4, w // 2))

    def read_packed(self, streams, slots, crop, dst: "PinnedArray", sync=True):
        s = np.ascontiguousarray(streams, np.int32)
        sl = np.ascontiguousarray(slots, np.int32)
        x, y, w, h = crop
        _check(self.lib, self.lib.lh264_pictures_read_packed(self.h, s.ctypes.data, sl.ctypes.data, len(s), x, y, w, h, dst.ptr, dst.nbytes, 1 if sync else 0))

    def output_wait(self):
        """Waits for the device-to-host copies started by read_packed(sync=False)."""
        _check(self.lib, self.lib.lh264_output_wait(self.h))

    def kernel_launches(self):
        return int(self.lib.lh264_kernel_launches(self.h))

    def kernel_ms(self, family):
        return float(self.lib.lh264_kernel_ms(self.h, family))

    # ---- recompression model (compress direction) -------------------------------------------------------
    def model_open(self, max_mbs, with_8x8=False):
        """Model state of one stream in HBM (the reference's MacroblockModel + FreqImage + tagged bool coders)."""
        return _check(self.lib, self.lib.lh264_model_open(self.h, max_mbs, 1 if with_8x8 else 0))

    def model_close(self, ms):
        _check(self.lib, self.lib.lh264_model_close(self.h, ms))

    def _model_descs(self, mpics, model_streams):
        arr = (ModelPictureDesc * len(mpics))()
        keep = []
        for i, (p, ms) in enumerate(zip(mpics, model_streams)):
            sl = np.ascontiguousarray(p.slices, MODEL_SLICE)
            rc = np.ascontiguousarray(p.recs, MODEL_MB)
            co = np.ascontiguousarray(p.coefs, np.int16)
            keep += [sl, rc, co]
            arr[i] = ModelPictureDesc(ms, p.mb_w, p.mb_h, len(sl), len(rc), len(co), sl.ctypes.data, rc.ctypes.data, co.ctypes.data if len(co) else None)
        return arr, keep

    def model_submit(self, mpics, model_streams, sync=True):
        """Codes one picture (records.ModelPicture) per stream; what WelsDecodeSlice's MB loop does in the reference."""
        arr, keep = self._model_descs(mpics, model_streams)
        self._keep_model = keep
        _check(self.lib, self.lib.lh264_model_submit(self.h, arr, len(mpics)))
        if sync:
            self.sync()

    def model_descs(self, mpics, model_streams):
        """The lh264_model_picture array of a submit, built once; see recon_descs()."""
        arr, keep = self._model_descs(mpics, model_streams)
        return (arr, len(mpics), keep)

    def model_submit_prepared(self, prepared, sync=True):
        arr, n, _ = prepared
        _check(self.lib, self.lib.lh264_model_submit(self.h, arr, n))
        if sync:
            self.sync()

    def model_upload_batch(self, mpics, model_streams):
        arr, keep = self._model_descs(mpics, model_streams)
        h = C.c_void_p()
        _check(self.lib, self.lib.lh264_model_batch_upload(self.h, arr, len(mpics), C.byref(h)))
        return h

    def model_run_resident(self, batch):
        _check(self.lib, self.lib.lh264_model_batch_run(self.h, batch))

    def model_free_batch(self, batch):
        self.lib.lh264_model_batch_free(self.h, batch)

    def model_finish(self, ms):
        """Ends all tag streams; returns {tag id: bytes} = the contents of the `.pip.<tag>` side files."""
        _check(self.lib, self.lib.lh264_model_finish(self.h, ms))
        out = {}
        for tag in range(NUM_TAGS):
            n = _check(self.lib, self.lib.lh264_model_read(self.h, ms, tag, None, 0, 0))
            if n > 0:
                buf = np.empty(n, np.uint8)
                _check(self.lib, self.lib.lh264_model_read(self.h, ms, tag, buf.ctypes.data, 0, n))
                out[tag] = buf.tobytes()
        return out

    def model_decisions(self, ms):
        return int(self.lib.lh264_model_decisions(self.h, ms))

    def model_decode(self, streams, out_capacity=None):
        """Decompress direction: one job per stream, all jobs concurrently on the GPU (lh264_model_decode).
        streams: list of dicts {tags: {tag id: bytes}, slices: ctypes array of DecodeSlice (or pointer + n_slices), max_mbs, with_8x8}.
        Returns a list of (bits: np.uint8 array, outs: ctypes array of DecodeSliceOut) — or raises; a job that needs a larger
        output buffer is retried with twice the capacity."""
        n = len(streams)
        caps = [int(out_capacity or (4 * sum(len(b) for b in s["tags"].values()) + (1 << 20))) for s in streams]
        while True:
            jobs = (DecodeJob * n)()
            keep = {}
            bits, outs = [], []
            for j, s in enumerate(streams):
                job = jobs[j]
                job.max_mbs = int(s["max_mbs"]); job.with_8x8 = int(bool(s.get("with_8x8", True)))
                for t, b in s["tags"].items():
                    buf = keep.get(id(b))                 # jobs that share a bytes object share the host copy
                    if buf is None:
                        buf = keep[id(b)] = np.frombuffer(bytes(b) + b"\0", np.uint8)
                    job.tag_data[int(t)] = buf.ctypes.data; job.tag_len[int(t)] = len(b)
                ns = int(s["n_slices"])
                job.slices = C.cast(s["slices"], C.c_void_p).value if ns else None
                job.n_slices = ns
                o = np.zeros(caps[j], np.uint8)
                so = (DecodeSliceOut * max(ns, 1))()
                bits.append(o); outs.append(so)
                job.out = o.ctypes.data; job.out_capacity = caps[j]; job.slice_out = C.cast(so, C.c_void_p).value
            rc = self.lib.lh264_model_decode(self.h, jobs, n)
            small = [j for j in range(n) if jobs[j].error == 22]
            if rc != 0 and small and all(jobs[j].error in (0, 22) for j in range(n)):
                for j in small:
                    caps[j] *= 2
                continue
            _check(self.lib, rc)
            return list(zip(bits, outs))

    # ---- kernel-level entry points ---------------------------------------------------------------
    def k_idct4x4_add(self, pix, coef):
        pix = np.ascontiguousarray(pix, np.uint8).copy()
        coef = np.ascontiguousarray(coef, np.int16)
        _check(self.lib, self.lib.lh264_k_idct4x4_add(self.h, pix.ctypes.data, coef.ctypes.data, len(pix)))
        return pix

    def k_idct8x8_add(self, pix, coef):
        pix = np.ascontiguousarray(pix, np.uint8).copy()
        coef = np.ascontiguousarray(coef, np.int16)
        _check(self.lib, self.lib.lh264_k_idct8x8_add(self.h, pix.ctypes.data, coef.ctypes.data, len(pix)))
        return pix

    def k_mc(self, ref, blocks, luma=True):
        ref = np.ascontiguousarray(ref, np.uint8)
        blocks = np.ascontiguousarray(blocks, np.int32)
        out = np.zeros((len(blocks), 16, 16), np.uint8)
        f = self.lib.lh264_k_mc_luma if luma else self.lib.lh264_k_mc_chroma
        _check(self.lib, f(self.h, ref.ctypes.data, ref.shape[1], ref.shape[0], blocks.ctypes.data, len(blocks), out.ctypes.data))
        return out

    def k_expand_plane(self, plane, pad):
        plane = np.ascontiguousarray(plane, np.uint8)
        h, w = plane.shape
        out = np.zeros((h + 2 * pad, w + 2 * pad), np.uint8)
        _check(self.lib, self.lib.lh264_k_expand_plane(self.h, plane.ctypes.data, w, h, pad, out.ctypes.data))
        return out
