"""numpy mirrors of the C-ABI records in include/lh264_b200.h and the .lhfb fixture container.

The record layouts replace the reference's SDqLayer per-MB SoA hand-off
(codec/decoder/core/inc/dec_frame.h:61-127).  `.lhfb` files are written by
oracle/ref_dump_hook.cpp (test infrastructure) and by the host decoder's --dump mode; they hold
one record per slice and are merged into pictures here.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

MAX_REFS = 16

MB_NONE, MB_I4x4, MB_I8x8, MB_I16x16, MB_IPCM, MB_P16x16, MB_P16x8, MB_P8x16, MB_P8x8, MB_PSKIP = range(10)
MBF_TRANSFORM8x8 = 1

MB_REC = np.dtype(
    [
        ("mb_type", "u1"),
        ("flags", "u1"),
        ("cbp", "u1"),
        ("qp_y", "u1"),
        ("qp_c", "u1", (2,)),
        ("i16_mode", "u1"),
        ("chroma_mode", "u1"),
        ("slice_idx", "<u2"),
        ("nxn_avail", "u1"),
        ("reserved0", "u1"),
        ("nz_mask", "<u4"),
        ("coded_mask", "<u4"),
        ("coef_off", "<u4"),
        ("ref_idx", "i1", (4,)),
        ("sub_mb_type", "u1", (4,)),
        ("intra_mode", "u1", (16,)),
        ("mv", "<i2", (16, 2)),
        ("reserved1", "u1", (16,)),
    ]
)
assert MB_REC.itemsize == 128

SLICE_REC = np.dtype(
    [
        ("first_mb", "<i4"),
        ("num_mbs", "<i4"),
        ("slice_type", "u1"),
        ("disable_deblocking_idc", "u1"),
        ("alpha_c0_offset", "i1"),
        ("beta_offset", "i1"),
        ("num_ref_idx_l0", "u1"),
        ("weighted_pred", "u1"),
        ("luma_log2_denom", "u1"),
        ("chroma_log2_denom", "u1"),
        ("ref_slot", "<i2", (MAX_REFS,)),
        ("luma_weight", "<i2", (MAX_REFS,)),
        ("luma_offset", "<i2", (MAX_REFS,)),
        ("chroma_weight", "<i2", (MAX_REFS, 2)),
        ("chroma_offset", "<i2", (MAX_REFS, 2)),
        ("scaling_flags", "u1"),
        ("reserved", "u1", (15,)),
        ("scale4", "u1", (6, 16)),
        ("scale8", "u1", (2, 64)),
    ]
)
assert SLICE_REC.itemsize == 16 + 2 * MAX_REFS * 7 + 240
SLF_SCALING, SLF_SCALING_CAVLC = 1, 2


def as_slice_recs(a):
    """Slice records saved before the scaling-list fields existed (240-byte records) widened to the current layout."""
    if a.dtype == SLICE_REC:
        return a
    out = np.zeros(len(a), SLICE_REC)
    for name in a.dtype.names:
        out[name] = a[name]
    return out


@dataclass
class Picture:
    """All slices of one coded picture in hand-off form."""

    mb_w: int
    mb_h: int
    cur_slot: int
    frame_num: int = 0
    poc_lsb: int = 0
    idr_pic_id: int = 0
    nal_ref_idc: int = 0
    serial: int = 0
    slices: np.ndarray = field(default_factory=lambda: np.zeros(0, SLICE_REC))
    mbs: np.ndarray = field(default_factory=lambda: np.zeros(0, MB_REC))
    coefs: np.ndarray = field(default_factory=lambda: np.zeros((0, 16), np.int16))
    scaled: np.ndarray | None = None  # reference-dequantised coefficients (dump only)
    output: bool = True  # False: decoded (a reference for later pictures) but not part of the output sequence
    pix_pre: list | None = None  # per slice: picture before that slice's deblocking
    pix_post: list | None = None  # per slice: picture after it

    @property
    def n_mbs(self) -> int:
        return self.mb_w * self.mb_h

    def final_yuv(self) -> np.ndarray | None:
        return self.pix_post[-1] if self.pix_post else None


def _read(f, fmt):
    size = struct.calcsize(fmt)
    b = f.read(size)
    if len(b) < size:
        raise EOFError
    return struct.unpack(fmt, b)


def read_lhfb(path, max_pictures=None):
    """Parses a .lhfb dump into a list of Picture (slices of one picture merged)."""
    pics: list[Picture] = []
    cur = None
    cur_key = None
    with open(path, "rb") as f:
        while True:
            try:
                magic, ver = _read(f, "<II")
            except EOFError:
                break
            if magic != 0x4246484C or ver != 1:
                raise ValueError("bad lhfb record")
            mb_w, mb_h, slot, frame_num, poc, idr_id, ref_idc, serial = _read(f, "<8i")
            srec = np.frombuffer(f.read(SLICE_REC.itemsize), SLICE_REC, 1).copy()
            (n_mbs,) = _read(f, "<i")
            raw = np.frombuffer(f.read(n_mbs * (4 + MB_REC.itemsize)), np.dtype([("xy", "<i4"), ("rec", MB_REC)]), n_mbs)
            (n_blk,) = _read(f, "<i")
            quant = np.frombuffer(f.read(n_blk * 32), "<i2").reshape(n_blk, 16).copy()
            scaled = np.frombuffer(f.read(n_blk * 32), "<i2").reshape(n_blk, 16).copy()
            (has_pix,) = _read(f, "<i")
            pre = post = None
            if has_pix:
                sz = mb_w * mb_h * 384
                pre = np.frombuffer(f.read(sz), np.uint8).copy()
                post = np.frombuffer(f.read(sz), np.uint8).copy()
            key = (serial, slot, frame_num, poc, idr_id, mb_w, mb_h)
            if cur is None or key != cur_key:
                if max_pictures is not None and len(pics) >= max_pictures:
                    break
                cur = Picture(mb_w, mb_h, slot, frame_num, poc, idr_id, ref_idc & 0xFF, serial)
                cur.output = not (ref_idc & 0x100)      # a concealed picture DecodeFrameNoDelay does not hand out: reconstruct, do not emit
                cur.mbs = np.zeros(mb_w * mb_h, MB_REC)
                cur.scaled = np.zeros((0, 16), np.int16)
                cur.pix_pre, cur.pix_post = [], []
                cur_key = key
                pics.append(cur)
            sidx = len(cur.slices)
            cur.slices = np.concatenate([cur.slices, srec])
            base = len(cur.coefs)
            recs = raw["rec"].copy()
            recs["slice_idx"] = sidx
            recs["coef_off"] += base
            cur.mbs[raw["xy"]] = recs
            cur.coefs = np.concatenate([cur.coefs, quant])
            cur.scaled = np.concatenate([cur.scaled, scaled])
            if has_pix:
                cur.pix_pre.append(pre)
                cur.pix_post.append(post)
    return pics


def write_lhfb(path, pics, with_pix=True):
    """Writes pictures back as one record per slice (used to trim fixtures)."""
    with open(path, "wb") as f:
        for p in pics:
            for s in range(len(p.slices)):
                sl = p.slices[s]
                f.write(struct.pack("<II", 0x4246484C, 1))
                f.write(struct.pack("<8i", p.mb_w, p.mb_h, p.cur_slot, p.frame_num, p.poc_lsb, p.idr_pic_id, p.nal_ref_idc, p.serial))
                f.write(sl.tobytes())
                xy = np.nonzero((p.mbs["slice_idx"] == s) & (np.arange(p.n_mbs) >= sl["first_mb"]) & (np.arange(p.n_mbs) < sl["first_mb"] + sl["num_mbs"]))[0]
                recs = p.mbs[xy].copy()
                nblk = np.array([bin(int(m)).count("1") for m in recs["coded_mask"]], np.int64)
                lo = int(recs["coef_off"][0]) if len(recs) else 0
                hi = int(recs["coef_off"][-1] + nblk[-1]) if len(recs) else 0
                recs["coef_off"] -= lo
                recs["slice_idx"] = 0
                f.write(struct.pack("<i", len(xy)))
                out = np.zeros(len(xy), np.dtype([("xy", "<i4"), ("rec", MB_REC)]))
                out["xy"] = xy
                out["rec"] = recs
                f.write(out.tobytes())
                f.write(struct.pack("<i", hi - lo))
                f.write(p.coefs[lo:hi].tobytes())
                sc = p.scaled[lo:hi] if p.scaled is not None and len(p.scaled) >= hi else np.zeros((hi - lo, 16), np.int16)
                f.write(sc.tobytes())
                if with_pix and p.pix_post:
                    f.write(struct.pack("<i", 1))
                    f.write(p.pix_pre[s].tobytes())
                    f.write(p.pix_post[s].tobytes())
                else:
                    f.write(struct.pack("<i", 0))


def save_fixture(path, pics):
    """Compact .npz form of a picture sequence (hand-off records + the reference's final pictures)."""
    meta = np.array([[p.mb_w, p.mb_h, p.cur_slot, p.frame_num, p.poc_lsb, p.idr_pic_id, p.nal_ref_idc, p.serial, len(p.slices), len(p.coefs)] for p in pics], np.int32)
    np.savez_compressed(
        path,
        meta=meta,
        slices=np.concatenate([p.slices for p in pics]),
        mbs=np.concatenate([p.mbs for p in pics]),
        coefs=np.concatenate([p.coefs for p in pics]) if sum(len(p.coefs) for p in pics) else np.zeros((0, 16), np.int16),
        yuv=np.concatenate([p.final_yuv() for p in pics]),
    )


def load_fixture(path):
    z = np.load(path)
    pics = []
    o_s = o_m = o_c = o_y = 0
    for row in z["meta"]:
        mb_w, mb_h, slot, fn, poc, idr, ref_idc, serial, n_sl, n_co = (int(v) for v in row)
        p = Picture(mb_w, mb_h, slot, fn, poc, idr, ref_idc, serial)
        p.slices = as_slice_recs(z["slices"][o_s : o_s + n_sl]).copy()
        p.mbs = z["mbs"][o_m : o_m + mb_w * mb_h].copy()
        p.coefs = z["coefs"][o_c : o_c + n_co].copy()
        ysz = mb_w * mb_h * 384
        p.pix_post = [z["yuv"][o_y : o_y + ysz].copy()]
        o_s += n_sl
        o_m += mb_w * mb_h
        o_c += n_co
        o_y += ysz
        pics.append(p)
    return pics


def n_slots_needed(pics):
    return max(max(p.cur_slot for p in pics), max(int(p.slices["ref_slot"].max()) for p in pics)) + 1


# ---- recompression-model hand-off (lh264_model_mb / lh264_model_slice, include/lh264_b200.h) -----------
MM_EMIT_SKIP_RUN, MM_STOP_IN_SKIP, MM_WRITE_BLOCK, MM_STOP_BIT, MM_T8_CODED, MM_T8 = 1, 2, 4, 8, 16, 32
NUM_TAGS = 70

MODEL_MB = np.dtype(
    [
        ("mb_addr", "<u4"),
        ("flags", "<u2"),
        ("skip_run", "<u2"),
        ("mb_in_slice", "<u2"),
        ("type_code", "u1"),
        ("cbp", "u1"),
        ("luma_qp", "u1"),
        ("chroma_mode", "u1"),
        ("i16_mode", "u1"),
        ("qp_delta", "i1"),
        ("sub_mb_type", "u1", (4,)),
        ("ref_idx", "i1", (4,)),
        ("pred_mode", "u1", (16,)),
        ("pred_avail", "u1", (16,)),
        ("coded_mask", "<u4"),
        ("coef_off", "<u4"),
        ("mvd", "<i2", (16, 2)),
    ]
)
assert MODEL_MB.itemsize == 128

MODEL_SLICE = np.dtype(
    [
        ("first_rec", "<i4"),
        ("n_recs", "<i4"),
        ("frame_num", "<i4"),
        ("slice_type", "u1"),
        ("num_ref_idx_l0", "u1"),
        ("n_pad_bits", "u1"),
        ("pad_bits", "u1"),
        ("chroma_format_idc", "u1"),
        ("reserved", "u1", (3,)),
    ]
)
assert MODEL_SLICE.itemsize == 20


@dataclass
class ModelPicture:
    """The recompression-model records of one coded picture (all slices, slice-data order)."""

    mb_w: int
    mb_h: int
    slices: np.ndarray
    recs: np.ndarray
    coefs: np.ndarray  # (n_blocks, 16) int16 quantised levels, the picture's coefficient pool


def read_lhmd(path, max_pictures=None):
    """Parses the 'LHMD' dump written by the host decoder (LhfbSink::submit_model)."""
    pics = []
    with open(path, "rb") as f:
        while max_pictures is None or len(pics) < max_pictures:
            try:
                magic, mb_w, mb_h, n_sl, n_rec, n_blk = _read(f, "<6i")
            except EOFError:
                break
            if magic != 0x444D484C:
                raise ValueError("bad lhmd record")
            sl = np.frombuffer(f.read(n_sl * MODEL_SLICE.itemsize), MODEL_SLICE, n_sl).copy()
            rc = np.frombuffer(f.read(n_rec * MODEL_MB.itemsize), MODEL_MB, n_rec).copy()
            co = np.frombuffer(f.read(n_blk * 32), "<i2").reshape(n_blk, 16).copy()
            pics.append(ModelPicture(mb_w, mb_h, sl, rc, co))
    return pics
