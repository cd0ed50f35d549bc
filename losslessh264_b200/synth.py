"""Seeded synthetic hand-off records: valid I/P pictures of any size without needing an encoder.

Used by the parity tests (CUDA vs oracle at 1080p) and by bench.py for the kernel-only roofline
leg.  The records obey the contract of include/lh264_b200.h: resolved intra modes only name
neighbours that exist, partition MVs are replicated over their 4x4 blocks, coded_mask / cbp /
nz_mask agree with the coefficient pool.  Content statistics loosely follow a QP-26 Constrained
Baseline IPPP stream (BASELINE.json config 4): mostly P16x16/skip, sparse low-frequency levels.
"""
from __future__ import annotations

import numpy as np

from .records import (
    MB_I4x4, MB_I8x8, MB_I16x16, MB_P16x16, MB_P16x8, MB_P8x16, MB_P8x8, MB_PSKIP,
    MB_REC, MBF_TRANSFORM8x8, SLICE_REC, Picture,
)

# z (decode order) -> (bx, by) of the 4x4 block
_ZX = np.array([(z & 1) | ((z & 4) >> 1) for z in range(16)])
_ZY = np.array([((z & 2) >> 1) | ((z & 8) >> 2) for z in range(16)])
_Z_OF = np.zeros((4, 4), np.int64)  # [by][bx] -> z
for _z in range(16):
    _Z_OF[_ZY[_z], _ZX[_z]] = _z

_QPC = np.array(list(range(30)) + [29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39])


def _choose(rng, table, n):
    """table: list of allowed values; returns n random picks."""
    t = np.asarray(table)
    return t[rng.integers(0, len(t), n)]


def _intra_modes_4x4(rng, left, top, tl, tr_mb_top, tr_mb_topright):
    """Resolved 4x4 modes for n MBs.  left/top/tl: MB-level availability (bool arrays)."""
    n = len(left)
    modes = np.zeros((n, 16), np.uint8)
    for z in range(16):
        bx, by = int(_ZX[z]), int(_ZY[z])
        a_left = left if bx == 0 else np.ones(n, bool)
        a_top = top if by == 0 else np.ones(n, bool)
        if bx == 0 and by == 0:
            a_tl = tl
        elif bx == 0:
            a_tl = left
        elif by == 0:
            a_tl = top
        else:
            a_tl = np.ones(n, bool)
        if by == 0:
            a_tr = tr_mb_top if bx < 3 else tr_mb_topright
        elif bx == 3:
            a_tr = np.zeros(n, bool)
        else:
            a_tr = np.full(n, _Z_OF[by - 1, bx + 1] < z)
        both = a_left & a_top
        m = np.full(n, 11, np.uint8)  # DC_128
        pick = rng.integers(0, 9, n)
        # both available: any of the nine modes (diagonal-right family needs top-left too)
        mb = pick.copy()
        need_tl = (mb == 4) | (mb == 5) | (mb == 6)
        mb = np.where(need_tl & ~a_tl, 2, mb)
        mb = np.where((mb == 3) & ~a_tr, 12, mb)
        mb = np.where((mb == 7) & ~a_tr, 13, mb)
        m = np.where(both, mb, m)
        ml = _choose(rng, [1, 8, 9], n)
        m = np.where(a_left & ~a_top, ml, m)
        mt = _choose(rng, [0, 3, 7, 10], n)
        mt = np.where((mt == 3) & ~a_tr, 12, mt)
        mt = np.where((mt == 7) & ~a_tr, 13, mt)
        m = np.where(a_top & ~a_left, mt, m)
        modes[:, z] = m
    return modes


def _intra_modes_8x8(rng, left, top, tl, topright):
    n = len(left)
    modes = np.zeros((n, 16), np.uint8)
    ones, zeros = np.ones(n, bool), np.zeros(n, bool)
    for q in range(4):
        a_left = left if (q & 1) == 0 else ones
        a_top = top if (q >> 1) == 0 else ones
        a_tl = [tl, top, left, ones][q]
        a_tr = [top, topright, ones, zeros][q]
        both = a_left & a_top
        m = np.full(n, 11, np.uint8)
        mb = rng.integers(0, 9, n)
        need_tl = (mb == 4) | (mb == 5) | (mb == 6)
        mb = np.where(need_tl & ~a_tl, 2, mb)
        mb = np.where((mb == 3) & ~a_tr, 12, mb)
        mb = np.where((mb == 7) & ~a_tr, 13, mb)
        m = np.where(both, mb, m)
        m = np.where(a_left & ~a_top, _choose(rng, [1, 8, 9], n), m)
        mt = _choose(rng, [0, 3, 7, 10], n)
        mt = np.where((mt == 3) & ~a_tr, 12, mt)
        mt = np.where((mt == 7) & ~a_tr, 13, mt)
        m = np.where(a_top & ~a_left, mt, m)
        modes[:, 4 * q] = m
    return modes


def _levels(rng, n_blocks, density, amp):
    """Sparse low-frequency quantised levels for n_blocks 4x4 blocks (raster inside the block)."""
    lev = np.zeros((n_blocks, 16), np.int16)
    # probability falls off with frequency index (x + y)
    freq = np.array([(i & 3) + (i >> 2) for i in range(16)])
    p = density * np.exp(-0.9 * freq)
    on = rng.random((n_blocks, 16)) < p
    mag = rng.geometric(1.0 / amp, (n_blocks, 16)).astype(np.int16)
    sign = rng.integers(0, 2, (n_blocks, 16)).astype(np.int16) * 2 - 1
    lev[on] = (mag * sign)[on]
    return lev


def synthetic_picture(rng, mb_w, mb_h, cur_slot, ref_slots, intra_only, qp=26, enable_8x8=True, two_slices=False, big_mv_rate=0.01):
    n = mb_w * mb_h
    mbx = np.arange(n) % mb_w
    mby = np.arange(n) // mb_w
    mbs = np.zeros(n, MB_REC)

    n_sl = 2 if two_slices and mb_h >= 2 else 1
    slices = np.zeros(n_sl, SLICE_REC)
    split = (mb_h // 2) * mb_w if n_sl == 2 else n
    slices["first_mb"] = [0, split][:n_sl]
    slices["num_mbs"] = [split, n - split][:n_sl]
    slices["slice_type"] = 2 if intra_only else 0
    slices["num_ref_idx_l0"] = max(1, len(ref_slots))
    for i in range(16):
        slices["ref_slot"][:, i] = ref_slots[min(i, len(ref_slots) - 1)] if ref_slots else 0
    slice_idx = (np.arange(n) >= split).astype(np.uint16)
    mbs["slice_idx"] = slice_idx
    first_row_of_slice = (mby == 0) | ((n_sl == 2) & (np.arange(n) >= split) & (np.arange(n) < split + mb_w))

    # ---- types ----
    if intra_only:
        r = rng.random(n)
        t = np.where(r < 0.55, MB_I4x4, np.where(r < 0.92 if enable_8x8 else r < 2, MB_I16x16, MB_I8x8))
    else:
        r = rng.random(n)
        edges = np.cumsum([0.30, 0.36, 0.08, 0.08, 0.11, 0.035, 0.025, 0.01])
        kinds = [MB_PSKIP, MB_P16x16, MB_P16x8, MB_P8x16, MB_P8x8, MB_I4x4, MB_I16x16, MB_I8x8 if enable_8x8 else MB_I16x16]
        t = np.full(n, MB_P16x16)
        for e, k in zip(edges[::-1], kinds[::-1]):
            t = np.where(r < e, k, t)
    mbs["mb_type"] = t
    intra = (t == MB_I4x4) | (t == MB_I8x8) | (t == MB_I16x16)
    inter = ~intra

    # ---- QP ----
    qpy = np.clip(qp + rng.integers(-2, 3, n) * (rng.random(n) < 0.1), 0, 51)
    mbs["qp_y"] = qpy
    mbs["qp_c"][:, 0] = _QPC[qpy]
    mbs["qp_c"][:, 1] = _QPC[qpy]

    # ---- intra modes (neighbours in another slice are unavailable) ----
    left = mbx > 0
    top = ~first_row_of_slice
    tl = left & top & ~((n_sl == 2) & (np.arange(n) == split + 0))
    if n_sl == 2:  # the MB above-left of the first MBs of slice 1 belongs to slice 0
        tl = tl & (slice_idx[np.maximum(np.arange(n) - mb_w - 1, 0)] == slice_idx)
    topright = top & (mbx < mb_w - 1)
    if n_sl == 2:
        topright = topright & (slice_idx[np.maximum(np.arange(n) - mb_w + 1, 0)] == slice_idx)
    m16 = np.full(n, 6, np.uint8)
    m16 = np.where(left & top, rng.integers(0, 4, n), m16)
    m16 = np.where((m16 == 3) & ~tl, 2, m16)
    m16 = np.where(left & ~top, _choose(rng, [1, 4], n), m16)
    m16 = np.where(top & ~left, _choose(rng, [0, 5], n), m16)
    mbs["i16_mode"] = m16
    mc = np.full(n, 6, np.uint8)
    mc = np.where(left & top, rng.integers(0, 4, n), mc)
    mc = np.where((mc == 3) & ~tl, 0, mc)
    mc = np.where(left & ~top, _choose(rng, [1, 4], n), mc)
    mc = np.where(top & ~left, _choose(rng, [2, 5], n), mc)
    mbs["chroma_mode"] = mc
    m4 = _intra_modes_4x4(rng, left, top, tl, top, topright)
    m8 = _intra_modes_8x8(rng, left, top, tl, topright)
    mbs["intra_mode"] = np.where((t == MB_I8x8)[:, None], m8, m4)
    mbs["nxn_avail"] = top * 1 + tl * 2 + left * 4 + topright * 8
    t8 = t == MB_I8x8
    if enable_8x8 and not intra_only:
        t8 = t8 | (((t == MB_P16x16) | (t == MB_P16x8) | (t == MB_P8x16)) & (rng.random(n) < 0.15))
    mbs["flags"] = np.where(t8, MBF_TRANSFORM8x8, 0)

    # ---- motion ----
    if not intra_only:
        gx, gy = rng.integers(-12, 13), rng.integers(-8, 9)
        base = np.stack([gx + rng.integers(-6, 7, n), gy + rng.integers(-6, 7, n)], 1)
        big = rng.random(n) < big_mv_rate
        base = np.where(big[:, None], rng.integers(-600, 600, (n, 2)), base)
        base = np.where((t == MB_PSKIP)[:, None] & (rng.random(n) < 0.5)[:, None], 0, base)
        mv = np.repeat(base[:, None, :], 16, 1).astype(np.int16)
        jit = rng.integers(-5, 6, (n, 16, 2)).astype(np.int16)
        r_of = np.arange(16)
        bx, by = r_of & 3, r_of >> 2
        # partition id per 4x4 block for each type
        pid_16x8 = by >> 1
        pid_8x16 = bx >> 1
        quad = (by >> 1) * 2 + (bx >> 1)
        sub = rng.integers(0, 4, (n, 4)).astype(np.uint8)
        mbs["sub_mb_type"] = np.where((t == MB_P8x8)[:, None], sub, 0)
        sub_blk = sub[:, quad]  # (n,16)
        pid_8x8 = np.where(sub_blk == 0, quad * 4,
                  np.where(sub_blk == 1, quad * 4 + (by & 1),
                  np.where(sub_blk == 2, quad * 4 + (bx & 1), quad * 4 + (by & 1) * 2 + (bx & 1))))
        pid = np.where((t == MB_P16x8)[:, None], pid_16x8[None, :],
              np.where((t == MB_P8x16)[:, None], pid_8x16[None, :],
              np.where((t == MB_P8x8)[:, None], pid_8x8, 0)))
        # jitter indexed by partition id so blocks of one partition share a vector
        pj = np.take_along_axis(jit, np.repeat(pid[:, :, None], 2, 2).astype(np.int64), 1)
        mv = mv + np.where(inter[:, None, None], pj, 0)
        mbs["mv"] = np.where(inter[:, None, None], mv, 0)
        nref = max(1, len(ref_slots))
        ri = rng.integers(0, nref, (n, 4)).astype(np.int8)
        # one ref per 8x8 quadrant; 16x16/skip share quadrant 0's, 16x8 per row pair, 8x16 per column pair
        ri = np.where(((t == MB_P16x16) | (t == MB_PSKIP))[:, None], ri[:, :1], ri)
        ri = np.where((t == MB_P16x8)[:, None], ri[:, [0, 0, 2, 2]], ri)
        ri = np.where((t == MB_P8x16)[:, None], ri[:, [0, 1, 0, 1]], ri)
        ri = np.where((t == MB_PSKIP)[:, None], 0, ri)
        mbs["ref_idx"] = np.where(inter[:, None], ri, -1)
    else:
        mbs["ref_idx"] = -1

    # ---- coefficients ----
    density = np.where(intra, 0.9, np.where(t == MB_PSKIP, 0.0, 0.35))
    mb_has = rng.random(n) < np.where(intra, 0.95, np.where(t == MB_PSKIP, 0.0, 0.6))
    lev = _levels(rng, n * 24, 1.0, 2.0).reshape(n, 24, 16)
    blk_on = rng.random((n, 24)) < density[:, None]
    lev = lev * (blk_on & mb_has[:, None])[:, :, None].astype(np.int16)
    # 8x8 transform: chroma stays 4x4; luma quadrant = 64 coefficients in raster 8x8, make it sparse too
    lev[:, 16:, :] = lev[:, 16:, :] * (rng.random((n, 8, 1)) < 0.5)
    i16 = t == MB_I16x16
    # I16x16: DC level in slot 0 of every block; AC either all-off or free
    ac_off = i16 & (rng.random(n) < 0.5)
    lev[ac_off, :16, 1:] = 0
    lev[i16, :16, 0] = (rng.integers(-6, 7, (int(i16.sum()), 16)) * (rng.random((int(i16.sum()), 16)) < 0.5)).astype(np.int16)
    any_blk = (lev != 0).any(2)  # (n,24)
    luma_any = any_blk[:, :16].copy()
    if t8.any():
        q_any = luma_any.reshape(n, 4, 4).any(2)
        luma_any = np.where(t8[:, None], np.repeat(q_any, 4, 1), luma_any)
    coded = np.concatenate([luma_any, any_blk[:, 16:]], 1)
    weights = (1 << np.arange(24)).astype(np.uint32)
    mbs["coded_mask"] = (coded * weights).sum(1).astype(np.uint32)
    nblk = coded.sum(1)
    mbs["coef_off"] = np.concatenate([[0], np.cumsum(nblk)[:-1]]).astype(np.uint32)
    coefs = lev[coded]  # row-major over (mb, block): blocks of an MB consecutive, increasing bit index
    # cbp
    ac_luma = (lev[:, :16, 1:] != 0).any(2) if True else None
    luma_q = np.where(i16[:, None], np.repeat(ac_luma.any(1)[:, None], 4, 1), luma_any.reshape(n, 4, 4).any(2))
    cbp_l = (luma_q * np.array([1, 2, 4, 8])).sum(1)
    c_ac = (lev[:, 16:, 1:] != 0).any((1, 2))
    c_dc = (lev[:, 16:, 0] != 0).any(1)
    cbp_c = np.where(c_ac, 2, np.where(c_dc, 1, 0))
    mbs["cbp"] = (cbp_l + 16 * cbp_c).astype(np.uint8)
    # nz_mask: luma in raster order of the blocks (I16x16: AC only); 8x8 transform marks the whole quadrant
    nz_z = np.where(i16[:, None], ac_luma, luma_any)
    z_of_r = np.array([_Z_OF[r >> 2, r & 3] for r in range(16)])
    nz_r = nz_z[:, z_of_r]
    nz = np.concatenate([nz_r, any_blk[:, 16:]], 1)
    mbs["nz_mask"] = (nz * weights).sum(1).astype(np.uint32)

    p = Picture(mb_w, mb_h, cur_slot)
    p.slices, p.mbs, p.coefs = slices, mbs, np.ascontiguousarray(coefs, np.int16)
    return p


def synthetic_sequence(mb_w, mb_h, n_pictures=4, seed=0, n_slots=3, qp=26, enable_8x8=True, two_slices=None):
    """I picture followed by P pictures (IPPP), slots rotating, one or two reference pictures."""
    rng = np.random.default_rng(seed)
    pics = []
    for t in range(n_pictures):
        cur = t % n_slots
        refs = [(t - 1) % n_slots] + ([(t - 2) % n_slots] if t >= 2 else [])
        ts = (seed + t) % 2 == 1 if two_slices is None else two_slices
        pics.append(synthetic_picture(rng, mb_w, mb_h, cur, refs if t else [], intra_only=(t == 0), qp=qp, enable_8x8=enable_8x8, two_slices=ts))
    return pics
