"""GPU: kernel-level anchors for the two stages that round 1 only checked at picture level — intra prediction and the deblocking
sample filters — against a SECOND oracle written here in numpy straight from the equations of ITU-T H.264 (8.3.1.2 Intra_4x4,
8.3.3 Intra_16x16, 8.3.4 chroma, 8.7.2.3 / 8.7.2.4 edge filters), the way the reference's own unit tests anchor its function
tables (test/decoder/DecUT_IntraPrediction.cpp:95-353 recompute every predictor from the neighbours;
DecUT_DeblockCommon.cpp:90-256 recompute the filters).  The pictures are built record by record: I_PCM macroblocks carry chosen
samples (so the neighbours of the macroblock under test are known exactly), the macroblock under test has no residual, and the
deblocking filter is switched per slice."""
import numpy as np
import pytest

from losslessh264_b200.records import MB_I4x4, MB_I16x16, MB_IPCM, MB_REC, SLICE_REC, Picture

Z2XY = [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (3, 0), (2, 1), (3, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 2), (3, 2), (2, 3), (3, 3)]   # z -> (bx, by)


class OracleBackend:
    """The first oracle (oracle/recon_oracle.c, pinned to the reference) behind the calls the tests make: the numpy formulas below
    are checked against it on the CPU before they judge the CUDA path."""

    def __init__(self):
        self.streams = {}

    def open_stream(self, mb_w, mb_h, n_slots):
        from oracle import oracle

        self.streams[len(self.streams)] = oracle.OracleStream(mb_w, mb_h, n_slots)
        return len(self.streams) - 1

    def recon(self, pics, streams):
        self.last = {s: self.streams[s].recon(p) for p, s in zip(pics, streams)}

    def read_picture(self, s, slot, mb_w, mb_h):
        return self.last[s]

    def close_stream(self, s):
        pass

    def close(self):
        pass


@pytest.fixture(scope="module", params=[pytest.param("oracle"), pytest.param("gpu", marks=pytest.mark.gpu)])
def ctx(request):
    if request.param == "oracle":
        c = OracleBackend()
    else:
        from losslessh264_b200.api import Context

        c = Context(0)
    yield c
    c.close()


def pcm_picture(rng, mb_w, mb_h, idc=1, qp=0):
    """A picture of I_PCM macroblocks with random samples; returns (Picture, Y, Cb, Cr) with the planes the records encode."""
    n = mb_w * mb_h
    mbs = np.zeros(n, MB_REC)
    Y = rng.integers(0, 256, (mb_h * 16, mb_w * 16), dtype=np.uint8)
    Cb = rng.integers(0, 256, (mb_h * 8, mb_w * 8), dtype=np.uint8)
    Cr = rng.integers(0, 256, (mb_h * 8, mb_w * 8), dtype=np.uint8)
    coefs = np.zeros((n * 24, 16), np.int16)
    for a in range(n):
        x, y = a % mb_w, a // mb_w
        mbs[a]["mb_type"] = MB_IPCM
        mbs[a]["coded_mask"] = 0xFFFFFF
        mbs[a]["nz_mask"] = 0xFFFFFF
        mbs[a]["coef_off"] = a * 24
        mbs[a]["ref_idx"] = -1
        mbs[a]["qp_y"] = qp
        mbs[a]["qp_c"] = qp
        flat = np.concatenate([Y[y * 16:y * 16 + 16, x * 16:x * 16 + 16].ravel(), Cb[y * 8:y * 8 + 8, x * 8:x * 8 + 8].ravel(), Cr[y * 8:y * 8 + 8, x * 8:x * 8 + 8].ravel()])
        coefs[a * 24:(a + 1) * 24] = flat.astype(np.int16).reshape(24, 16)
    sl = np.zeros(1, SLICE_REC)
    sl[0]["first_mb"], sl[0]["num_mbs"], sl[0]["slice_type"], sl[0]["disable_deblocking_idc"] = 0, n, 2, idc
    return Picture(mb_w=mb_w, mb_h=mb_h, cur_slot=0, slices=sl, mbs=mbs, coefs=coefs), Y, Cb, Cr


def planes(flat, mb_w, mb_h):
    W, H = mb_w * 16, mb_h * 16
    return flat[:W * H].reshape(H, W), flat[W * H:W * H * 5 // 4].reshape(H // 2, W // 2), flat[W * H * 5 // 4:].reshape(H // 2, W // 2)


def clip1(v):
    return int(min(255, max(0, v)))


# ---- 8.3.3 Intra_16x16 and 8.3.4 chroma, from the neighbouring samples --------------------------------------------------------
def pred16(mode, top, left, tl):
    p = np.zeros((16, 16), np.int64)
    if mode == 0:
        p[:] = top[None, :]
    elif mode == 1:
        p[:] = left[:, None]
    elif mode == 2:
        p[:] = (int(top.sum()) + int(left.sum()) + 16) >> 5
    else:
        t = np.concatenate([[tl], top]).astype(np.int64)      # t[1 + x] = p[x, -1]
        l = np.concatenate([[tl], left]).astype(np.int64)
        Hh = sum((k + 1) * (t[1 + 8 + k] - t[1 + 6 - k]) for k in range(8))
        Vv = sum((k + 1) * (l[1 + 8 + k] - l[1 + 6 - k]) for k in range(8))
        a, b, c = 16 * (int(left[15]) + int(top[15])), (5 * Hh + 32) >> 6, (5 * Vv + 32) >> 6
        for y in range(16):
            for x in range(16):
                p[y, x] = clip1((a + b * (x - 7) + c * (y - 7) + 16) >> 5)
    return p.astype(np.uint8)


def predc(mode, top, left, tl):            # record numbering: 0 DC, 1 H, 2 V, 3 plane (LH264_C_*)
    p = np.zeros((8, 8), np.int64)
    if mode == 2:
        p[:] = top[None, :]
    elif mode == 1:
        p[:] = left[:, None]
    elif mode == 0:
        t0, t1, l0, l1 = int(top[:4].sum()), int(top[4:].sum()), int(left[:4].sum()), int(left[4:].sum())
        p[:4, :4] = (t0 + l0 + 4) >> 3
        p[:4, 4:] = (t1 + 2) >> 2
        p[4:, :4] = (l1 + 2) >> 2
        p[4:, 4:] = (t1 + l1 + 4) >> 3
    else:
        t = np.concatenate([[tl], top]).astype(np.int64)
        l = np.concatenate([[tl], left]).astype(np.int64)
        Hh = sum((k + 1) * (t[1 + 4 + k] - t[1 + 2 - k]) for k in range(4))
        Vv = sum((k + 1) * (l[1 + 4 + k] - l[1 + 2 - k]) for k in range(4))
        a, b, c = 16 * (int(left[7]) + int(top[7])), (34 * Hh + 32) >> 6, (34 * Vv + 32) >> 6
        for y in range(8):
            for x in range(8):
                p[y, x] = clip1((a + b * (x - 3) + c * (y - 3) + 16) >> 5)
    return p.astype(np.uint8)


@pytest.mark.parametrize("i16,chroma", [(0, 2), (1, 1), (2, 0), (3, 3), (3, 0), (0, 3)])
def test_intra16x16_and_chroma_predictors_against_the_standard(ctx, i16, chroma):
    rng = np.random.default_rng(100 + 10 * i16 + chroma)
    pic, Y, Cb, Cr = pcm_picture(rng, 3, 3)
    c = 4                                       # the centre macroblock: all eight neighbours are I_PCM
    m = pic.mbs[c]
    m["mb_type"], m["i16_mode"], m["chroma_mode"], m["cbp"] = MB_I16x16, i16, chroma, 0
    m["coded_mask"], m["nz_mask"], m["coef_off"], m["qp_y"], m["qp_c"] = 0, 0, 0, 26, 26
    s = ctx.open_stream(3, 3, 2)
    try:
        ctx.recon([pic], [s])
        gy, gu, gv = planes(ctx.read_picture(s, 0, 3, 3), 3, 3)
    finally:
        ctx.close_stream(s)
    want = pred16(i16, Y[15, 16:32], Y[16:32, 15], int(Y[15, 15]))
    assert np.array_equal(gy[16:32, 16:32], want), f"Intra16x16 mode {i16}"
    for got, P in ((gu, Cb), (gv, Cr)):
        assert np.array_equal(got[8:16, 8:16], predc(chroma, P[7, 8:16], P[8:16, 7], int(P[7, 7]))), f"chroma mode {chroma}"
    # the I_PCM neighbours come back untouched (deblocking is off for the slice)
    assert np.array_equal(gy[:16], Y[:16]) and np.array_equal(gy[32:], Y[32:])


# ---- 8.3.1.2 Intra_4x4 ---------------------------------------------------------------------------------------------------------
def pred4(mode, p, x0, y0, tr_avail):
    """Predicts block at (x0, y0) of plane p (int array with the samples decoded so far) with mode 0..8."""
    t = [int(p[y0 - 1, x0 + i]) for i in range(4)]
    t += [int(p[y0 - 1, x0 + 4 + i]) for i in range(4)] if tr_avail else [t[3]] * 4
    l = [int(p[y0 + i, x0 - 1]) for i in range(4)]
    q = int(p[y0 - 1, x0 - 1])
    o = np.zeros((4, 4), np.int64)
    e = [l[3], l[2], l[1], l[0], q] + t      # e[4 + 1 + x] = p[x, -1], e[3 - y] = p[-1, y], e[4] = p[-1, -1]

    def P(x, y):                               # p[x, y] for x = -1 or y = -1
        return e[4 + 1 + x] if y == -1 else e[3 - y]
    for y in range(4):
        for x in range(4):
            if mode == 0:
                v = t[x]
            elif mode == 1:
                v = l[y]
            elif mode == 2:
                v = (sum(t[:4]) + sum(l) + 4) >> 3
            elif mode == 3:                    # diagonal down left
                v = (t[6] + 3 * t[7] + 2) >> 2 if x == 3 and y == 3 else (t[x + y] + 2 * t[x + y + 1] + t[x + y + 2] + 2) >> 2
            elif mode == 4:                    # diagonal down right
                if x > y:
                    v = (P(x - y - 2, -1) + 2 * P(x - y - 1, -1) + P(x - y, -1) + 2) >> 2
                elif x < y:
                    v = (P(-1, y - x - 2) + 2 * P(-1, y - x - 1) + P(-1, y - x) + 2) >> 2
                else:
                    v = (t[0] + 2 * q + l[0] + 2) >> 2
            elif mode == 5:                    # vertical right
                zvr = 2 * x - y
                if zvr >= 0 and zvr % 2 == 0:
                    v = (P(x - (y >> 1) - 1, -1) + P(x - (y >> 1), -1) + 1) >> 1
                elif zvr >= 0:
                    v = (P(x - (y >> 1) - 2, -1) + 2 * P(x - (y >> 1) - 1, -1) + P(x - (y >> 1), -1) + 2) >> 2
                elif zvr == -1:
                    v = (l[0] + 2 * q + t[0] + 2) >> 2
                else:
                    v = (P(-1, y - 1) + 2 * P(-1, y - 2) + P(-1, y - 3) + 2) >> 2
            elif mode == 6:                    # horizontal down
                zhd = 2 * y - x
                if zhd >= 0 and zhd % 2 == 0:
                    v = (P(-1, y - (x >> 1) - 1) + P(-1, y - (x >> 1)) + 1) >> 1
                elif zhd >= 0:
                    v = (P(-1, y - (x >> 1) - 2) + 2 * P(-1, y - (x >> 1) - 1) + P(-1, y - (x >> 1)) + 2) >> 2
                elif zhd == -1:
                    v = (l[0] + 2 * q + t[0] + 2) >> 2
                else:
                    v = (P(x - 1, -1) + 2 * P(x - 2, -1) + P(x - 3, -1) + 2) >> 2
            elif mode == 7:                    # vertical left
                v = (t[x + (y >> 1)] + t[x + (y >> 1) + 1] + 1) >> 1 if y % 2 == 0 else (t[x + (y >> 1)] + 2 * t[x + (y >> 1) + 1] + t[x + (y >> 1) + 2] + 2) >> 2
            else:                              # horizontal up
                zhu = x + 2 * y
                if zhu > 5:
                    v = l[3]
                elif zhu == 5:
                    v = (l[2] + 3 * l[3] + 2) >> 2
                elif zhu % 2 == 0:
                    v = (l[y + (x >> 1)] + l[y + (x >> 1) + 1] + 1) >> 1
                else:
                    v = (l[y + (x >> 1)] + 2 * l[y + (x >> 1) + 1] + l[y + (x >> 1) + 2] + 2) >> 2
            o[y, x] = v
    return o


@pytest.mark.parametrize("mode", range(9))
def test_intra4x4_predictors_against_the_standard(ctx, mode):
    rng = np.random.default_rng(200 + mode)
    pic, Y, Cb, Cr = pcm_picture(rng, 3, 3)
    m = pic.mbs[4]
    m["mb_type"], m["chroma_mode"], m["cbp"] = MB_I4x4, 0, 0
    m["coded_mask"], m["nz_mask"], m["coef_off"], m["qp_y"], m["qp_c"] = 0, 0, 0, 26, 26
    # every block uses `mode`; where the top-right neighbour does not exist yet the hand-off names the replicate variant
    # (LH264_I4_DDL_TOP 12 / LH264_I4_VL_TOP 13, include/lh264_b200.h)
    work = Y.astype(np.int64).copy()
    for z, (bx, by) in enumerate(Z2XY):
        tr = (by == 0) or (bx < 3 and Z2XY.index((bx + 1, by - 1)) < z)      # top MB row is available (I_PCM); (3, 0) uses the top-right MB
        m["intra_mode"][z] = {3: 12, 7: 13}.get(mode, mode) if not tr and mode in (3, 7) else mode
        x0, y0 = 16 + 4 * bx, 16 + 4 * by
        work[y0:y0 + 4, x0:x0 + 4] = pred4(mode, work, x0, y0, tr)
    m["nxn_avail"] = 15
    s = ctx.open_stream(3, 3, 2)
    try:
        ctx.recon([pic], [s])
        gy, _, _ = planes(ctx.read_picture(s, 0, 3, 3), 3, 3)
    finally:
        ctx.close_stream(s)
    assert np.array_equal(gy[16:32, 16:32], work[16:32, 16:32].astype(np.uint8)), f"Intra4x4 mode {mode}"


# ---- 8.7.2.3 / 8.7.2.4: the edges between and inside two flat intra macroblocks ------------------------------------------------
ALPHA = [0] * 16 + [4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255]
BETA = [0] * 16 + [2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18]


# Table 8-17: tc0 for bS = 1, 2, 3 by indexA (zero below 17)
TC0 = [(0, 0, 0)] * 17 + [(0, 0, 1), (0, 0, 1), (0, 0, 1), (0, 0, 1), (0, 1, 1), (0, 1, 1), (1, 1, 1), (1, 1, 1), (1, 1, 1), (1, 1, 1), (1, 1, 2), (1, 1, 2), (1, 1, 2),
                          (1, 1, 2), (1, 2, 3), (1, 2, 3), (2, 2, 3), (2, 2, 4), (2, 3, 4), (2, 3, 4), (3, 3, 5), (3, 4, 6), (3, 4, 6), (4, 5, 7), (4, 5, 8), (4, 6, 9),
                          (5, 7, 10), (6, 8, 11), (6, 8, 13), (7, 10, 14), (8, 11, 16), (9, 12, 18), (10, 13, 20), (11, 15, 23), (13, 17, 25)]


def filter_luma_lt4(p, q, bs, index_a, alpha, beta):
    """8.7.2.3 for luma, bS 1..3: p[0..2], q[0..2]; returns the filtered copies."""
    P, Q = list(p), list(q)
    if not (abs(p[0] - q[0]) < alpha and abs(p[1] - p[0]) < beta and abs(q[1] - q[0]) < beta):
        return P, Q
    tc0 = TC0[index_a][bs - 1]
    ap, aq = abs(p[2] - p[0]), abs(q[2] - q[0])
    tc = tc0 + (ap < beta) + (aq < beta)
    d = min(tc, max(-tc, (((q[0] - p[0]) << 2) + (p[1] - q[1]) + 4) >> 3))
    P[0], Q[0] = clip1(p[0] + d), clip1(q[0] - d)
    if ap < beta:
        P[1] = p[1] + min(tc0, max(-tc0, (p[2] + ((p[0] + q[0] + 1) >> 1) - (p[1] << 1)) >> 1))
    if aq < beta:
        Q[1] = q[1] + min(tc0, max(-tc0, (q[2] + ((p[0] + q[0] + 1) >> 1) - (q[1] << 1)) >> 1))
    return P, Q


def filter_bs4(p, q, alpha, beta, chroma):
    """One line across an edge with bS = 4: p[0..3] and q[0..3] (p[0], q[0] next to the edge); returns the filtered copies."""
    P, Q = list(p), list(q)
    if not (abs(p[0] - q[0]) < alpha and abs(p[1] - p[0]) < beta and abs(q[1] - q[0]) < beta):
        return P, Q
    small = abs(p[0] - q[0]) < ((alpha >> 2) + 2)
    if not chroma and small and abs(p[2] - p[0]) < beta:
        P[0] = (p[2] + 2 * p[1] + 2 * p[0] + 2 * q[0] + q[1] + 4) >> 3
        P[1] = (p[2] + p[1] + p[0] + q[0] + 2) >> 2
        P[2] = (2 * p[3] + 3 * p[2] + p[1] + p[0] + q[0] + 4) >> 3
    else:
        P[0] = (2 * p[1] + p[0] + q[1] + 2) >> 2
    if not chroma and small and abs(q[2] - q[0]) < beta:
        Q[0] = (p[1] + 2 * p[0] + 2 * q[0] + 2 * q[1] + q[2] + 4) >> 3
        Q[1] = (p[0] + q[0] + q[1] + q[2] + 2) >> 2
        Q[2] = (2 * q[3] + 3 * q[2] + q[1] + q[0] + p[0] + 4) >> 3
    else:
        Q[0] = (2 * q[1] + q[0] + p[1] + 2) >> 2
    return P, Q


@pytest.mark.parametrize("qp,step", [(30, 3), (36, 9), (44, 20), (51, 40), (20, 2)])
def test_macroblock_edge_filter_bs4_against_the_standard(ctx, qp, step):
    """A 2x1 picture of I_PCM macroblocks with flat values v and v + step; the records carry the qp the filter is to use on both
    sides (the edge filters read qp_y / qp_c of the two macroblocks, whatever their type)."""
    rng = np.random.default_rng(300 + qp)
    pic, Y, Cb, Cr = pcm_picture(rng, 2, 1, idc=0, qp=qp)
    v = int(rng.integers(40, 180))
    for a, val in ((0, v), (1, v + step)):
        pic.coefs[a * 24:(a + 1) * 24] = val
    Y[:, :16], Y[:, 16:], Cb[:, :8], Cb[:, 8:], Cr[:, :8], Cr[:, 8:] = v, v + step, v, v + step, v, v + step
    s = ctx.open_stream(2, 1, 2)
    try:
        ctx.recon([pic], [s])
        gy, gu, gv = planes(ctx.read_picture(s, 0, 2, 1), 2, 1)
    finally:
        ctx.close_stream(s)
    # luma, one row, in the standard's order: the left macroblock's edges see flat samples; then the macroblock edge at x = 16 with
    # bS 4 (both sides intra, qp the average of two equal qps); then the right macroblock's internal edge at x = 20 with bS 3, which
    # sees what the macroblock edge left behind in x = 16..19 (the later edges at 24 and 28 are out of reach of x < 22)
    row = [v] * 16 + [v + step] * 16
    P, Q = filter_bs4(row[15:11:-1], row[16:20], ALPHA[qp], BETA[qp], False)
    row[12:16], row[16:20] = P[::-1], Q
    P, Q = filter_luma_lt4(row[19:16:-1], row[20:23], 3, qp, ALPHA[qp], BETA[qp])
    row[17:20], row[20:23] = P[::-1], Q
    assert np.array_equal(gy[5, 10:24], np.array(row[10:24], np.uint8)), f"luma edges qp {qp} step {step}: {gy[5, 10:24]} vs {row[10:24]}"
    assert np.all(gy[:, :12] == v)
    Pc, Qc = filter_bs4([v] * 4, [v + step] * 4, ALPHA[qp], BETA[qp], True)     # (the records carry qp_c = qp: the filter indexes with it as is)
    for g in (gu, gv):
        assert np.array_equal(g[3, 6:10], np.array([v, Pc[0], Qc[0], v + step], np.uint8)), f"chroma edge qp {qp}"
