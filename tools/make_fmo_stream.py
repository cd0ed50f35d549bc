"""Writes small Baseline-profile streams that use slice groups (FMO map types 0 and 1 - the two the reference decodes,
codec/decoder/core/src/fmo.cpp:53-104), several slices per slice group and, optionally, arbitrary slice order (ASO).
The reference encoder cannot produce either, and no corpus stream has them, so this is a tiny encoder of its own: it does no
rate/distortion work, it only emits VALID syntax whose decoding exercises the macroblock walk (FmoNextMb), neighbour
availability across slice groups (intra modes, motion-vector prediction, P_Skip inference, CAVLC nC) and deblocking inside
slices.  Goldens come from the unmodified reference decoder (tests/golden/add_stream_golden.py).  Build container only.

Macroblock kinds: I_PCM (random samples), Intra16x16 without residual (a prediction mode its available neighbours allow),
P_Skip, P_L0_16x16 / 16x8 / 8x16 / P_8x8 with random motion-vector differences and no residual.

    python tools/make_fmo_stream.py out.264 --mbw 8 --mbh 6 --pics 6 --groups 3 --map 1 --slices-per-group 2 --aso --idc 2 --seed 5
"""
import argparse
import random


class BitWriter:
    def __init__(self):
        self.bits = []

    def u(self, n, v):
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def align_zero(self):
        while len(self.bits) & 7:
            self.bits.append(0)

    def bytes_(self, b):
        assert len(self.bits) % 8 == 0
        for x in b:
            self.u(8, x)

    def rbsp(self):
        b = self.bits + [1]
        while len(b) & 7:
            b.append(0)
        return bytes(sum(b[i + k] << (7 - k) for k in range(8)) for i in range(0, len(b), 8))


def escape(p):
    out, z = bytearray(), 0
    for b in p:
        if z >= 2 and b <= 3:
            out.append(3)
            z = 0
        out.append(b)
        z = z + 1 if b == 0 else 0
    return bytes(out)


def nal(ref_idc, typ, rbsp):
    return b"\x00\x00\x00\x01" + bytes([(ref_idc << 5) | typ]) + escape(rbsp)


def group_map(n, w, groups, typ, runs):
    m = [0] * n
    if typ == 0:
        i = 0
        while i < n:
            for g in range(groups):
                if i >= n:
                    break
                for j in range(runs[g]):
                    if i + j < n:
                        m[i + j] = g
                i += runs[g]
    else:
        for i in range(n):
            m[i] = ((i % w) + (((i // w) * groups) >> 1)) % groups
    return m


PCM, I16, SKIP, P16, P16x8, P8x16, P8x8 = range(7)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--mbw", type=int, default=8)
    ap.add_argument("--mbh", type=int, default=6)
    ap.add_argument("--pics", type=int, default=6)
    ap.add_argument("--groups", type=int, default=2)
    ap.add_argument("--map", type=int, default=1, choices=[0, 1])
    ap.add_argument("--slices-per-group", type=int, default=1)
    ap.add_argument("--aso", action="store_true")
    ap.add_argument("--idc", type=int, default=2, choices=[0, 1, 2], help="disable_deblocking_filter_idc of every slice")
    ap.add_argument("--qp", type=int, default=32)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-pcm", action="store_true", help="no I_PCM macroblocks (the recompression model does not carry their samples)")
    a = ap.parse_args()
    rng = random.Random(a.seed)
    w, h, n = a.mbw, a.mbh, a.mbw * a.mbh
    runs = [rng.randrange(1, w + 3) for _ in range(a.groups)]
    gmap = group_map(n, w, a.groups, a.map, runs)

    out = bytearray()
    sps = BitWriter()
    sps.u(8, 66), sps.u(8, 0), sps.u(8, 30)
    sps.ue(0)                       # sps id
    sps.ue(0)                       # log2_max_frame_num_minus4
    sps.ue(2)                       # pic_order_cnt_type 2: output order = decoding order
    sps.ue(1)                       # max_num_ref_frames
    sps.u(1, 0)                     # gaps_in_frame_num_value_allowed_flag
    sps.ue(w - 1), sps.ue(h - 1)
    sps.u(1, 1)                     # frame_mbs_only_flag
    sps.u(1, 0)                     # direct_8x8_inference_flag
    sps.u(1, 0)                     # frame_cropping_flag
    sps.u(1, 0)                     # vui_parameters_present_flag
    out += nal(3, 7, sps.rbsp())
    pps = BitWriter()
    pps.ue(0), pps.ue(0)
    pps.u(1, 0), pps.u(1, 0)        # CAVLC, no bottom_field_pic_order
    pps.ue(a.groups - 1)
    if a.groups > 1:
        pps.ue(a.map)
        if a.map == 0:
            for g in range(a.groups):
                pps.ue(runs[g] - 1)
    pps.ue(0), pps.ue(0)            # num_ref_idx defaults
    pps.u(1, 0), pps.u(2, 0)        # no weighted prediction
    pps.se(a.qp - 26), pps.se(0), pps.se(0)
    pps.u(1, 1)                     # deblocking_filter_control_present_flag
    pps.u(1, 0), pps.u(1, 0)        # constrained_intra_pred, redundant_pic_cnt_present
    out += nal(3, 8, pps.rbsp())

    for pic in range(a.pics):
        idr = pic == 0
        # slices: every slice group's macroblocks (in address order) cut into consecutive pieces
        slices = []
        for g in range(a.groups):
            mbs = [i for i in range(n) if gmap[i] == g]
            k = min(a.slices_per_group, len(mbs))
            cuts = sorted(rng.sample(range(1, len(mbs)), k - 1)) if k > 1 else []
            for lo, hi in zip([0] + cuts, cuts + [len(mbs)]):
                slices.append(mbs[lo:hi])
        slice_of = {}
        for si, sl in enumerate(slices):
            for m in sl:
                slice_of[m] = si
        kind = [None] * n
        order = list(range(len(slices)))
        if a.aso:
            rng.shuffle(order)
        for si in order:
            sl = slices[si]
            bw = BitWriter()
            bw.ue(sl[0])
            bw.ue(7 if idr else 5)
            bw.ue(0)
            bw.u(4, 0 if idr else pic % 16)
            if idr:
                bw.ue(0)                              # idr_pic_id
            if not idr:
                bw.u(1, 0)                            # num_ref_idx_active_override_flag
                bw.u(1, 0)                            # ref_pic_list_modification_flag_l0
            if idr:
                bw.u(1, 0), bw.u(1, 0)                # no_output_of_prior_pics, long_term_reference_flag
            else:
                bw.u(1, 0)                            # adaptive_ref_pic_marking_mode_flag
            bw.se(0)                                  # slice_qp_delta
            bw.ue(a.idc)
            if a.idc != 1:
                bw.se(rng.randrange(-2, 3)), bw.se(rng.randrange(-2, 3))

            def avail(m, dx, dy):
                x, y = m % w + dx, m // w + dy
                if x < 0 or x >= w or y < 0:
                    return None
                t = y * w + x
                return t if slice_of.get(t) == si and t < m else None

            def write_intra(m, base):
                if not a.no_pcm and rng.random() < 0.3:
                    kind[m] = PCM
                    bw.ue(base + 25)
                    bw.align_zero()
                    lo = rng.randrange(0, 200)
                    bw.bytes_(bytes(rng.randrange(lo, lo + 56) for _ in range(384)))
                    return
                kind[m] = I16
                L, T, TL = avail(m, -1, 0), avail(m, 0, -1), avail(m, -1, -1)
                modes = [2] + ([0] if T is not None else []) + ([1] if L is not None else []) + ([3] if None not in (L, T, TL) else [])
                bw.ue(base + 1 + rng.choice(modes))   # Intra16x16, cbp 0
                cm = [0] + ([1] if L is not None else []) + ([2] if T is not None else []) + ([3] if None not in (L, T, TL) else [])
                bw.ue(rng.choice(cm))                 # intra_chroma_pred_mode
                bw.se(0)                              # mb_qp_delta
                # Intra16x16DCLevel with no coefficients: coeff_token(TotalCoeff 0) under nC from the neighbours' block counts
                na = None if L is None else (16 if kind[L] == PCM else 0)
                nb = None if T is None else (16 if kind[T] == PCM else 0)
                nc = (na + nb + 1) >> 1 if None not in (na, nb) else (na if na is not None else (nb if nb is not None else 0))
                if nc < 2:
                    bw.u(1, 1)
                elif nc < 4:
                    bw.u(2, 3)
                elif nc < 8:
                    bw.u(4, 15)
                else:
                    bw.u(6, 3)

            def mvd():
                bw.se(rng.randrange(-7, 8)), bw.se(rng.randrange(-7, 8))

            if idr:
                for m in sl:
                    write_intra(m, 0)
            else:
                run = 0
                for m in sl:
                    r = rng.random()
                    if r < 0.35:
                        kind[m] = SKIP
                        run += 1
                        continue
                    bw.ue(run)
                    run = 0
                    if r < 0.50:
                        write_intra(m, 5)
                        continue
                    t = rng.choice([P16, P16, P16x8, P8x16, P8x8])
                    kind[m] = t
                    if t == P16:
                        bw.ue(0), mvd()
                    elif t == P16x8:
                        bw.ue(1), mvd(), mvd()
                    elif t == P8x16:
                        bw.ue(2), mvd(), mvd()
                    else:
                        bw.ue(3)
                        subs = [rng.randrange(4) for _ in range(4)]
                        for s_ in subs:
                            bw.ue(s_)
                        for s_ in subs:
                            for _ in range((1, 2, 2, 4)[s_]):
                                mvd()
                    bw.ue(0)                          # coded_block_pattern 0 (inter mapping: codeNum 0)
                if run:
                    bw.ue(run)
            out += nal(3 if idr else 2, 5 if idr else 1, bw.rbsp())
    open(a.out, "wb").write(out)


if __name__ == "__main__":
    main()
