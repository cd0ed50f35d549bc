"""Gives an existing Annex-B stream scaling lists: rewrites its SPS / PPS NAL units (everything else is copied verbatim) so that
the SAME quantised levels are dequantised through seq / pic scaling matrices.  The reference encoder cannot write scaling
lists and the only corpus stream that carries them (test_scalinglist_jm) is refused by the reference decoder, so these are
the streams the scaling-list path (WelsCalcDeqCoeffScalingList, decode_slice.cpp:1231-1277) is checked on;
the goldens come from the unmodified reference decoder (tests/golden/add_stream_golden.py).  Build container only.

    python tools/make_scaling_stream.py in.264 out.264 --sps-lists SEED --pps-lists SEED

A stream whose SPS is not High profile is re-labelled profile_idc 100 (chroma_format_idc 1, 8 bit), which changes nothing else
of its decoding.  List i of a matrix is, depending on the seed, absent (fall-back rule), "use default" (first delta -8) or
explicit with values that wander between 4 and --max-scale.
"""
import argparse
import random


class BitReader:
    def __init__(self, data):
        self.d, self.p = data, 0

    def u(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | ((self.d[self.p >> 3] >> (7 - (self.p & 7))) & 1)
            self.p += 1
        return v

    def ue(self):
        z = 0
        while self.u(1) == 0:
            z += 1
        return (1 << z) - 1 + (self.u(z) if z else 0)

    def se(self):
        k = self.ue()
        return (k + 1) // 2 if k & 1 else -(k // 2)


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

    def copy(self, br, upto):
        while br.p < upto:
            self.bits.append(br.u(1))

    def rbsp(self):
        b = self.bits + [1]
        while len(b) & 7:
            b.append(0)
        return bytes(sum(b[i + k] << (7 - k) for k in range(8)) for i in range(0, len(b), 8))


def unescape(p):
    out, z = bytearray(), 0
    for b in p:
        if z >= 2 and b == 3:
            z = 0
            continue
        out.append(b)
        z = z + 1 if b == 0 else 0
    return bytes(out)


def escape(p):
    out, z = bytearray(), 0
    for b in p:
        if z >= 2 and b <= 3:
            out.append(3)
            z = 0
        out.append(b)
        z = z + 1 if b == 0 else 0
    return bytes(out)


def rbsp_end(data):
    """bit position of the rbsp stop bit"""
    i = len(data) - 1
    while data[i] == 0:
        i -= 1
    b = data[i]
    k = 0
    while not (b >> k) & 1:
        k += 1
    return 8 * i + 7 - k


ZIG4 = [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15]


def write_lists(bw, rng, n_lists, hi=60):
    for i in range(n_lists):
        kind = rng.choice(["absent", "default", "explicit", "explicit", "short"])
        if kind == "absent":
            bw.u(1, 0)
            continue
        bw.u(1, 1)
        size = 16 if i < 6 else 64
        if kind == "default":
            bw.se(-8)                                  # next_scale becomes 0 at j == 0: use the default matrix
            continue
        last = 8
        stop_at = rng.randrange(3, size) if kind == "short" else size
        for j in range(size):
            if j == stop_at:
                bw.se(-last if last <= 128 else 256 - last)   # next_scale = 0 (mod 256): the rest of the list repeats last_scale
                break
            nxt = min(hi, max(4, last + rng.randrange(-6, 9) * (1 if hi <= 60 else 6)))
            bw.se(nxt - last)
            last = nxt


def rewrite_sps(p, seed, hi=60):
    br = BitReader(p)
    bw = BitWriter()
    profile = br.u(8)
    flags, level = br.u(8), br.u(8)
    sps_id = br.ue()
    high = profile in (100, 110, 122, 244, 44, 83, 86, 118, 128)
    bw.u(8, profile if high else 100)
    bw.u(8, flags if high else 0)
    bw.u(8, level)
    bw.ue(sps_id)
    if high:
        cf = br.ue()
        assert cf == 1
        bdl, bdc, bypass = br.ue(), br.ue(), br.u(1)
        assert br.u(1) == 0, "the stream already carries a seq scaling matrix"
    else:
        cf, bdl, bdc, bypass = 1, 0, 0, 0
    bw.ue(cf), bw.ue(bdl), bw.ue(bdc), bw.u(1, bypass)
    if seed is None:
        bw.u(1, 0)
    else:
        bw.u(1, 1)
        write_lists(bw, random.Random(seed), 8, hi)
    bw.copy(br, rbsp_end(p))
    return bw.rbsp()


def rewrite_pps(p, seed, hi=60):
    br = BitReader(p)
    bw = BitWriter()
    end = rbsp_end(p)
    bw.ue(br.ue()), bw.ue(br.ue())                     # pps id, sps id
    bw.u(1, br.u(1)), bw.u(1, br.u(1))                 # entropy coding mode, bottom_field_pic_order
    nsg = br.ue()
    assert nsg == 0
    bw.ue(0)
    bw.ue(br.ue()), bw.ue(br.ue())                     # num_ref_idx defaults
    bw.u(1, br.u(1)), bw.u(2, br.u(2))                 # weighted pred, bipred idc
    bw.se(br.se()), bw.se(br.se())                     # pic_init_qp / qs
    cqo = br.se()
    bw.se(cqo)
    bw.u(3, br.u(3))                                   # deblocking control, constrained intra, redundant pic cnt
    t8, second = 0, cqo
    if br.p < end:
        t8 = br.u(1)
        assert br.u(1) == 0, "the stream already carries a pic scaling matrix"
        second = br.se()
    bw.u(1, t8)
    if seed is None:
        bw.u(1, 0)
    else:
        bw.u(1, 1)
        write_lists(bw, random.Random(seed), 8, hi)    # the reference reads 8 lists whatever transform_8x8_mode_flag says
    bw.se(second)
    return bw.rbsp()


def split_annexb(data):
    """[(start code bytes, nal bytes)]"""
    out, i, n = [], 0, len(data)
    starts = []
    while i + 3 <= n:
        if data[i] == 0 and data[i + 1] == 0 and data[i + 2] == 1:
            starts.append(i)
            i += 3
        else:
            i += 1
    for k, s in enumerate(starts):
        e = starts[k + 1] if k + 1 < len(starts) else n
        pre = s
        while pre > (starts[k - 1] + 3 if k else 0) and data[pre - 1] == 0:
            pre -= 1
        nal_end = e
        if k + 1 < len(starts):
            while nal_end > s + 3 and data[nal_end - 1] == 0:
                nal_end -= 1
        out.append((data[pre:s + 3], data[s + 3:nal_end]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--sps-lists", type=int, default=None)
    ap.add_argument("--pps-lists", type=int, default=None)
    ap.add_argument("--max-scale", type=int, default=60, help="largest list entry (above 60: multipliers overflow the reference's 16-bit tables at high QP)")
    a = ap.parse_args()
    data = open(a.src, "rb").read()
    out = bytearray()
    for sc, nal in split_annexb(data):
        t = nal[0] & 31
        if t == 7:
            nal = nal[:1] + escape(rewrite_sps(unescape(nal[1:]), a.sps_lists, a.max_scale))
        elif t == 8:
            nal = nal[:1] + escape(rewrite_pps(unescape(nal[1:]), a.pps_lists, a.max_scale))
        out += (sc if len(sc) >= 3 else b"\x00\x00\x01") + nal
    open(a.dst, "wb").write(out)


if __name__ == "__main__":
    main()
