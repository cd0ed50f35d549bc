"""CPU: crafted (not merely bit-flipped) inputs that round 1's byte-flip fuzzing could not reach.

1. A slice whose first_mb_in_slice is coded with 32 leading zero bits (ue() = 0xffffffff): stored in an int it became -1,
   passed the `>= n_mbs` test and indexed the macroblock arrays at a negative address (heap overflow under ASAN).
2. The same value in the default stream of a `.pip` (decompress direction): md_slice() indexed last[], tcoef[] and mb_slice[].
3. A picture parameter set with num_slice_groups_minus1 ~ 2^30 and slice_group_map_type 6: the parser shifted past bit 31 and
   looped for minutes (denial of service).  The reference drops such a parameter set (au_parser.cpp:1315, :1322-1326).
Every case must be refused quickly and cleanly (exit code of a parse error, no sanitizer report, no hang)."""
import os
import subprocess

import pytest

from streams import LIB_DIR, STREAM_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "losslessh264_b200", "host")
CORE = ["decoder.cpp", "h264_syntax.cpp", "cavlc.cpp", "cavlc_writer.cpp", "pip_container.cpp", "pip_expand.cpp"]
SAN = ["-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-w"]
ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:exitcode=99", UBSAN_OPTIONS="halt_on_error=1:exitcode=98")


class Bits:
    def __init__(self):
        self.b = []

    def u(self, n, v):
        self.b += [(v >> (n - 1 - i)) & 1 for i in range(n)]

    def ue(self, v):
        n = (v + 1).bit_length()
        self.u(n - 1, 0)
        self.u(n, v + 1)

    def raw_zero_prefix_ue(self, zeros, info):
        """`zeros` leading zero bits, the marker bit, then `zeros` info bits (any value): what a 32-bit ue() reader sees as overflow"""
        self.u(zeros, 0)
        self.u(1, 1)
        self.u(zeros, info)

    def rbsp(self):
        bits = self.b + [1]
        bits += [0] * (-len(bits) % 8)
        return bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))


def escape(payload):
    out, zeros = bytearray(), 0
    for x in payload:
        if zeros >= 2 and x <= 3:
            out.append(3)
            zeros = 0
        out.append(x)
        zeros = zeros + 1 if x == 0 else 0
    return bytes(out)


def nal(header, bits):
    return b"\x00\x00\x00\x01" + bytes([header]) + escape(bits.rbsp())


def split_units(data):
    idx, i = [], 0
    while True:
        j = data.find(b"\x00\x00\x01", i)
        if j < 0:
            break
        idx.append(j)
        i = j + 3
    return [data[a:b] for a, b in zip(idx, idx[1:] + [len(data)])]


@pytest.fixture(scope="module")
def asan_tools(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("asan"))
    parse = os.path.join(d, "parse_asan")
    subprocess.check_call(["g++"] + SAN + ["-o", parse] + [os.path.join(HOST, f) for f in ["tool_parse_dump.cpp", "sinks_file.cpp"] + CORE])
    expand = os.path.join(d, "expand_asan")
    subprocess.check_call(["g++"] + SAN + ["-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-I" + os.path.join(ROOT, "losslessh264_b200", "csrc"),
                           "-o", expand, os.path.join(ROOT, "oracle", "model_decode_cpu.cpp")] + [os.path.join(HOST, f) for f in CORE])
    return parse, expand


def crafted_first_mb_stream():
    """SPS + PPS of a real QCIF stream, then an IDR slice whose first_mb_in_slice has 32 leading zero bits."""
    units = split_units(open(os.path.join(STREAM_DIR, "SVA_BA1_B.264"), "rb").read())
    head = b"".join(u for u in units if (u[3] & 31) in (7, 8))
    s = Bits()
    s.raw_zero_prefix_ue(32, 0)      # first_mb_in_slice: ue() returns 0xffffffff
    s.ue(7)                          # slice_type I
    s.ue(0)                          # pps id
    s.u(64, 0x123456789ABCDEF0)      # whatever follows
    return head + nal(0x65, s)


def test_negative_first_mb_in_slice_is_refused_by_the_parser(asan_tools, tmp_path):
    parse, _ = asan_tools
    case = os.path.join(str(tmp_path), "c.264")
    open(case, "wb").write(crafted_first_mb_stream())
    for extra in ([], ["1", os.path.join(str(tmp_path), "c.lhmd"), os.path.join(str(tmp_path), "c.pip")]):
        r = subprocess.run([parse, case, os.path.join(str(tmp_path), "c.lhfb")] + extra, capture_output=True, text=True, timeout=30, env=ENV)
        assert r.returncode in (0, 1, 2, 3, 4, 5), f"rc {r.returncode}\n{r.stderr[-800:]}"
        assert "pictures 0 " in r.stderr, r.stderr[-300:]


def test_first_mb_beyond_the_picture_is_refused_in_a_pip_default_stream(asan_tools, tmp_path):
    """Decompress direction: a real `.pip` whose first slice header is replaced by one naming a macroblock outside the picture.  The
    planner (host) refuses it; the decompress core itself refuses it too when handed the slice directly."""
    parse, expand = asan_tools
    d = str(tmp_path)
    pip = os.path.join(d, "x.pip")
    r = subprocess.run([os.path.join(LIB_DIR, "lh264_parse_dump"), os.path.join(STREAM_DIR, "SVA_BA1_B.264"), os.path.join(d, "d.lhfb"), "1", os.path.join(d, "d.lhmd"), pip], capture_output=True)
    assert r.returncode == 0
    data = open(pip, "rb").read()
    units = split_units(data)
    k = next(i for i, u in enumerate(units) if (u[3] & 31) in (1, 5))
    for first_mb, zeros in ((99, None), (5000, None), (None, 32)):      # QCIF has 99 macroblocks: 99 is the first address outside
        s = Bits()
        if zeros:
            s.raw_zero_prefix_ue(zeros, 0)
        else:
            s.ue(first_mb)
        s.ue(7); s.ue(0); s.u(4, 0); s.ue(0); s.u(4, 0); s.u(2, 0); s.ue(0)          # a plausible rest of an IDR header
        crafted = data[: data.find(units[k])] + b"\x00\x00\x01" + bytes([units[k][3]]) + escape(s.rbsp()) + b"".join(units[k + 1:])
        open(pip, "wb").write(crafted)
        r = subprocess.run([expand, pip, os.path.join(d, "o.264")], capture_output=True, text=True, timeout=60, env=ENV)
        assert r.returncode not in (0, 98, 99) and r.returncode > 0, f"first_mb {first_mb}: rc {r.returncode}\n{r.stderr[-800:]}"


def test_huge_slice_group_count_does_not_hang_the_pps_parser(asan_tools, tmp_path):
    parse, _ = asan_tools
    units = split_units(open(os.path.join(STREAM_DIR, "SVA_BA1_B.264"), "rb").read())
    sps = next(u for u in units if (u[3] & 31) == 7)
    p = Bits()
    p.ue(0); p.ue(0); p.u(1, 0); p.u(1, 0)
    p.ue((1 << 30) - 1)              # num_slice_groups_minus1
    p.ue(6)                          # slice_group_map_type 6
    p.ue((1 << 30) - 1)              # pic_size_in_map_units_minus1
    p.u(64, 0)
    case = os.path.join(str(tmp_path), "c.264")
    open(case, "wb").write(sps + nal(0x68, p) + b"".join(u for u in units if (u[3] & 31) in (1, 5))[:4000])
    r = subprocess.run([parse, case, os.path.join(str(tmp_path), "c.lhfb")], capture_output=True, text=True, timeout=20, env=ENV)
    assert r.returncode in (0, 1, 2, 3, 4, 5), f"rc {r.returncode}\n{r.stderr[-800:]}"
    assert "pictures 0 " in r.stderr
