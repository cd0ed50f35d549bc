"""CPU: the host decoder (NAL/slice/CAVLC parse -> hand-off records) checked end to end through the
oracle: parsed records -> oracle reconstruction -> SHA1 of the cropped I420 == the reference's SHA1
(tests/golden/stream_sha1.json: 36 of them are the constants of the reference's
test/api/decoder_test.cpp:90-128, the rest were measured from the compiled reference)."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

from losslessh264_b200.records import n_slots_needed, read_lhfb
from oracle import oracle
from streams import LIB_DIR, SHA1, STREAM_DIR, decodable_streams

TOOL = os.path.join(LIB_DIR, "lh264_parse_dump")


@pytest.fixture(scope="module", autouse=True)
def _built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(LIB_DIR), "host"), "../_lib/lh264_parse_dump"])


def decode_with_oracle(stream, tmp_path):
    out = os.path.join(tmp_path, "d.lhfb")
    r = subprocess.run([TOOL, stream, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    m = re.search(r"crop (\d+) (\d+) (\d+) (\d+)", r.stderr)
    crop = tuple(int(v) for v in m.groups()) if m else (0, 0, 0, 0)      # (no line: nothing was handed out)
    pics = read_lhfb(out)
    cx, cy, cw, ch = crop
    st, geo = None, None
    h = hashlib.sha1()
    n_out = 0
    for p in pics:
        if (p.mb_w, p.mb_h) != geo:         # a new sequence of another size: new picture storage (Error_I_P)
            geo = (p.mb_w, p.mb_h)
            st = oracle.OracleStream(p.mb_w, p.mb_h, n_slots_needed(pics))
        mb_w, mb_h = geo
        W, H = mb_w * 16, mb_h * 16
        o = st.recon(p)
        if not p.output:                    # concealed and lost to a DecodeFrameNoDelay caller: a reference only
            continue
        n_out += 1
        h.update(np.ascontiguousarray(o[: W * H].reshape(H, W)[cy : cy + ch, cx : cx + cw]).tobytes())
        for k in (0, 1):
            pl = o[W * H + k * W * H // 4 : W * H + (k + 1) * W * H // 4].reshape(H // 2, W // 2)
            h.update(np.ascontiguousarray(pl[cy // 2 : cy // 2 + ch // 2, cx // 2 : cx // 2 + cw // 2]).tobytes())
    return h.hexdigest(), n_out


# the long streams are left to the GPU suite to keep the CPU suite within a few minutes
CPU_SET = [s for s in decodable_streams() if s not in ("LS_SVA_D.264", "BA1_FT_C.264", "CI1_FT_B.264", "syn1080p_cbp_ippp_16f.264", "CVPCMNL1_SVA_C.264")]


@pytest.mark.parametrize("name", CPU_SET)
def test_host_parse_plus_oracle_matches_reference_sha1(name, tmp_path):
    sha, n = decode_with_oracle(os.path.join(STREAM_DIR, name), str(tmp_path))
    assert sha == SHA1[name]["sha1"], f"{name}: {n} pictures decoded, SHA1 differs from the reference"


def test_cavlc_tables_are_prefix_codes(tmp_path):
    # the tool refuses to run when the Kraft/prefix self-check of its VLC tables fails (exit code 3)
    r = subprocess.run([TOOL, os.path.join(STREAM_DIR, "SVA_BA2_D.264"), os.path.join(str(tmp_path), "x.lhfb")], capture_output=True)
    assert r.returncode == 0


def test_damaged_stream_freezes_output_like_the_reference(tmp_path):
    """Error_I_P.264: three sequences of two resolutions, five of its six pictures have lost slices.  The reference (freeze error
    concealment modes) hands out exactly one picture, the complete 640x480 IDR; so does the host decoder."""
    sha, n = decode_with_oracle(os.path.join(STREAM_DIR, "Error_I_P.264"), str(tmp_path))
    assert n == 1 and sha == SHA1["Error_I_P.264"]["sha1"]


def test_stream_the_reference_refuses_is_refused_with_the_same_default_stream(tmp_path):
    """test_scalinglist_jm.264: every PPS carries weighted_bipred_idc != 0, which the reference's ParsePps rejects, so no slice finds
    its parameter set: the reference hands out 0 frames and writes a `.pip` holding only the default stream.
    The host decoder must refuse the same way: 0 pictures, no side streams, byte-identical default stream."""
    import json

    from streams import REFUSED

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pip_tag_sha1.json")))
    for name in sorted(REFUSED):
        d = str(tmp_path)
        pip = os.path.join(d, "o.pip")
        r = subprocess.run([TOOL, os.path.join(STREAM_DIR, name), os.path.join(d, "o.lhfb"), "1", os.path.join(d, "o.lhmd"), pip], capture_output=True, text=True)
        assert re.search(r"pictures 0 ", r.stderr), r.stderr[-300:]
        assert os.path.getsize(os.path.join(d, "o.lhfb")) == 0 and os.path.getsize(os.path.join(d, "o.lhmd")) == 0
        assert SHA1[name]["frames"] == 0 and gold[name]["tags"] == {}
        assert hashlib.sha1(open(pip, "rb").read()).hexdigest() == gold[name]["default"]["sha1"], f"{name}: default stream differs from the reference's"
