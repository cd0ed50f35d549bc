"""GPU: whole streams through the reference-facing API (libopenh264_lh264.so: WelsCreateDecoder /
DecodeFrameNoDelay, driven by the h264dec-compatible CLI) — SHA1 of the decoded YUV must equal the
reference's (BASELINE.json config 2, the conformance batch; config 1/4 streams included)."""
import hashlib
import os
import subprocess

import pytest

from streams import LIB_DIR, SHA1, STREAM_DIR, cavlc_streams, decodable_streams

pytestmark = pytest.mark.gpu
CLI = os.path.join(LIB_DIR, "h264dec_lh264")


@pytest.mark.parametrize("name", decodable_streams())
def test_decoded_yuv_sha1_matches_reference(name, tmp_path):
    out = os.path.join(str(tmp_path), "o.yuv")
    r = subprocess.run([CLI, os.path.join(STREAM_DIR, name), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    sha = hashlib.sha1(open(out, "rb").read()).hexdigest()
    assert sha == SHA1[name]["sha1"], f"{name}: SHA1 {sha} != reference {SHA1[name]['sha1']}\n{r.stderr[-400:]}"


# ---- compress mode of the CLI: `h264dec in.264 out.pip out.yuv` writes the same file set as the reference ------------------
import hashlib
import json

PIP_GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pip_tag_sha1.json")))
PIP_SET = ["tibby.264", "tibby8x8cavlc.264", "BA1_FT_C.264", "black.264", "BAMQ2_JVC_C.264", "MR2_TANDBERG_E.264", "SVA_FM1_E.264", "CVPCMNL1_SVA_C.264",
           "syn_qcif_4slices_8f.264", "syn720p_8f.264", "Error_I_P.264"]      # the last one: damaged, switches to a BIGGER picture mid-stream


@pytest.mark.parametrize("name", PIP_SET)
def test_cli_compress_writes_the_reference_pip_file_set(name, tmp_path):
    pip = os.path.join(str(tmp_path), "o.pip")
    yuv = os.path.join(str(tmp_path), "o.yuv")
    r = subprocess.run([os.path.join(LIB_DIR, "h264dec_lh264"), os.path.join(STREAM_DIR, name), pip, yuv], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    g = PIP_GOLD[name]
    b = open(pip, "rb").read()
    assert hashlib.sha1(b).hexdigest() == g["default"]["sha1"], f"{name}: default stream differs"
    files = sorted(f[6:] for f in os.listdir(str(tmp_path)) if f.startswith("o.pip."))
    assert files == sorted(g["tags"]), f"{name}: tag file set {files} != {sorted(g['tags'])}"
    for t in files:
        tb = open(pip + "." + t, "rb").read()
        assert hashlib.sha1(tb).hexdigest() == g["tags"][t]["sha1"], f"{name}: .pip.{t} differs from the reference"
    assert hashlib.sha1(open(yuv, "rb").read()).hexdigest() == SHA1[name]["sha1"], f"{name}: YUV differs"


# ---- lossless round trip (BASELINE config 1, roundtriptest/run.sh:27-39): the `.pip` file set written by the B200 path is fed to the
# UNMODIFIED reference decompressor (oracle/_ref/h264dec x.pip x.264, test infrastructure) and must give back the original
# file byte for byte.  Streams on which the reference itself does not round-trip are not in the list.
REF_DEC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "h264dec")
ROUNDTRIP_SET = ["tibby.264", "black.264", "BAMQ2_JVC_C.264", "BA1_FT_C.264", "tibby8x8cavlc.264", "BA_MW_D.264", "MR2_TANDBERG_E.264", "syn720p_8f.264", "tibbycabac.264"]


@pytest.mark.skipif(not os.path.exists(REF_DEC), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("name", ROUNDTRIP_SET)
def test_pip_written_on_the_gpu_round_trips_through_the_reference_decompressor(name, tmp_path):
    src = os.path.join(STREAM_DIR, name)
    pip = os.path.join(str(tmp_path), "o.pip")
    r = subprocess.run([os.path.join(LIB_DIR, "h264dec_lh264"), src, pip], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    back = os.path.join(str(tmp_path), "back.264")
    r = subprocess.run([REF_DEC, pip, back], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-400:]
    assert open(back, "rb").read() == open(src, "rb").read(), f"{name}: the reference decompressor does not reproduce the original from our .pip"
