"""GPU: whole streams through the reference-facing API (libopenh264_lh264.so: WelsCreateDecoder /
DecodeFrameNoDelay, driven by the h264dec-compatible CLI) — SHA1 of the decoded YUV must equal the
reference's (BASELINE.json config 2, the conformance batch; config 1/4 streams included)."""
import hashlib
import os
import subprocess

import pytest

from streams import LIB_DIR, SHA1, STREAM_DIR, cavlc_streams

pytestmark = pytest.mark.gpu
CLI = os.path.join(LIB_DIR, "h264dec_lh264")


@pytest.mark.parametrize("name", cavlc_streams())
def test_decoded_yuv_sha1_matches_reference(name, tmp_path):
    out = os.path.join(str(tmp_path), "o.yuv")
    r = subprocess.run([CLI, os.path.join(STREAM_DIR, name), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    sha = hashlib.sha1(open(out, "rb").read()).hexdigest()
    assert sha == SHA1[name]["sha1"], f"{name}: SHA1 {sha} != reference {SHA1[name]['sha1']}\n{r.stderr[-400:]}"
