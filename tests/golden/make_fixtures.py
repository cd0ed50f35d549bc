#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the compiled reference (oracle/_ref/h264dec_dump).

Runs only where /root/reference exists (the build container).  Each fixture holds the first N
pictures of a reference test stream in hand-off form (include/lh264_b200.h records, converted
from the reference's SDqLayer by oracle/ref_dump_hook.cpp) together with the picture the
reference reconstructed for it.  Also records the reference's YUV SHA1 per stream in
tests/golden/ref_sha1.json (cross-checked against test/api/decoder_test.cpp:90-128).
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from losslessh264_b200.records import read_lhfb, save_fixture  # noqa: E402

REF = os.environ.get("LH264_REFERENCE", "/root/reference")
DUMP = os.path.join(ROOT, "oracle", "_ref", "h264dec_dump")

# name -> (stream, number of leading pictures kept)
FIXTURES = {
    "ba_mw_d": ("res/BA_MW_D.264", 8),
    "tibby8x8cavlc": ("roundtriptest/tibby8x8cavlc.264", 5),
    "ci1_ft_b": ("res/CI1_FT_B.264", 3),
    "mr2_tandberg_e": ("res/MR2_TANDBERG_E.264", 20),
    "allipcm": ("res/QCIF_2P_I_allIPCM.264", 2),
    "basqp1_sony_c": ("res/BASQP1_Sony_C.jsv", 4),
    "nlmq2_jvc_c": ("res/NLMQ2_JVC_C.264", 4),
    "mps_mw_a": ("res/MPS_MW_A.264", 6),
    "sva_cl1_e": ("res/SVA_CL1_E.264", 6),
    "cvfc1_sony_c": ("res/CVFC1_Sony_C.jsv", 3),
    "test_vd_rc": ("res/test_vd_rc.264", 3),
}


def main():
    sha = {}
    for name, (stream, n) in FIXTURES.items():
        with tempfile.TemporaryDirectory() as td:
            env = dict(os.environ, LH264_DUMP=os.path.join(td, "d.lhfb"), LH264_DUMP_PIX="1")
            subprocess.run([DUMP, os.path.join(REF, stream), os.path.join(td, "o.pip"), os.path.join(td, "o.yuv")], env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            pics = read_lhfb(os.path.join(td, "d.lhfb"), max_pictures=n)
            save_fixture(os.path.join(HERE, name + ".npz"), pics)
            sha[os.path.basename(stream)] = hashlib.sha1(open(os.path.join(td, "o.yuv"), "rb").read()).hexdigest()
            print(name, len(pics), "pictures", os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")
    json.dump(sha, open(os.path.join(HERE, "fixture_stream_sha1.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
