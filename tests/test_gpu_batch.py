"""GPU: the multi-stream host runtime (lh264_batch: parse threads -> pinned staging -> one recon + one model submit per
step for all streams) gives, for every stream of a mixed batch, the reference's YUV SHA1 and the reference's `.pip` file set."""
import hashlib
import json
import os
import re
import subprocess

import pytest

from streams import LIB_DIR, SHA1, STREAM_DIR

pytestmark = pytest.mark.gpu
TOOL = os.path.join(LIB_DIR, "lh264_batch")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pip_tag_sha1.json")))


def run(names, copies, threads, tmp_path, extra=()):
    d = os.path.join(str(tmp_path), "pip")
    os.makedirs(d, exist_ok=True)
    r = subprocess.run([TOOL, "--copies", str(copies), "--threads", str(threads), "--pip-dir", d, *extra] + [os.path.join(STREAM_DIR, n) for n in names],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-400:] + r.stderr[-400:]
    lines = r.stdout.strip().splitlines()
    summary = json.loads(lines[-1])
    per = {}
    for ln in lines[:-1]:
        m = re.match(r"(\S+) pictures (\d+) yuv_sha1 (\S+) pip_sha1 (\S+)", ln)
        per[os.path.basename(m.group(1))] = (int(m.group(2)), m.group(3), m.group(4))
    return summary, per, d


def check_pip_files(d, idx, name):
    g = GOLD[name]
    pre = f"s{idx}.pip"
    assert hashlib.sha1(open(os.path.join(d, pre), "rb").read()).hexdigest() == g["default"]["sha1"], f"{name}: default stream differs"
    tags = sorted(f[len(pre) + 1:] for f in os.listdir(d) if f.startswith(pre + "."))
    assert tags == sorted(g["tags"]), f"{name}: tag set {tags}"
    for t in tags:
        assert hashlib.sha1(open(os.path.join(d, f"{pre}.{t}"), "rb").read()).hexdigest() == g["tags"][t]["sha1"], f"{name}: .pip.{t} differs"


def test_mixed_batch_of_streams_matches_reference(tmp_path):
    names = ["tibby.264", "BA1_FT_C.264", "black.264", "syn_qcif_4slices_8f.264", "BA_MW_D_P_LOST.264", "CVFC1_Sony_C.jsv", "tibby8x8cavlc.264"]
    summary, per, d = run(names, 3, 4, tmp_path)
    assert summary["copy_mismatches"] == 0 and summary["streams"] == 3 * len(names)
    for i, n in enumerate(names):
        pics, yuv, _ = per[n]
        assert "frames" not in SHA1[n] or pics == SHA1[n]["frames"], f"{n}: {pics} pictures"
        assert yuv == SHA1[n]["sha1"], f"{n}: YUV SHA1 differs from the reference"
        for c in range(3):                                  # every copy wrote the reference's file set
            check_pip_files(d, c * len(names) + i, n)


def test_1080p_streams_in_one_batch(tmp_path):
    n = "syn1080p_cbp_ippp_16f.264"
    summary, per, d = run([n], 12, 6, tmp_path)
    assert summary["copy_mismatches"] == 0 and summary["pictures"] == 12 * 16
    assert per[n][1] == SHA1[n]["sha1"]
    check_pip_files(d, 0, n)
    check_pip_files(d, 11, n)
