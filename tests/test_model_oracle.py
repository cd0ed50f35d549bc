"""CPU: the recompression-model oracle (oracle/model_oracle.c) fed by the host decoder's model hand-off
records, pinned tag by tag against the `.pip.<tag>` side streams written by the compiled, unmodified
reference (tests/golden/pip_tag_sha1.json, made by tests/golden/make_pip_golden.py)."""
import hashlib
import json
import os
import subprocess

import pytest

from losslessh264_b200.records import read_lhmd
from oracle import oracle
from streams import CABAC, LIB_DIR, STREAM_DIR, decodable_streams

TOOL = os.path.join(LIB_DIR, "lh264_parse_dump")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pip_tag_sha1.json")))

# long streams are left to the GPU suite
CPU_SET = [s for s in decodable_streams() if s in GOLD and "tags" in GOLD[s] and s not in ("LS_SVA_D.264", "BA1_FT_C.264", "CI1_FT_B.264", "CVPCMNL1_SVA_C.264")]


@pytest.fixture(scope="module", autouse=True)
def _built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(LIB_DIR), "host"), "../_lib/lh264_parse_dump"])


def model_pictures(name, tmp_path):
    out = os.path.join(tmp_path, "d.lhfb")
    md = os.path.join(tmp_path, "d.lhmd")
    r = subprocess.run([TOOL, os.path.join(STREAM_DIR, name), out, "1", md], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return read_lhmd(md)


def check_tags(name, got):
    want = GOLD[name]["tags"]
    assert sorted(str(t) for t in got) == sorted(want), f"{name}: tag set differs: {sorted(got)} vs {sorted(want)}"
    for t, b in got.items():
        w = want[str(t)]
        assert len(b) == w["len"] and hashlib.sha1(b).hexdigest() == w["sha1"], f"{name}: side stream .pip.{t} differs from the reference ({len(b)} vs {w['len']} bytes)"


@pytest.mark.parametrize("name", CPU_SET)
def test_model_oracle_matches_reference_pip_tags(name, tmp_path):
    mo = oracle.ModelOracle()
    for p in model_pictures(name, str(tmp_path)):
        mo.picture(p)
    check_tags(name, mo.finish())
    mo.close()
