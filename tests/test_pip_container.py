"""CPU: the `.pip` default stream ("boilerplate": start codes, NAL headers, parameter sets, slice HEADER bits, stop bits,
trailing zeros) written by the host container code (losslessh264_b200/host/pip_container.cpp) is byte-identical to the
one the compiled reference writes (tests/golden/pip_tag_sha1.json: "default"), for every CAVLC stream of the corpus —
including the damaged ones (lost slices, errors), whose boilerplate does not depend on reconstruction."""
import hashlib
import json
import os
import subprocess

import pytest

from streams import LIB_DIR, SHA1, STREAM_DIR

TOOL = os.path.join(LIB_DIR, "lh264_parse_dump")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pip_tag_sha1.json")))
NAMES = sorted(n for n in GOLD if "default" in GOLD[n])


@pytest.fixture(scope="module", autouse=True)
def _built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(LIB_DIR), "host"), "../_lib/lh264_parse_dump"])


@pytest.mark.parametrize("name", NAMES)
def test_default_stream_matches_reference(name, tmp_path):
    out = os.path.join(str(tmp_path), "o.pip")
    subprocess.run([TOOL, os.path.join(STREAM_DIR, name), os.path.join(str(tmp_path), "o.lhfb"), "1", os.path.join(str(tmp_path), "o.lhmd"), out], capture_output=True)
    b = open(out, "rb").read()
    want = GOLD[name]["default"]
    assert len(b) == want["len"] and hashlib.sha1(b).hexdigest() == want["sha1"], f"{name}: default .pip stream differs from the reference"


def test_corpus_is_covered():
    assert len(NAMES) >= 40
