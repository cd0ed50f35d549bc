"""CPU: decompress direction, stage A — the CAVLC slice_data() re-serialiser
(losslessh264_b200/host/cavlc_writer.cpp; replaces EncoderState::writeMbSynCavlc, decode_slice.cpp:1442-1835 ->
encoder/core/src/svc_set_mb_syn_cavlc.cpp) reproduces, from the model hand-off records alone, exactly the bits the parser
consumed: every slice of every CAVLC stream of the corpus is re-written and compared bit for bit."""
import os
import re
import subprocess

import pytest

from streams import CABAC, LIB_DIR, PCM_SYNTH, REFUSED, STREAM_DIR

TOOL = os.path.join(LIB_DIR, "lh264_parse_dump")
PCM = {"CVPCMNL1_SVA_C.264"} | PCM_SYNTH          # I_PCM samples are not part of the model records (the reference does not model them either)
NAMES = sorted(n for n in os.listdir(STREAM_DIR) if n.rsplit(".", 1)[-1] in ("264", "jsv", "h264") and n not in CABAC | PCM | REFUSED)


@pytest.fixture(scope="module", autouse=True)
def _built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(LIB_DIR), "host"), "../_lib/lh264_parse_dump"])


def _check(name):
    env = dict(os.environ, LH264_CHECK_REWRITE="1")
    r = subprocess.run([TOOL, os.path.join(STREAM_DIR, name), "/dev/null", "1", "/dev/null"], capture_output=True, text=True, env=env)
    m = re.search(r"rewrite check: (\d+) slices re-serialised, (\d+) mismatches", r.stderr)
    assert m, r.stderr[-400:]
    return int(m.group(1)), int(m.group(2))


@pytest.mark.parametrize("name", NAMES)
def test_rewritten_slice_data_is_bit_identical(name):
    slices, bad = _check(name)
    assert slices > 0 and bad == 0, f"{name}: {bad} of {slices} slices differ after re-serialisation"


def test_pcm_is_reported_not_silently_wrong():
    slices, bad = _check("CVPCMNL1_SVA_C.264")
    assert slices > 0 and bad > 0          # the writer refuses I_PCM records (write_mb returns false): counted as mismatches


def test_corpus_is_covered():
    assert len(NAMES) >= 40
