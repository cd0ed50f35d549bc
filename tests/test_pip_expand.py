"""CPU: decompress direction — host container expansion (losslessh264_b200/host/pip_expand.cpp) around the CPU
build of the decompress core that the CUDA kernel k_model_decode runs (oracle/model_decode_cpu.cpp compiles
csrc/model_decode_core.inl with MD_FN = static inline).  For every CAVLC stream of the corpus: write the `.pip` file set
(default stream from the host container code, side streams from the compress oracle — both pinned to the reference's bytes by
tests/golden/pip_tag_sha1.json), expand it, and require the ORIGINAL .264 back, byte for byte."""
import os
import subprocess

import pytest

from oracle import oracle
from streams import FMO, LIB_DIR, PCM_SYNTH, STREAM_DIR
from test_model_oracle import CPU_SET as MODEL_SET, TOOL

# CAVLC and CABAC streams alike (the core re-serialises both); I_PCM macroblocks are not in the model, as in the reference
CPU_SET = [s for s in MODEL_SET if s != "QCIF_2P_I_allIPCM.264" and s not in PCM_SYNTH | FMO]
from losslessh264_b200.records import read_lhmd


@pytest.fixture(scope="module")
def expand_cpu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(LIB_DIR), "host"), "../_lib/lh264_parse_dump"])
    return oracle.build_expand_checker()


def write_pip_cpu(name, d):
    """x.pip + x.pip.<tag> made on the CPU: host parser + container code + compress oracle."""
    pip = os.path.join(d, "x.pip")
    md = os.path.join(d, "d.lhmd")
    r = subprocess.run([TOOL, os.path.join(STREAM_DIR, name), os.path.join(d, "d.lhfb"), "1", md, pip], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mo = oracle.ModelOracle()
    for p in read_lhmd(md):
        mo.picture(p)
    for t, b in mo.finish().items():
        open(f"{pip}.{t}", "wb").write(b)
    mo.close()
    return pip


@pytest.mark.parametrize("name", CPU_SET)
def test_expanded_pip_is_the_original_file(name, tmp_path, expand_cpu):
    d = str(tmp_path)
    pip = write_pip_cpu(name, d)
    out = os.path.join(d, "out.264")
    r = subprocess.run([expand_cpu, pip, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(STREAM_DIR, name), "rb").read(), f"{name}: expanded file differs from the original"


@pytest.mark.parametrize("name", ["SVA_BA1_B.264", "test_qcif_cabac.264"])
def test_truncated_side_stream_is_not_silently_accepted(name, tmp_path, expand_cpu):
    d = str(tmp_path)
    pip = write_pip_cpu(name, d)
    big = max((f for f in os.listdir(d) if f.startswith("x.pip.")), key=lambda f: os.path.getsize(os.path.join(d, f)))
    b = open(os.path.join(d, big), "rb").read()
    open(os.path.join(d, big), "wb").write(b[: len(b) // 2])
    out = os.path.join(d, "out.264")
    r = subprocess.run([expand_cpu, pip, out], capture_output=True, text=True)
    same = r.returncode == 0 and open(out, "rb").read() == open(os.path.join(STREAM_DIR, name), "rb").read()
    assert not same


def test_slice_groups_are_refused_not_mangled(tmp_path, expand_cpu):
    """The decompress direction walks macroblocks in raster order; a stream with slice groups is refused (the reference's own
    decompressor aborts on it).  Arbitrary slice order alone is fine: aso_4slices_nopcm is in CPU_SET."""
    d = str(tmp_path)
    pip = write_pip_cpu("fmo_dispersed3_2slices_aso_nopcm.264", d)
    r = subprocess.run([expand_cpu, pip, os.path.join(d, "out.264")], capture_output=True, text=True)
    assert r.returncode != 0 and "slice groups" in r.stderr
    assert "aso_4slices_nopcm.264" in CPU_SET


def test_cabac_pcm_stream_is_refused_like_the_reference(tmp_path, expand_cpu):
    """I_PCM in a CABAC stream: the reference aborts in computeNeighborPriorsCabac (decode_slice.cpp:1754); here error 21."""
    d = str(tmp_path)
    pip = write_pip_cpu("QCIF_2P_I_allIPCM.264", d)
    r = subprocess.run([expand_cpu, pip, os.path.join(d, "o.264")], capture_output=True, text=True)
    assert r.returncode != 0
