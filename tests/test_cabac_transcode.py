"""CPU: CABAC paths that no stream of the corpus reaches (8x8-transform residual blocks, transform_size_8x8_flag contexts), checked by
TRANSCODING: the first (intra) picture of a CAVLC stream that uses the 8x8 transform is re-serialised with the CABAC slice-data
writer of the decompress core (oracle/expand_cpu --intra-as-cabac; an I slice header is the same for both entropy coders), the PPS
gets entropy_coding_mode_flag = 1, and the result must decode — with this repository's host CABAC parser, and with the compiled,
unmodified reference where it is present — to exactly the picture the original stream starts with."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

from losslessh264_b200.records import n_slots_needed, read_lhfb, read_lhmd
from oracle import oracle
from streams import LIB_DIR, ROOT, STREAM_DIR
from test_pip_expand import expand_cpu  # noqa: F401  (fixture)

TOOL = os.path.join(LIB_DIR, "lh264_parse_dump")
REF_DEC = os.path.join(ROOT, "oracle", "_ref", "h264dec")


def first_picture(stream, tmp, tag):
    """I420 bytes of the first decoded picture: host parse -> oracle reconstruction."""
    out = os.path.join(tmp, tag + ".lhfb")
    r = subprocess.run([TOOL, stream, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pics = read_lhfb(out)
    st = oracle.OracleStream(pics[0].mb_w, pics[0].mb_h, n_slots_needed(pics))
    return np.asarray(st.recon(pics[0])).tobytes(), pics[0]


def set_pps_cabac(nal):
    """entropy_coding_mode_flag := 1 in a PPS NAL unit (pic_parameter_set_id ue, seq_parameter_set_id ue, then the flag)."""
    b = bytearray(nal)
    pos = 8                                            # bit position behind the NAL header byte
    for _ in range(2):                                 # two ue(v)
        z = 0
        while not (b[pos >> 3] >> (7 - (pos & 7))) & 1:
            z += 1
            pos += 1
        pos += z + 1
    b[pos >> 3] |= 1 << (7 - (pos & 7))
    return bytes(b)


@pytest.mark.parametrize("name", ["tibby8x8cavlc.264", "syn_qcif_4slices_8f.264"])
def test_intra_picture_transcoded_to_cabac_decodes_to_the_same_samples(name, tmp_path, expand_cpu):
    d = str(tmp_path)
    src = os.path.join(STREAM_DIR, name)
    want, pic0 = first_picture(src, d, "orig")
    # `.pip` of the first picture only: default stream cut behind its last slice, side streams from the model oracle with the
    # CABAC convention for the PADBYTE stream (seven bits per slice, here zeros)
    pip, md = os.path.join(d, "x.pip"), os.path.join(d, "d.lhmd")
    r = subprocess.run([TOOL, src, os.path.join(d, "d.lhfb"), "1", md, pip], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p0 = read_lhmd(md, max_pictures=1)[0]
    n_slices = len(p0.slices)
    p0.slices["n_pad_bits"] = 7
    p0.slices["pad_bits"] = 0
    mo = oracle.ModelOracle()
    mo.picture(p0)
    for t, b in mo.finish().items():
        open(f"{pip}.{t}", "wb").write(b)
    mo.close()
    default = open(pip, "rb").read()
    starts = [m.start() for m in re.finditer(b"\x00\x00\x01", default)]
    vcl, cut = 0, None
    for k, at in enumerate(starts):                    # cut behind the picture's last slice unit (parameter sets and SEI stay)
        if vcl == n_slices:
            cut = at - (1 if at and default[at - 1] == 0 else 0)
            break
        if default[at + 3] & 31 in (1, 5):
            vcl += 1
    assert cut is not None
    open(pip, "wb").write(default[:cut])
    out = os.path.join(d, "cabac.264")
    r = subprocess.run([expand_cpu, pip, out, "--intra-as-cabac"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # flip the PPS flag
    data = open(out, "rb").read()
    idx = [m.start() for m in re.finditer(b"\x00\x00\x01", data)] + [len(data)]
    nals = [data[idx[i] + 3 : idx[i + 1]] for i in range(len(idx) - 1)]
    k = [n[0] & 31 for n in nals].index(8)
    data = data[: idx[k] + 3] + set_pps_cabac(nals[k]) + data[idx[k + 1]:]
    open(out, "wb").write(data)
    got, gpic = first_picture(out, d, "cabac")
    assert got == want, f"{name}: the CABAC transcode of the first picture decodes to different samples"
    if "8x8" in name:                                  # the point of the exercise: 8x8-transform macroblocks went through CABAC
        assert int((gpic.mbs["flags"] & 1).sum()) > 0
    if os.path.exists(REF_DEC):                        # the unmodified reference's CABAC parser agrees
        yuv = os.path.join(d, "ref.yuv")
        r = subprocess.run([REF_DEC, out, os.path.join(d, "r.pip"), yuv], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-300:]
        ref = open(yuv, "rb").read()
        sha_want = hashlib.sha1(open_first_frame_cropped(src, d)).hexdigest()
        assert hashlib.sha1(ref).hexdigest() == sha_want, f"{name}: the reference decodes the CABAC transcode to different samples"


def open_first_frame_cropped(src, d):
    """First frame of the reference's own output for the original stream."""
    yuv = os.path.join(d, "orig.yuv")
    r = subprocess.run([REF_DEC, src, os.path.join(d, "o.pip"), yuv], capture_output=True, text=True)
    assert r.returncode == 0
    m = re.search(r"iWidth:\s+(\d+)\s+height:\s+(\d+)", r.stderr)
    w, h = int(m.group(1)), int(m.group(2))
    return open(yuv, "rb").read()[: w * h * 3 // 2]
