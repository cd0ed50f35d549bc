"""GPU: the multi-stream runtime through its C entry points (include/lh264_runtime.h, ctypes binding losslessh264_b200/runtime.py):
bitstream bytes in host memory -> pictures handed to a callback + `.pip` bytes read back, against the reference's goldens."""
import hashlib
import json
import os

import pytest

from streams import SHA1, STREAM_DIR

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pip_tag_sha1.json")))


def test_runtime_api_pictures_and_pip_match_reference():
    from losslessh264_b200.runtime import Runtime

    names = ["tibby.264", "SVA_BA1_B.264", "BA_MW_D_IDR_LOST.264", "Error_I_P.264", "syn_cif_cabac_4slices_ltr_6f.264", "tibby8x8cavlc.264"]
    rt = Runtime(threads=4, pipelines=2, depth=3)
    for n in names * 2:
        rt.add(open(os.path.join(STREAM_DIR, n), "rb").read())
    hashes = [hashlib.sha1() for _ in range(rt.n)]
    counts = [0] * rt.n

    def on_picture(stream, index, w, h, data):
        assert index == counts[stream]
        counts[stream] += 1
        hashes[stream].update(data)

    rt.run(on_picture)
    for i in range(rt.n):
        n = names[i % len(names)]
        r = rt.result(i)
        assert hashes[i].hexdigest() == SHA1[n]["sha1"], f"{n}: pictures handed to the callback differ from the reference"
        assert r["yuv_sha1"] == SHA1[n]["sha1"] and r["pictures_output"] == counts[i]
        if "frames" in SHA1[n]:
            assert counts[i] == SHA1[n]["frames"]
        files = rt.pip(i)
        g = GOLD[n]
        assert hashlib.sha1(files[0]).hexdigest() == g["default"]["sha1"], f"{n}: default stream differs"
        assert sorted(str(t) for t in files if t) == sorted(g["tags"]), f"{n}: tag set"
        for t, b in files.items():
            if t:
                assert hashlib.sha1(b).hexdigest() == g["tags"][str(t)]["sha1"], f"{n}: .pip.{t} differs"
    st = rt.stats()
    assert st["pictures"] >= sum(counts) and st["kernel_launches"] > 0 and st["h2d_bytes"] > 0 and st["d2h_bytes"] > 0
    rt.close()


def test_runtime_model_only_leg():
    """x.264 -> x.pip without a YUV sink: no reconstruction kernels, no picture comes back, same `.pip` bytes."""
    from losslessh264_b200.runtime import Runtime

    n = "syn720p_8f.264"
    rt = Runtime(threads=2, want_yuv=False, want_hash=False)
    for _ in range(4):
        rt.add(open(os.path.join(STREAM_DIR, n), "rb").read())
    rt.run()
    g = GOLD[n]
    for i in range(4):
        files = rt.pip(i)
        assert hashlib.sha1(files[0]).hexdigest() == g["default"]["sha1"]
        for t, b in files.items():
            if t:
                assert hashlib.sha1(b).hexdigest() == g["tags"][str(t)]["sha1"], f".pip.{t} differs"
    assert rt.stats()["d2h_bytes"] == 0
    rt.close()
