#!/usr/bin/env python
"""Makes the bench corpus of BASELINE.json config 4 ("synthetic 1920x1080 Constrained Baseline IPPP streams, 1024 independent
streams"): EIGHT distinct 16-picture 1080p streams (same generator and encoder flags as tests/golden/streams/
syn1080p_cbp_ippp_16f.264, different seeds -> different textures, motion vectors, coefficient counts), encoded with the reference
encoder (oracle/_ref/h264enc), and pins each with what the UNMODIFIED reference decoder makes of it (oracle/_ref/h264dec): YUV
SHA1, frame count, SHA1 over the `.pip` file set.  Runs where /root/reference exists; the streams and bench_streams.json are
committed so that bench.py can run (and check its own output) on the GPU box.

    python tools/make_bench_streams.py        # writes tests/golden/bench_streams/
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_synth_stream import encode  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_streams")
REF_DEC = os.path.join(ROOT, "oracle", "_ref", "h264dec")
SEEDS = [1234, 2001, 2002, 2003, 2004, 2005, 2006, 2007]       # 1234 = the stream the round-1 bench replicated


def pip_sha1(prefix):
    """SHA1 over default stream, then (tag byte, tag stream) for every tag file in ascending tag order (= lh264_rt's pip_sha1)."""
    h = hashlib.sha1(open(prefix, "rb").read())
    d = os.path.dirname(prefix)
    tags = sorted(int(f.rsplit(".", 1)[1]) for f in os.listdir(d) if f.startswith(os.path.basename(prefix) + "."))
    total = os.path.getsize(prefix)
    for t in tags:
        b = open(f"{prefix}.{t}", "rb").read()
        h.update(bytes([t]))
        h.update(b)
        total += len(b)
    return h.hexdigest(), total


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = {}
    for seed in SEEDS:
        name = f"syn1080p_seed{seed}_16f.264"
        path = os.path.join(OUT, name)
        size = encode(1920, 1080, 16, seed, path)
        with tempfile.TemporaryDirectory() as td:
            r = subprocess.run([REF_DEC, path, os.path.join(td, "o.pip"), os.path.join(td, "o.yuv")], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-300:]
            yuv = open(os.path.join(td, "o.yuv"), "rb").read()
            ps, pb = pip_sha1(os.path.join(td, "o.pip"))
        meta[name] = {"seed": seed, "bytes": size, "frames": len(yuv) // (1920 * 1080 * 3 // 2), "yuv_sha1": hashlib.sha1(yuv).hexdigest(), "pip_sha1": ps, "pip_bytes": pb}
        print(name, meta[name], flush=True)
    json.dump(meta, open(os.path.join(OUT, "bench_streams.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
