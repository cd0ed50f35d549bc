"""Adds streams to tests/golden/stream_sha1.json: SHA1 of the I420 output of the compiled, UNMODIFIED reference
(oracle/_ref/h264dec in.264 out.pip out.yuv), then their `.pip` goldens through make_pip_golden.py.  Build container only.

    python tests/golden/add_stream_golden.py name.264 [...]
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "h264dec")
OUT = os.path.join(HERE, "stream_sha1.json")

table = json.load(open(OUT))
for name in sys.argv[1:]:
    src = os.path.join(HERE, "streams", name)
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([REF, src, os.path.join(d, "o.pip"), os.path.join(d, "o.yuv")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-400:]
        yuv = open(os.path.join(d, "o.yuv"), "rb").read()
        pip = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.startswith("o.pip"))
    table[name] = {"bytes": os.path.getsize(src), "in_decoder_test": False, "pip_bytes": pip, "sha1": hashlib.sha1(yuv).hexdigest(), "synthetic": True, "yuv_bytes": len(yuv)}
    print(name, table[name])
json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
subprocess.check_call([sys.executable, os.path.join(HERE, "make_pip_golden.py")] + sys.argv[1:])
