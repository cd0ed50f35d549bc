"""Generates tests/golden/pip_tag_sha1.json: for every stream of tests/golden/streams, the SHA1 and length of
each `.pip.<tag>` side stream (and of the default `.pip` stream) written by the compiled, UNMODIFIED reference
(oracle/_ref/h264dec in.264 out.pip).  Run in the build container (needs oracle/_ref, ~9.5 GB RAM per run).

    python tests/golden/make_pip_golden.py [stream ...]
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
STREAMS = os.path.join(HERE, "streams")
OUT = os.path.join(HERE, "pip_tag_sha1.json")


def main():
    names = sys.argv[1:] or sorted(os.listdir(STREAMS))
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        with tempfile.TemporaryDirectory() as d:
            pip = os.path.join(d, "o.pip")
            r = subprocess.run([REF, os.path.join(STREAMS, name), pip], capture_output=True)
            if r.returncode != 0 or not os.path.exists(pip):
                table[name] = {"error": f"reference exit code {r.returncode}"}
                continue
            entry = {"default": {"sha1": hashlib.sha1(open(pip, "rb").read()).hexdigest(), "len": os.path.getsize(pip)}, "tags": {}}
            for f in sorted(os.listdir(d)):
                if f.startswith("o.pip."):
                    b = open(os.path.join(d, f), "rb").read()
                    entry["tags"][f[6:]] = {"sha1": hashlib.sha1(b).hexdigest(), "len": len(b)}
            table[name] = entry
            print(name, entry["default"]["len"], sum(v["len"] for v in entry["tags"].values()), flush=True)
    json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
