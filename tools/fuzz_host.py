#!/usr/bin/env python
"""Robustness fuzzing of the host side under ASAN/UBSAN (development tool, not part of the product or the tests).

  python tools/fuzz_host.py parse  [seed] [cases per stream]   corrupted .264 files through lh264_parse_dump (decode and model mode)
  python tools/fuzz_host.py expand [seed] [cases per stream]   corrupted .pip file sets through the CPU build of the decompress core

Both binaries are compiled with -fsanitize=address,undefined into a temporary directory; a crash, a sanitizer report or a hang
(time-out) is printed with the case kept next to it.  End of round 1: 3 600 parse cases and 900 expand cases, none left.
"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "losslessh264_b200", "host")
STREAMS = os.path.join(ROOT, "tests", "golden", "streams")
SAN = ["-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-w"]
ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:exitcode=99", UBSAN_OPTIONS="halt_on_error=1:exitcode=98")
CORE = ["decoder.cpp", "h264_syntax.cpp", "cavlc.cpp", "cavlc_writer.cpp", "pip_container.cpp", "pip_expand.cpp"]


def corrupt(data, rng):
    b = bytearray(data)
    for _ in range(rng.choice([1, 2, 5, 20, 100])):
        i = rng.randrange(len(b))
        b[i] = b[i] ^ (1 << rng.randrange(8)) if rng.random() < 0.5 else rng.randrange(256)
    if rng.random() < 0.2:
        b = b[: rng.randrange(len(b))]
    return bytes(b)


def run(cmd, timeout):
    try:
        return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=ENV).returncode
    except subprocess.TimeoutExpired:
        return "TIMEOUT"


def fuzz_parse(seed, n, tmp):
    exe = os.path.join(tmp, "parse_asan")
    srcs = [os.path.join(HOST, f) for f in ["tool_parse_dump.cpp", "sinks_file.cpp"] + CORE if os.path.exists(os.path.join(HOST, f))]
    subprocess.check_call(["g++"] + SAN + ["-o", exe] + srcs)
    rng = random.Random(seed)
    names = ["test_qcif_cabac.264", "tibbycabac.264", "syn_cif_cabac_4slices_ltr_6f.264", "QCIF_2P_I_allIPCM.264", "tibby8x8cavlc.264", "SVA_FM1_E.264",
             "MR2_TANDBERG_E.264", "CVPCMNL1_SVA_C.264", "Error_I_P.264"]
    bad = 0
    for name in names:
        data = open(os.path.join(STREAMS, name), "rb").read()
        for it in range(n):
            case = os.path.join(tmp, "case.264")
            open(case, "wb").write(corrupt(data, rng))
            for extra in ([], ["1", os.path.join(tmp, "c.lhmd"), os.path.join(tmp, "c.pip")]):
                rc = run([exe, case, os.path.join(tmp, "c.lhfb")] + extra, 30)
                if rc not in (0, 1, 2, 3, 4, 5):
                    bad += 1
                    keep = os.path.join(ROOT, f"fuzz_parse_{seed}_{bad}.264")
                    os.replace(case, keep)
                    print(name, it, "rc", rc, "->", keep)
                    break
    print("parse fuzz done, bad =", bad)
    return bad


def fuzz_expand(seed, n, tmp):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    from test_pip_expand import write_pip_cpu

    exe = os.path.join(tmp, "expand_asan")
    subprocess.check_call(["g++"] + SAN + ["-o", exe, os.path.join(ROOT, "oracle", "model_decode_cpu.cpp")] +
                          [os.path.join(HOST, f) for f in ("pip_expand.cpp", "pip_container.cpp", "h264_syntax.cpp")])
    subprocess.check_call(["make", "-s", "-C", HOST, "../_lib/lh264_parse_dump"])
    rng = random.Random(seed)
    bad = 0
    for name in ["test_qcif_cabac.264", "syn_cif_cabac_4slices_ltr_6f.264", "tibby8x8cavlc.264", "SVA_FM1_E.264", "Error_I_P.264", "MR2_TANDBERG_E.264"]:
        d = os.path.join(tmp, name + ".d")
        os.makedirs(d)
        pip = write_pip_cpu(name, d)
        files = ["x.pip"] + [f for f in os.listdir(d) if f.startswith("x.pip.")]
        orig = {f: open(os.path.join(d, f), "rb").read() for f in files}
        for it in range(n):
            f = rng.choice(files) if rng.random() < 0.7 else "x.pip"
            if not orig[f]:
                continue
            open(os.path.join(d, f), "wb").write(corrupt(orig[f], rng))
            rc = run([exe, pip, os.path.join(d, "o.264")], 60)
            if rc not in (0, 1, 2, 3, 4, 5):
                bad += 1
                print(name, it, f, "rc", rc)
            open(os.path.join(d, f), "wb").write(orig[f])
    print("expand fuzz done, bad =", bad)
    return bad


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "parse"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with tempfile.TemporaryDirectory() as tmp:
        sys.exit(1 if (fuzz_parse if mode == "parse" else fuzz_expand)(seed, n, tmp) else 0)
