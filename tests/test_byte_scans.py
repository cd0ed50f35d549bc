"""CPU: the memchr-based byte scans of the host parser (start codes B.1, emulation prevention 7.4.1, the unit loop of
h264dec.cpp:246-364) against byte-by-byte restatements, on random buffers dense in zeros and 00 00 0x patterns
(tests/host/byte_scans_check.cpp).  The corpus streams rarely hold runs of zeros; damaged streams do."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "losslessh264_b200", "host")


def test_scans_agree_with_the_plain_walks(tmp_path):
    exe = os.path.join(str(tmp_path), "byte_scans_check")
    src = [os.path.join(ROOT, "tests", "host", "byte_scans_check.cpp")] + [os.path.join(HOST, f) for f in
           ("h264_syntax.cpp", "pip_container.cpp", "decoder.cpp", "cavlc.cpp", "cavlc_writer.cpp", "pip_expand.cpp")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-w", "-fsanitize=address,undefined", "-I" + HOST, "-o", exe] + src)
    r = subprocess.run([exe, "3000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-400:] + r.stderr[-800:]
