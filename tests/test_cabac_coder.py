"""CPU: the CABAC arithmetic coder pair — encoder of the decompress core (csrc/model_decode_core.inl, the code k_model_decode runs,
compiled for the CPU) against the host parser's decoding engine (host/cabac.h) — on random and adversarial bin sequences
(oracle/cabac_selftest.cpp; the analogue of the reference's bool-coder self-test, console/dec/src/test_arithmetic_coding.cpp:18-71)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabac_encoder_and_decoder_round_trip(tmp_path):
    exe = os.path.join(str(tmp_path), "cabac_selftest")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", exe, os.path.join(ROOT, "oracle", "cabac_selftest.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "240 cases" in r.stdout
