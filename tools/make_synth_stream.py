#!/usr/bin/env python
"""Makes the synthetic test streams of BASELINE.json configs 3-5 with the reference encoder
(oracle/_ref/h264enc, test tooling only; runs where /root/reference exists).

YUV generator: luma = 0.6*(((x+3t)//2 + (y+2t)//3) mod 256) +
0.4*tex[2t:, 3t:] with tex = 8x8-replicated uniform noise, + N(0,2) noise, clipped; chroma = slow
ramps.  Encoder flags as in SURVEY §8d.
"""
import argparse
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LH264_REFERENCE", "/root/reference")


def synth_yuv(w, h, frames, seed):
    rng = np.random.default_rng(seed)
    tex = np.kron(rng.integers(0, 256, ((h + 2 * frames) // 8 + 2, (w + 3 * frames) // 8 + 2), dtype=np.uint8), np.ones((8, 8), np.uint8))
    x = np.arange(w)[None, :]
    y = np.arange(h)[:, None]
    out = bytearray()
    for t in range(frames):
        base = (((x + 3 * t) // 2 + (y + 2 * t) // 3) % 256).astype(np.float32)
        lum = 0.6 * base + 0.4 * tex[2 * t : 2 * t + h, 3 * t : 3 * t + w] + rng.normal(0, 2, (h, w))
        out += np.clip(lum + 0.5, 0, 255).astype(np.uint8).tobytes()
        cx = np.arange(w // 2)[None, :]
        cy = np.arange(h // 2)[:, None]
        out += ((cx // 8 + t) % 256 + 0 * cy).astype(np.uint8).tobytes()
        out += ((cy // 8 + 2 * t) % 256 + 0 * cx).astype(np.uint8).tobytes()
    return bytes(out)


def encode(w, h, frames, seed, out_path, qp=26, slices=1, ltr=0, cabac=0):
    enc = os.path.join(ROOT, "oracle", "_ref", "h264enc")
    with tempfile.TemporaryDirectory() as td:
        for f in ("welsenc.cfg", "layer2.cfg"):
            shutil.copy(os.path.join(REF, "testbin", f), td)
        open(os.path.join(td, "in.yuv"), "wb").write(synth_yuv(w, h, frames, seed))
        cmd = [enc, "welsenc.cfg", "-org", "in.yuv", "-sw", str(w), "-sh", str(h), "-bf", "out.264", "-frms", str(frames),
               "-numtl", "1", "-ltr", str(ltr), "-dw", "0", str(w), "-dh", "0", str(h), "-frout", "0", "30", "-lqp", "0", str(qp),
               "-rc", "-1", "-aq", "0", "-bgd", "0", "-scene", "0"]
        cmd += ["-slcmd", "0", "1", "-slcnum", "0", str(slices)] if slices > 1 else ["-slcmd", "0", "0"]
        if ltr:
            cmd += ["-ltrnum", "2"]
        if cabac:
            cmd += ["-cabac", "1"]
        subprocess.run(cmd, cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        shutil.copy(os.path.join(td, "out.264"), out_path)
    return os.path.getsize(out_path)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "streams"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    print("syn1080p", encode(1920, 1080, 16, 1234, os.path.join(a.out, "syn1080p_cbp_ippp_16f.264")))
    print("syn720p", encode(1280, 720, 8, 720, os.path.join(a.out, "syn720p_8f.264")))
    print("syn_qcif_multislice", encode(176, 144, 8, 5, os.path.join(a.out, "syn_qcif_4slices_8f.264"), slices=4))
    # CABAC twins (host CABAC parse): multi-slice + two long-term references, and a coarse-QP single-slice stream
    print("syn_cif_cabac", encode(352, 288, 6, 7, os.path.join(a.out, "syn_cif_cabac_4slices_ltr_6f.264"), qp=28, slices=4, ltr=1, cabac=1))
    print("syn_qcif_cabac", encode(176, 144, 8, 9, os.path.join(a.out, "syn_qcif_cabac_qp38_8f.264"), qp=38, cabac=1))
