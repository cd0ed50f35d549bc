#!/usr/bin/env python
"""Writes tests/golden/api_probe/<stream>.<mode>.txt: what the UNMODIFIED reference (oracle/_ref/api_probe_ref = oracle/api_probe.cpp
linked against the reference's decoder objects) shows an API client, call by call.  Runs where /root/reference exists."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "oracle", "_ref", "api_probe_ref")
OUT = os.path.join(ROOT, "tests", "golden", "api_probe")
# error-free streams under BaseDecoderTest's ERROR_CON_SLICE_COPY (the probe's default); the damaged ones under every error
# concealment mode that behaves differently on them: 2 ERROR_CON_SLICE_COPY (hands out what precedes a lost IDR picture, predicted
# from a grey stand-in reference), 4 ERROR_CON_SLICE_COPY_CROSS_IDR, 7 ERROR_CON_SLICE_MV_COPY_CROSS_IDR_FREEZE_RES_CHANGE (what
# h264dec selects, h264dec.cpp:156-157: output frozen until the first IDR picture)
STREAMS = [("SVA_BA1_B.264", None), ("syn_cif_cabac_4slices_ltr_6f.264", None), ("MPS_MW_A.264", None), ("MR1_BT_A.h264", None), ("test_vd_rc.264", None)]
for _n in ("BA_MW_D_IDR_LOST.264", "BA_MW_D_P_LOST.264", "Error_I_P.264"):
    STREAMS += [(_n, 2), (_n, 4), (_n, 7)]

# streams with lost slices (tests/golden/streams/ec_*: the last slice of some pictures taken out of a four-slice stream) under the
# modes that conceal differently: 1 frame copy, 2 slice copy, 4 slice copy across IDR, 6 motion-vector copy, 7 the same with the
# freeze on a change of resolution.  DecodeFrame2 hands the concealed pictures out; DecodeFrameNoDelay loses them (its second
# half clears the output of the first)
for _n in ("ec_qcif_4slices_tail_lost.264", "ec_qcif_4slices_idr_tail_lost.264", "ec_cif_cabac_ltr_tail_lost.264"):
    STREAMS += [(_n, 1), (_n, 2), (_n, 4), (_n, 6), (_n, 7)]

os.makedirs(OUT, exist_ok=True)
for name, ec in STREAMS:
    for mode in ("nodelay", "frame2"):
        r = subprocess.run([PROBE, mode, os.path.join(ROOT, "tests", "golden", "streams", name)] + ([str(ec)] if ec is not None else []), capture_output=True, text=True)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith(("call ", "after "))]      # (the reference also prints its billing table)
        tag = f".ec{ec}" if ec is not None else ""
        open(os.path.join(OUT, f"{name}{tag}.{mode}.txt"), "w").write("\n".join(lines) + "\n")
        print(name, ec, mode, len(lines) // 2, "calls")
