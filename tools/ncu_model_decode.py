"""Profiling driver: one lh264_model_decode call (k_model_decode) on N copies of a corpus stream's `.pip`, through the product's
C ABI.  Used under ncu (`ncu --set full -k regex:k_model_decode ...`); prints the kernel time of the un-profiled call otherwise.
    python tools/ncu_model_decode.py [stream] [copies]"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from losslessh264_b200.api import Context
from losslessh264_b200 import pip as pipmod
from losslessh264_b200.records import read_lhmd

name = sys.argv[1] if len(sys.argv) > 1 else "BA1_FT_C.264"
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 592
src = os.path.join(ROOT, "tests", "golden", "streams", name)
tool = os.path.join(ROOT, "losslessh264_b200", "_lib", "lh264_parse_dump")
ctx = Context(0)
with tempfile.TemporaryDirectory() as t:
    subprocess.check_call([tool, src, os.path.join(t, "s.lhfb"), "1", os.path.join(t, "s.lhmd"), os.path.join(t, "s.pip")], stderr=subprocess.DEVNULL)
    default = open(os.path.join(t, "s.pip"), "rb").read()
    pics = read_lhmd(os.path.join(t, "s.lhmd"))
ms = ctx.model_open(max(p.mb_w * p.mb_h for p in pics))
for p in pics:
    ctx.model_submit([p], [ms])
tags = ctx.model_finish(ms)
decisions = ctx.model_decisions(ms)
ctx.model_close(ms)
plan = pipmod.Plan(default)
ctx.kernel_ms(-1)
res = ctx.model_decode([plan.job(tags)] * copies)
ms_ = ctx.kernel_ms(6)
ok = plan.assemble(*res[0]) == open(src, "rb").read()
print(f"{name} x {copies}: k_model_decode {ms_:.1f} ms, {len(pics) * copies / (ms_ / 1000.0):.0f} frames/s, {decisions * copies / (ms_ / 1000.0) / 1e9:.2f} G decisions/s, identical={ok}")
plan.close()
