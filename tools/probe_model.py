"""Quick device-time probe of the model kernels: S copies of the 1080p workload stream, one picture per stream per submit."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from losslessh264_b200.api import Context
from losslessh264_b200.records import read_lhmd

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
stream = os.path.join(ROOT, "tests/golden/streams", sys.argv[2] if len(sys.argv) > 2 else "syn1080p_cbp_ippp_16f.264")
tool = os.path.join(ROOT, "losslessh264_b200/_lib/lh264_parse_dump")
with tempfile.TemporaryDirectory() as tmp:
    subprocess.check_call([tool, stream, tmp + "/a.lhfb", "1", tmp + "/a.lhmd"], stderr=subprocess.DEVNULL)
    pics = read_lhmd(tmp + "/a.lhmd")
ctx = Context(0)
hs = [ctx.model_open(pics[0].mb_w * pics[0].mb_h) for _ in range(S)]
ctx.model_submit([pics[0]] * S, hs)          # warm
ctx.kernel_ms(-1)
t0 = time.perf_counter()
for p in pics[1:]:
    ctx.model_submit([p] * S, hs)
wall = time.perf_counter() - t0
gen, code = ctx.kernel_ms(4), ctx.kernel_ms(5)
ctx.kernel_ms(-2)
dec = ctx.model_decisions(hs[0])
n = len(pics) - 1
print(f"S={S} pictures={n} decisions/stream={dec} gen {gen:.1f} ms code {code:.1f} ms wall {wall*1000:.1f} ms -> {S*n/(wall):.0f} pictures/s; code: {S*dec/ (code/1000)/1e9:.3f} G decisions/s")
