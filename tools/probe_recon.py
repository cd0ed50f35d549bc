"""Device-time probe of the reconstruction kernels: S copies of the 1080p workload stream, resident batches."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from losslessh264_b200.api import Context
from losslessh264_b200.records import read_lhfb, n_slots_needed
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
stream = os.path.join(ROOT, "tests/golden/streams/syn1080p_cbp_ippp_16f.264")
with tempfile.TemporaryDirectory() as tmp:
    subprocess.check_call([os.path.join(ROOT, "losslessh264_b200/_lib/lh264_parse_dump"), stream, tmp + "/a.lhfb"], stderr=subprocess.DEVNULL)
    pics = read_lhfb(tmp + "/a.lhfb")
ctx = Context(0)
st = [ctx.open_stream(pics[0].mb_w, pics[0].mb_h, n_slots_needed(pics)) for _ in range(S)]
bs = [ctx.upload_batch([p] * S, st) for p in pics[:6]]
for b in bs: ctx.recon_resident(b)
ctx.sync()
ctx.kernel_ms(-1)
for b in bs[1:]: ctx.recon_resident(b)
print(f"S={S} 5 P pictures: k_decode_rows (reconstruction + deblocking) {ctx.kernel_ms(0):.2f} ms")
