"""Device-time probe of k_decode_rows on INTRA pictures: S copies of the IDR picture of the 1080p workload stream, resident batches."""
import os, subprocess, sys, tempfile, time
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
b = ctx.upload_batch([pics[0]] * S, st)
ctx.recon_resident(b)
ctx.sync()
for i in range(3):
    ctx.kernel_ms(-1)
    t0 = time.perf_counter()
    ctx.recon_resident(b)
    ctx.sync()
    print(f"S={S} I picture launch {i}: k_decode_rows {ctx.kernel_ms(0):.2f} ms by events, {1000 * (time.perf_counter() - t0):.2f} ms wall", flush=True)
