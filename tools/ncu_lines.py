"""Aggregates the warp-stall samples of an ncu report (--set full --import-source on) per CUDA source line.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel substring> [top N]

ncu's CSV source page lists SASS with samples but no line numbers; nvdisasm -g on the cubin extracted from the in-tree
library gives the SASS offset -> source line table.  The two are joined on the instruction offset inside the function.
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("NCU_LINES_BIN") or os.path.join(ROOT, "losslessh264_b200", "_lib", "liblh264_b200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
base = int(data[0][ix["Address"]], 16)
samples = {int(r[ix["Address"]], 16) - base: (int(r[ix["# Samples"]] or 0), r[ix["Source"]], int(r[ix["Instructions Executed"]] or 0)) for r in data}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = {int(r[ix["Address"]], 16) - base: {h[6:]: int(r[ix[h]] or 0) for h in stall_cols} for r in data}
with tempfile.TemporaryDirectory() as t:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=t, capture_output=True)
    cubin = [f for f in os.listdir(t) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(t, cubin)], capture_output=True, text=True).stdout
# find the function section
line_of = {}
cur_line, in_fn = None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln) or re.match(r"//-+ \.text\.(\S+)", ln)
    if m:
        in_fn = kern in m.group(1)
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur_line
per = collections.Counter(); inst = collections.Counter(); why = collections.defaultdict(collections.Counter)
tot = 0
for off, (n, src, ie) in samples.items():
    key = line_of.get(off, ("?", 0))
    per[key] += n; inst[key] += ie; tot += n
    for k, v in stalls[off].items():
        why[key][k] += v
srcs = {}
print(f"total samples {tot}")
for key, n in per.most_common(top):
    f, l = key
    if f not in srcs:
        pth = os.path.join(ROOT, "losslessh264_b200", "csrc", f)
        srcs[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    w = ", ".join(f"{k} {v}" for k, v in why[key].most_common(2))
    print(f"{100 * n / tot:5.1f}%  {inst[key]:>10d} inst  {f}:{l:<4d} {text}   [{w}]")
