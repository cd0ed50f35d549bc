#!/usr/bin/env python
"""Per-source-line summary of an `ncu --set full --import-source on` capture, from ncu's OWN source correlation
(`ncu -i x.ncu-rep --page source --print-source cuda,sass --csv`): no external SASS <-> line join (the round-1 tool that joined
nvdisasm line info by address mis-attributed lines).

    python tools/ncu_src.py gpurun_out/x.ncu-rep [top N] > profiles/x_lines.txt

Prints, per CUDA source line, warp-level instructions executed (share of the kernel), stall samples and the dominant stall reasons."""
import csv
import io
import os
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    path, func, hdr = None, None, None
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "" or not r[0].isdigit():
            continue                                   # SASS rows carry no line number
        d = dict(zip(hdr, r))
        # the header names "Source" twice (CUDA text, then SASS text): the first is the CUDA line
        src = r[1]
        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        stalls = {k[6:]: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        lines.append((os.path.basename(path or "?"), int(r[0]), src.strip(), num("Instructions Executed"), num("# Samples"), stalls))
    tot_i = sum(l[3] for l in lines) or 1.0
    tot_s = sum(l[4] for l in lines) or 1.0
    print(f"# {rep}: {func}")
    print(f"# warp instructions executed {tot_i:.4g}, stall samples {tot_s:.4g}; top {top} source lines by instructions")
    print(f"# {'file:line':32s} {'inst %':>7s} {'samples %':>9s}  top stalls | source")
    for f, ln, src, ins, smp, st in sorted(lines, key=lambda l: -l[3])[:top]:
        ts = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        tss = " ".join(f"{k}:{int(v)}" for k, v in ts if v > 0)
        print(f"{f + ':' + str(ln):34s} {100 * ins / tot_i:6.2f}% {100 * smp / tot_s:8.2f}%  {tss:38s} | {src[:110]}")
    print(f"# by file")
    byf = {}
    for f, ln, src, ins, smp, st in lines:
        a = byf.setdefault(f, [0.0, 0.0])
        a[0] += ins
        a[1] += smp
    for f, (i, s) in sorted(byf.items(), key=lambda kv: -kv[1][0]):
        print(f"#   {f:28s} inst {100 * i / tot_i:6.2f}%  samples {100 * s / tot_s:6.2f}%")
    tot_st = {}
    for l in lines:
        for k, v in l[5].items():
            tot_st[k] = tot_st.get(k, 0.0) + v
    print("# stall samples by reason: " + ", ".join(f"{k} {100 * v / tot_s:.1f}%" for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
