#!/usr/bin/env python
"""Summary of an `ncu --set full` capture (one kernel launch) as JSON, from ncu's own raw page:

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "S=1024 P pictures, tools/probe_recon.py 1024" > profiles/r2_x_ncu_summary.json

bench.py reads `dram_bytes_per_unit` (dram__bytes_read.sum + dram__bytes_write.sum divided by the units of the captured launch)
from the committed summary of the dominant kernel for its `roofline.traffic`."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_static": "static_smem",
    "lts__t_bytes.sum": "l2_bytes",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving_per_issue",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction_per_issue",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar_per_issue",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle_per_issue",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle_per_issue",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    units_per_launch = float(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    s = {"report": rep, "note": note}
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            s["kernel"] = v.split("(")[0]
        if h in WANT and v != "":
            x = float(v.replace(",", ""))
            if u in SCALE and ("bytes" in h or "duration" in h):
                x *= SCALE[u]
            s[WANT[h]] = x
    s["duration_unit"] = "ms"
    if "dram_read" in s and "dram_write" in s:
        s["dram_bytes"] = s["dram_read"] + s["dram_write"]
        if units_per_launch:
            s["units_per_launch"] = units_per_launch
            s["dram_bytes_per_unit"] = s["dram_bytes"] / units_per_launch
            if "warp_instructions" in s:
                s["warp_instructions_per_unit"] = s["warp_instructions"] / units_per_launch
    print(json.dumps(s, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
