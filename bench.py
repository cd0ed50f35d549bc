#!/usr/bin/env python
"""bench.py — 1080p decode frames/s of the B200 hot path (BASELINE.json config 4).

Workload: S independent 1920x1080 Constrained-Baseline IPPP streams per GPU.  Every stream decodes
the same 16-picture synthetic sequence (tests/golden/streams/syn1080p_cbp_ippp_16f.264, made by the
reference encoder from the seeded generator of tools/make_synth_stream.py); stream state (DPB
slots, records) is per stream, so the device does S times the work.  The hand-off records are
produced by the host CAVLC parser (losslessh264_b200/host) before the timed region.

One "step" = one coded picture for every stream = S pictures reconstructed (inter pass, intra
row wavefront, deblocking row wavefront: 2 kernel launches).
  value : records resident in HBM when the timed region starts (lh264_batch_recon).
  e2e   : records in pinned HOST memory -> lh264_recon_submit (H2D inside) -> decoded I420 of every
          picture back to pinned host memory (lh264_pictures_read_packed), all inside the timed region.
  roofline : dominant kernel family, algorithmic bytes / CUDA-event time, vs the
          measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline / --impl reference : the compiled reference (oracle/_ref/h264dec: parse + model +
          reconstruction, C kernels) on the box's host cores, same stream.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAM = os.path.join(ROOT, "tests", "golden", "streams", "syn1080p_cbp_ippp_16f.264")
PARSE_TOOL = os.path.join(ROOT, "losslessh264_b200", "_lib", "lh264_parse_dump")
REF_DEC = os.path.join(ROOT, "oracle", "_ref", "h264dec")
METRIC, UNIT = "1080p decode frames/s", "frames/s"
WORKLOAD = "syn1080p_cbp_ippp_16f x S independent streams (BASELINE config 4), 1 picture per stream per step"
MB_PER_PIC = 120 * 68
PIC_BYTES = 1920 * 1088 * 3 // 2


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# algorithmic bytes per picture (SURVEY.md §8d / DESIGN.md): P*(write + deblock read + deblock write + r) + N*(768+128)
def algorithmic_bytes(is_p):
    return PIC_BYTES * (3 + (1 if is_p else 0)) + MB_PER_PIC * 896


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons, "samples": len(self.rows)}


def run_reference_once(k, tmp):
    """k concurrent reference processes (the reference is single-threaded with process-global model
    state, one stream per process).  Returns (frames/s by the tool's own decode time, frames/s by wall)."""
    procs = []
    t0 = time.time()
    for i in range(k):
        d = os.path.join(tmp, f"p{i}")
        os.makedirs(d, exist_ok=True)
        procs.append(subprocess.Popen([REF_DEC, STREAM, os.path.join(d, "o.pip"), os.path.join(d, "o.yuv")], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True))
    fps_core, frames = 0.0, 0
    for p in procs:
        err = p.communicate()[1]
        m = re.search(r"Frames:\s+(\d+)\s+decode time:\s+([\d.]+) sec", err)
        if p.returncode != 0 or not m:
            raise RuntimeError("reference h264dec failed: " + err[-300:])
        frames += int(m.group(1))
        fps_core += int(m.group(1)) / float(m.group(2))
    wall = time.time() - t0
    return fps_core, frames / wall, frames


def reference_concurrency():
    cores = os.cpu_count() or 1
    try:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
    except Exception:
        mem_gb = 16
    # each reference process needs ~9.3 GB for its dense prior table; cap to keep the run bounded
    return max(1, min(cores, int(mem_gb // 10), 32)), cores


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_DEC):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/h264dec not built"}))
        return
    k, cores = reference_concurrency()
    with tempfile.TemporaryDirectory() as tmp:
        for _ in range(min(args.warmup, 1)):
            run_reference_once(k, tmp)
        steps = max(1, min(args.steps, 3))
        t0 = time.time()
        core, wall, frames = 0.0, 0.0, 0
        for _ in range(steps):
            c, w, f = run_reference_once(k, tmp)
            core += c
            wall += w
            frames += f
        dt = time.time() - t0
    v = core / steps
    sample = f"{k} concurrent h264dec processes x 16 frames of the workload stream, {steps} runs; value = sum of per-process frames / tool-reported decode time (parse + model + reconstruction; excludes the ~8 s prior-table construction per process); including start-up: {wall / steps:.2f} frames/s"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": round(1000 * dt / steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": k, "host_cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def parse_stream():
    from losslessh264_b200.records import read_lhfb

    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "s.lhfb")
        r = subprocess.run([PARSE_TOOL, STREAM, out, "3"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host parser failed: " + r.stderr[-300:])
        m = re.search(r"\(([\d.]+) pictures/s\)", r.stderr)
        return read_lhfb(out), float(m.group(1)) if m else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--streams", type=int, default=256, help="independent 1080p streams per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return bench_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from losslessh264_b200.api import Context, PinnedArray
    from losslessh264_b200.records import n_slots_needed

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, K, W = args.streams, args.steps, max(args.warmup, 3)

    pics, parse_pps = parse_stream()
    n_frames = len(pics)
    mb_w, mb_h = pics[0].mb_w, pics[0].mb_h
    tstream = torch.cuda.Stream()
    ctx = Context(local, tstream.cuda_stream)
    streams = [ctx.open_stream(mb_w, mb_h, n_slots_needed(pics)) for _ in range(S)]
    batches = [ctx.upload_batch([p] * S, streams) for p in pics]   # S independent copies of every picture's records
    ctx.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: records resident in HBM -------------------------------------------------------
    # every step touches S x (records + 2-4 picture planes) = S x ~14 MB >> 126 MB L2: inputs larger than L2
    step_no = 0

    def resident_steps(n):
        nonlocal step_no
        for _ in range(n):
            ctx.recon_resident(batches[step_no % n_frames])
            step_no += 1

    with torch.cuda.stream(tstream):
        resident_steps(W + (-W) % n_frames)            # warm-up ends on a sequence boundary
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = ctx.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tstream)
        resident_steps(K)
        e1.record(tstream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = ctx.kernel_launches() - l0
        sampler.stop_flag = True
    value = world * S * K / (ms / 1000.0)

    # ---- roofline of the dominant kernel family (CUDA events on the launching stream) ------------
    step_no = 0
    ctx.kernel_ms(-1)                                   # reset + enable per-family event timing
    resident_steps(n_frames)
    fam = [ctx.kernel_ms(i) for i in range(3)]
    ctx.kernel_ms(-2)
    fam_names = ["k_recon_rows", "(unused)", "k_deblock_rows"]
    dom = int(np.argmax(fam))
    n_p = sum(1 for p in pics if int(p.slices["slice_type"][0]) == 0)
    alg_seq = S * (n_p * algorithmic_bytes(True) + (n_frames - n_p) * algorithmic_bytes(False))
    peak, peak_kind = measured_peak()
    # whole reconstruction (all three families) against the whole-picture algorithmic bytes
    ach_all = alg_seq / (sum(fam) / 1000.0) / 1e9
    # dominant family: its own share of the algorithmic bytes: k_recon_rows = reference read (P pictures) +
    # reconstruction write + records and coefficients; k_deblock_rows = picture read + write + MB metadata
    share = [S * (n_p * PIC_BYTES + n_frames * (PIC_BYTES + MB_PER_PIC * 896)), 0, S * n_frames * (2 * PIC_BYTES + MB_PER_PIC * 128)]
    ach_dom = share[dom] / (fam[dom] / 1000.0) / 1e9
    roofline = {"bound": "hbm", "kernel": fam_names[dom], "achieved": round(ach_dom, 1), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": round(ach_dom / peak, 4), "traffic": None,
                "whole_step": {"achieved": round(ach_all, 1), "frac": round(ach_all / peak, 4)},
                "family_ms_per_16_steps": {n: round(v, 3) for n, v in zip(fam_names, fam) if n[0] != "("}}

    # ---- e2e: host buffers in, decoded pictures back to the host ---------------------------------
    Se = min(S, 64)                                    # pictures per submit on the host-buffer path
    out_pin = PinnedArray(Se * 1920 * 1080 * 3 // 2)
    e_streams = streams[:Se]
    h2d = Se * int(pics[1].mbs.nbytes + pics[1].coefs.nbytes + pics[1].slices.nbytes)
    d2h = Se * 1920 * 1080 * 3 // 2

    def e2e_steps(n, start):
        for i in range(n):
            p = pics[(start + i) % n_frames]
            ctx.recon([p] * Se, e_streams, sync=False)
            ctx.read_packed(e_streams, [p.cur_slot] * Se, (0, 0, 1920, 1080), out_pin)

    e2e_steps(n_frames, 0)
    barrier()
    Ke = max(n_frames, min(K, 32))
    t0 = time.perf_counter()
    e2e_steps(Ke, 0)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
    e2e_value = world * Se * Ke / (e2e_ms / 1000.0)

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms / K, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams_per_gpu": S, "pictures_per_step_per_gpu": S, "resolution": "1920x1080 (coded 1920x1088)",
                   "l2": "inputs larger than L2 (each step touches > 1 GB of records and pictures)", "host_parse_pictures_per_s_per_core": parse_pps},
        "roofline": roofline,
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "pictures_per_step": Se,
                "path": "lh264_recon_submit (host records) + lh264_pictures_read_packed (I420 to pinned host)"},
        "gpu_launches": int(launches), "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_DEC):
        k, cores = reference_concurrency()
        with tempfile.TemporaryDirectory() as tmp:
            core, wall, frames = run_reference_once(k, tmp)
        line["cpu_baseline"] = {"value": round(core, 2), "unit": UNIT, "cores": k, "host_cores": cores, "kind": "reference",
                                "sample": f"{k} concurrent oracle/_ref/h264dec processes x 16 frames of the workload stream (parse + model + reconstruction, tool-reported decode time; {wall:.2f} frames/s including the ~8 s start-up per process)"}
    for b in batches:
        ctx.free_batch(b)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
