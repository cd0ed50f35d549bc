"""Static sharding of independent streams over the GPUs of one box.

Nothing inside a stream can be split byte-exactly (adaptive priors and the previous-frame context chain the
whole file), so the unit is the stream: stream i goes to rank i mod world, or — for mixed resolutions — to
the least loaded rank by macroblock count (greedy, deterministic).  No collective is needed on the data path;
only throughput counters are summed at the end.
"""
from __future__ import annotations


def assign_round_robin(n_streams: int, world: int, rank: int) -> list[int]:
    return [i for i in range(n_streams) if i % world == rank]


def assign_by_load(mbs_per_stream: list[int], world: int) -> list[list[int]]:
    """Greedy longest-processing-time assignment; ties broken by the lower rank, so every rank computes the
    same table without communicating."""
    order = sorted(range(len(mbs_per_stream)), key=lambda i: (-mbs_per_stream[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += mbs_per_stream[i]
    for lst in out:
        lst.sort()
    return out


def aggregate_counters(local: dict, dist=None) -> dict:
    """Sums per-rank counters (pictures, decisions, bytes) and takes the max of times over ranks.  `dist` is
    torch.distributed (nccl on GPUs, gloo in the CPU tests) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    import torch

    keys = sorted(local)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    sums = torch.tensor([float(local[k]) for k in keys if not k.endswith("_ms")], dtype=torch.float64, device=dev)
    maxs = torch.tensor([float(local[k]) for k in keys if k.endswith("_ms")], dtype=torch.float64, device=dev)
    if sums.numel():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if maxs.numel():
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out, si, mi = {}, 0, 0
    for k in keys:
        if k.endswith("_ms"):
            out[k] = float(maxs[mi]); mi += 1
        else:
            out[k] = float(sums[si]); si += 1
    return out
