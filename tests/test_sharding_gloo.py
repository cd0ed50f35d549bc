"""CPU, world_size 2 over gloo: the N>1 host logic — every stream lands on exactly one rank, both ranks derive
the same table without communicating, and the end-of-run counter aggregation (sum of work, max of time) is
what bench.py prints."""
import os

import torch.distributed as dist
import torch.multiprocessing as mp

from losslessh264_b200.sharding import aggregate_counters, assign_by_load, assign_round_robin


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = assign_round_robin(1024, world, rank)
    table = assign_by_load([8160] * 5 + [396] * 7 + [99] * 3, world)
    local = {"pictures": 16 * len(mine), "decisions": 1000 * (rank + 1), "step_ms": 10.0 + rank}
    agg = aggregate_counters(local, dist)
    q.put((rank, mine[:4], len(mine), table, agg))
    dist.destroy_process_group()


def test_two_ranks_partition_streams_and_aggregate():
    world, port = 2, 29617
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, first0, n0, table0, agg0), (r1, first1, n1, table1, agg1) = res
    assert first0 == [0, 2, 4, 6] and first1 == [1, 3, 5, 7] and n0 + n1 == 1024
    assert table0 == table1                                  # same table on both ranks, no communication needed
    flat = sorted(i for lst in table0 for i in lst)
    assert flat == list(range(15))                           # every stream on exactly one rank
    sizes = [8160] * 5 + [396] * 7 + [99] * 3
    loads = [sum(sizes[i] for i in lst) for lst in table0]
    assert abs(loads[0] - loads[1]) <= 8160
    assert agg0 == agg1 and agg0["pictures"] == 16 * 1024 and agg0["decisions"] == 3000 and agg0["step_ms"] == 11.0
