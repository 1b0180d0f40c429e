"""N>1 host logic on CPU (gloo, world_size 2): every rank builds ITS shard of the config-2 table
(row_start = rank * rows, no data-path collective), scans it with the oracle port, and only the
row / selected counts are all-reduced -- the same structure bench.py uses with NCCL. The shards
together must equal one process scanning both shards."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, rows, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import oracle_binding as ora
    w, _ = bench.build_workload(rows, rank * rows, seed=2, chunk_rows=rows // 2)
    total, sel, cs = ora.scan_table_mt(w.table, w.filter, w.proj, batch_size=256, n_threads=2)
    t = torch.tensor([total, sel, cs % (1 << 62)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((t.tolist(), ms.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_process():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import bench
    import oracle_binding as ora
    rows, world = 40_000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, rows, 29613, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, max_ms = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_total = exp_sel = exp_cs = 0
    for r in range(world):
        w, _ = bench.build_workload(rows, r * rows, seed=2, chunk_rows=rows // 2)
        t, s, c = ora.scan_table_mt(w.table, w.filter, w.proj, batch_size=256, n_threads=1)
        exp_total, exp_sel, exp_cs = exp_total + t, exp_sel + s, exp_cs + c % (1 << 62)
    assert got == [exp_total, exp_sel, exp_cs]
    assert exp_total == world * rows and 0.24 < exp_sel / exp_total < 0.26
    assert max_ms == float(world)  # MAX over ranks picks the slowest rank


def test_rank_shards_are_disjoint_continuations():
    # shard r starts where shard r-1 ends: the sorted PK keeps increasing across shards
    sys.path.insert(0, ROOT)
    import bench
    import oracle_binding as ora
    a, _ = bench.build_workload(20_000, 0, seed=2, chunk_rows=10_000)
    b, _ = bench.build_workload(20_000, 20_000, seed=2, chunk_rows=10_000)
    last_a = ora.Block(a.table.block(a.table.n_blocks - 1))
    first_b = ora.Block(b.table.block(0))
    assert last_a.cell(0, last_a.row_count - 1) < first_b.cell(0, 0)
