"""The one exchange step of the path on CPU (gloo, world_size 2 and 3): range-partitioned major merge.
Every rank holds some of the K runs, splitters are picked from gathered rowkey samples, run slices travel
to the rank owning their rowkey range (batch_isend_irecv -- NCCL on the GPU box), every rank merges its
range (here with the oracle standing in for the device merge) and the per-rank outputs concatenated in
rank order must equal ONE process merging all the runs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _to_decoded(run):
    import torch
    from oceanbase_b200.compaction import DecodedRun
    return DecodedRun(torch.from_numpy(run["key"].copy()), torch.from_numpy(run["flag"].copy()),
                      [torch.from_numpy(v.copy()) for v in run["vals"]], [torch.from_numpy(e.copy()) for e in run["ext"]])


def _oracle_merge_fn(runs):
    import oracle_binding as ora
    as_np = [{"key": r.key.numpy(), "flag": None if r.flag is None else r.flag.numpy(),
              "vals": [v.numpy() for v in r.vals], "ext": [e.numpy() for e in r.ext]} for r in runs]
    return ora.major_merge(as_np, 3)


def _worker(rank, world, n_runs, window, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oceanbase_b200.synth import make_config5_runs
    from oceanbase_b200.compaction import distributed_major_merge
    runs = make_config5_runs(n_runs=n_runs, window=window, seed=21, encode=False)
    local = {q: _to_decoded(runs[q]) for q in range(n_runs) if q % world == rank}
    m, splitters, recv_rows = distributed_major_merge(local, n_runs, 3, _oracle_merge_fn, samples_per_run=64)
    out.put((rank, m["key"], m["vals"], m["null"], m["dropped"], m["fused"], splitters.numpy(), recv_rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_runs", [(2, 4), (3, 5)])
def test_range_partitioned_merge_equals_single_process(world, n_runs):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle_binding as ora
    from oceanbase_b200.synth import make_config5_runs
    window = 6000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, n_runs, window, 29640 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    runs = make_config5_runs(n_runs=n_runs, window=window, seed=21, encode=False)
    want = ora.major_merge(runs, 3)
    keys = np.concatenate([g[1] for g in got])
    assert np.array_equal(keys, want["key"])                     # rank order == global rowkey order
    for c in range(3):
        assert np.array_equal(np.concatenate([g[2][c] for g in got]), want["vals"][c])
        assert np.array_equal(np.concatenate([g[3][c] for g in got]), want["null"][c])
    assert sum(g[4] for g in got) == want["dropped"] and sum(g[5] for g in got) == want["fused"]
    # every rank agrees on the splitters, every input row was sent to exactly one rank, load is balanced
    for g in got[1:]:
        assert np.array_equal(g[6], got[0][6])
    assert sum(int(g[7].sum()) for g in got) == sum(len(r["key"]) for r in runs)
    share = [int(g[7].sum()) for g in got]
    assert max(share) < 1.5 * sum(share) / world
    # all rows of one rowkey land on one rank
    for a, b in zip(got[:-1], got[1:]):
        assert a[1][-1] < b[1][0]


def _worker_composite(rank, world, n_runs, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_binding as ora
    from test_major_merge_kat import composite_runs
    from oceanbase_b200.compaction import DecodedRun, distributed_major_merge
    runs = composite_runs(np.random.default_rng(7), n_runs, 4000, 2)
    t = torch.from_numpy
    local = {q: DecodedRun(t(r["key"].copy()), t(r["flag"].copy()), [t(r["vals"][0].copy())], [t(r["ext"][0].copy())],
                           [t(k.copy()) for k in r["more_keys"]]) for q, r in enumerate(runs) if q % world == rank}

    def merge_fn(rs):
        return ora.major_merge([{"key": r.key.numpy(), "flag": r.flag.numpy(), "vals": [r.vals[0].numpy()], "ext": [r.ext[0].numpy()],
                                 "more_keys": [k.numpy() for k in r.more_keys]} for r in rs], 1)

    m, _, _ = distributed_major_merge(local, n_runs, 1, merge_fn, samples_per_run=64)
    out.put((rank, m["key"], m["more_keys"], m["vals"][0], m["null"][0]))
    dist.barrier()
    dist.destroy_process_group()


def test_range_partitioned_merge_with_composite_rowkeys():
    # the partition cuts on the first rowkey column only, so rows that tie on it stay on one rank; the other rowkey
    # columns travel with the payload
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle_binding as ora
    from test_major_merge_kat import composite_runs
    world, n_runs = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_composite, args=(r, world, n_runs, 29677, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = ora.major_merge(composite_runs(np.random.default_rng(7), n_runs, 4000, 2), 1)
    assert np.array_equal(np.concatenate([g[1] for g in got]), want["key"])
    for c in range(2):
        assert np.array_equal(np.concatenate([g[2][c] for g in got]), want["more_keys"][c])
    assert np.array_equal(np.concatenate([g[3] for g in got]), want["vals"][0])
    assert np.array_equal(np.concatenate([g[4] for g in got]), want["null"][0])
    assert all(len(g[1]) > 0 for g in got) and got[0][1][-1] < got[1][1][0]
