"""Phase B of the range-partitioned compaction on CPU (gloo, world_size 2): every rank encodes the merged rows of ITS rowkey range
(on the GPU box: obgpu_merge_result_encode; here the host writer stands in, the device encoder being byte-identical to it) and
computes the range's column checksums (oracle); then
  * the column checksums, all_reduced (SUM, wrapping int64), equal the checksums of the whole merged stream -- the property the
    reference relies on when it adds the micro-blocks' column checksums up per SSTable (ob_micro_block_checksum_helper.cpp:127-146);
  * the ranks' block images, concatenated in rank order, decode (oracle) to exactly the merged stream ONE process produces."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_binding as ora
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_table
    from oceanbase_b200.synth import make_config5_runs
    from test_gloo_compaction import _to_decoded, _oracle_merge_fn
    from oceanbase_b200.compaction import distributed_major_merge
    runs = make_config5_runs(n_runs=4, window=5000, seed=31, encode=False)
    local = {q: _to_decoded(runs[q]) for q in range(4) if q % world == rank}
    m, _splitters, _recv = distributed_major_merge(local, 4, 3, _oracle_merge_fn, samples_per_run=64)
    n = len(m["key"])
    o = ora.oracle()
    cks = [o.ora_column_checksum(np.ascontiguousarray(m["key"]).ctypes.data, None, n, 8)]
    for c in range(3):
        cks.append(o.ora_column_checksum(np.ascontiguousarray(m["vals"][c]).ctypes.data, np.ascontiguousarray(m["null"][c]).ctypes.data, n, 8))
    t = torch.tensor(cks, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)     # int64 addition wraps exactly like the reference's column checksum sums
    cols = [Column(capi.OBJ_INT, capi.ENC_RAW, m["key"])] + [
        Column(capi.OBJ_INT, capi.ENC_RAW, m["vals"][c], nulls=m["null"][c] if m["null"][c].any() else None) for c in range(3)]
    table = encode_table(cols, 500, rowkey_cnt=1) if n else None
    out.put((rank, n, t.numpy().copy(), None if table is None else (table.image, table.offsets, table.sizes)))
    dist.barrier()
    dist.destroy_process_group()


def test_checksums_add_up_and_images_concatenate():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle_binding as ora
    from oceanbase_b200.synth import make_config5_runs
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29671, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    runs = make_config5_runs(n_runs=4, window=5000, seed=31, encode=False)
    want = ora.major_merge(runs, 3)
    n = len(want["key"])
    assert sum(g[1] for g in got) == n
    o = ora.oracle()
    whole = [o.ora_column_checksum(np.ascontiguousarray(want["key"]).ctypes.data, None, n, 8)] + [
        o.ora_column_checksum(np.ascontiguousarray(want["vals"][c]).ctypes.data, np.ascontiguousarray(want["null"][c]).ctypes.data, n, 8)
        for c in range(3)]
    for g in got:
        assert [int(x) for x in g[2]] == whole          # every rank holds the reduced sums
    # the ranks' blocks, in rank order, decode to the merged stream
    at = 0
    for g in got:
        if g[3] is None:
            continue
        image, offsets, sizes = g[3]
        for b in range(len(offsets)):
            blk = ora.Block(np.ascontiguousarray(image[offsets[b]:offsets[b] + sizes[b]]))
            assert blk.verify_checksums() == 0
            for r in (0, blk.row_count // 2, blk.row_count - 1):
                assert (blk.cell(0, r) & 0xffffffffffffffff) == int(want["key"][at + r]) & 0xffffffffffffffff
                for c in range(3):
                    cell = blk.cell(1 + c, r)
                    if want["null"][c][at + r]:
                        assert cell is None
                    else:
                        assert (cell & 0xffffffffffffffff) == int(want["vals"][c][at + r]) & 0xffffffffffffffff
            at += blk.row_count
    assert at == n
