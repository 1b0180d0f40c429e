"""BASELINE config 2 at its full size (100 M rows x 8 INT64, one range predicate): the oracle cannot scan that in test
time, so parity is checked through size-independent properties against the generator's own columns: the selected row
count, per projected column the exact 128-bit SUM / MIN / MAX / COUNT (pushed-down aggregates) and the XOR of the dense
output, the strictly increasing rowkey (order preserved), block-sliced selection offsets, and idempotence of a second scan.
Skipped when the box has too little host memory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROWS = 100_000_000
CHUNK = 4_000_000     # bench.build_workload's chunking (every chunk restarts the RLE run structure)


def test_config2_full_size_properties():
    import os
    import sys
    import torch
    import oceanbase_b200 as ob
    from oceanbase_b200.synth import config2_columns
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 1 << 40
    if avail < 24 << 30:
        pytest.skip("needs ~20 GB of host memory")
    seed = 1234
    w, _ = bench.build_workload(ROWS, 0, seed)
    table = w.table
    # the generator's columns, chunk by chunk -> expected aggregates over the selected rows
    exp_cnt = 0
    exp_sum = [0] * 8
    exp_xor = [np.uint64(0)] * 8
    exp_min = [None] * 8
    exp_max = [None] * 8
    sel_per_block = []
    for s0 in range(0, ROWS, CHUNK):
        n = min(CHUNK, ROWS - s0)
        cols = config2_columns(n, seed, s0)
        m = (cols[4] >= 32) & (cols[4] <= 63)
        exp_cnt += int(m.sum())
        sel_per_block.append(np.add.reduceat(m.astype(np.int64), np.arange(0, n, w.rows_per_block)))
        for c in range(8):
            v = cols[c][m]
            exp_sum[c] += sum(int(x.sum(dtype=np.int64)) for x in np.array_split(v, 16))   # pieces stay far below 2^63
            exp_xor[c] ^= np.bitwise_xor.reduce(v.view(np.uint64)) if len(v) else np.uint64(0)
            if len(v):
                exp_min[c] = int(v.min()) if exp_min[c] is None else min(exp_min[c], int(v.min()))
                exp_max[c] = int(v.max()) if exp_max[c] is None else max(exp_max[c], int(v.max()))
        del cols, m
    sel_per_block = np.concatenate(sel_per_block)
    assert len(sel_per_block) == table.n_blocks

    dev = torch.device("cuda", 0)
    d_image = torch.empty(table.image.size + 64, dtype=torch.uint8, device=dev)
    d_image[:table.image.size].copy_(torch.from_numpy(table.image))
    d_image[table.image.size:].zero_()
    torch.cuda.synchronize()
    ctx = ob.ScanContext(0)
    batch = ctx.open_batch(table, device_image_ptr=d_image.data_ptr())
    assert batch.total_rows == ROWS
    res = batch.scan(w.filter, w.proj, max_selected_rows=int(ROWS * 0.26))
    assert res.selected_rows == exp_cnt
    assert np.array_equal(np.diff(res.fetch_sel_offsets()), sel_per_block)          # every block's slice has the right size
    M = 1 << 128
    for c in range(8):
        assert res.aggregate(ob.AGG_COUNT, c) == exp_cnt
        assert res.aggregate(ob.AGG_SUM, c) % M == exp_sum[c] % M, f"SUM of column {c}"
        assert res.aggregate(ob.AGG_MIN, c) == exp_min[c] and res.aggregate(ob.AGG_MAX, c) == exp_max[c]
        data, _, nulls = res.fetch_col(c)
        assert not nulls.any() and res.col(c).has_null == 0
        assert np.bitwise_xor.reduce(data.view(np.uint64)) == exp_xor[c], f"XOR of column {c}"
        if c == 0:
            assert bool(np.all(np.diff(data.view(np.int64)) > 0))                   # rowkey order: strictly increasing
            first = data.copy()
    # idempotence: a second scan of the same batch gives the same bytes
    res2 = batch.scan(w.filter, w.proj, max_selected_rows=int(ROWS * 0.26))
    assert res2.selected_rows == exp_cnt and np.array_equal(res2.fetch_col(0)[0], first)
    res2.free()
    res.free()
    batch.close()
    ctx.close()
