"""GPU parity: fused scan through the C-ABI vs the CPU oracle, bit-exact (integer / byte work)."""
import numpy as np
import pytest

import oracle_binding as ora

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import oceanbase_b200 as ob
    c = ob.ScanContext(0)
    yield c
    c.close()


def assert_scan_matches(ctx, w, want_row_ids=True, max_selected_rows=0, agg=None):
    table = w.table
    base = 0x10_0000_0000  # arbitrary non-zero rebasing address for VEC_DISCRETE pointers
    batch = ctx.open_batch(table)
    if agg is not None:     # (aggregate rows, offsets): the scan prunes with the skip index
        batch.set_agg_rows(*agg)
    res = batch.scan(w.filter, w.proj, want_row_ids=want_row_ids, string_base=base,
                     max_selected_rows=max_selected_rows)
    want = ora.scan_table(table, w.filter, w.proj, w.proj_is_string, w.proj_elem_len, string_base=base)
    n = res.selected_rows
    assert res.info().total_rows == table.total_rows == want["total_rows"]
    assert n == want["selected"]
    assert np.array_equal(res.fetch_sel_offsets(), want["sel_offset"])
    if want_row_ids:
        assert np.array_equal(res.fetch_row_ids(), want["row_ids"])
    for i, col in enumerate(w.proj):
        data, lens, nulls = res.fetch_col(i)
        assert np.array_equal(nulls, want["nulls"][i]), f"null bitmap of column {col}"
        assert np.array_equal(data, want["data"][i]), f"payload of column {col}"
        if w.proj_is_string[i]:
            assert np.array_equal(lens, want["lens"][i]), f"lens of column {col}"
            # pointed-to bytes (rule 8c.6): spot check through the image
            off = (data - np.uint64(base)).astype(np.int64)
            for k in range(0, n, max(n // 64, 1)):
                if lens[k] > 0:
                    assert off[k] >= 0 and off[k] + lens[k] <= table.image.size
        assert res.col(i).has_null == int(want["has_null"][i])
    # selection bitmap of a few blocks == ObBitmap image of the oracle's filter
    for b in sorted({0, table.n_blocks // 2, table.n_blocks - 1}):
        blk = ora.Block(table.block(b))
        exp = blk.filter_tree(w.filter) if w.filter is not None else np.ones(blk.row_count, dtype=np.uint8)
        assert np.array_equal(res.fetch_bitmap(b), exp)
    res.free()
    batch.close()
    return n


def test_config1_plumbing(ctx):
    from oceanbase_b200.synth import make_config1
    w = make_config1(rows=100_000, rows_per_block=500, seed=1)
    n = assert_scan_matches(ctx, w)
    assert n == 100_000


@pytest.mark.parametrize("shape", ["bt", "and"])
def test_config2_range_predicate(ctx, shape):
    from oceanbase_b200.synth import make_config2_like
    w = make_config2_like(rows=150_000, rows_per_block=1400, seed=2, shape=shape)
    n = assert_scan_matches(ctx, w)
    assert 0.24 < n / w.table.total_rows < 0.26


def test_config3_dict_and_strings(ctx):
    from oceanbase_b200.synth import make_config3_like
    w = make_config3_like(rows=60_000, rows_per_block=700, seed=3)
    assert_scan_matches(ctx, w)


def test_capacity_overflow_is_reported(ctx):
    import oceanbase_b200 as ob
    from oceanbase_b200.synth import make_config2_like
    w = make_config2_like(rows=50_000, rows_per_block=1400, seed=5)
    batch = ctx.open_batch(w.table)
    res = batch.scan(w.filter, w.proj, max_selected_rows=1000)
    with pytest.raises(ob.ObGpuError) as ei:
        res.info()
    assert ei.value.code == ob.OB_BUF_NOT_ENOUGH
    need = res._info.selected_rows
    res.free()
    res = batch.scan(w.filter, w.proj, max_selected_rows=need)
    assert res.selected_rows == need
    res.free()
    batch.close()
