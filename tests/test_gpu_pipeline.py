"""Host-buffer pipeline (oceanbase_b200.pipeline): batched open/scan/fetch on several streams gives the
same rows as one oracle scan of the whole table."""
import numpy as np
import pytest

import oracle_binding as ora

pytestmark = pytest.mark.gpu


def test_pipeline_matches_oracle():
    from oceanbase_b200.pipeline import HostScanPipeline
    from oceanbase_b200.synth import make_config2_like
    w = make_config2_like(rows=120_000, rows_per_block=1400, seed=11)
    pipe = HostScanPipeline(0, n_workers=3)
    outs = pipe.scan(w.table, w.filter, w.proj, blocks_per_batch=7, selectivity_hint=0.05).batches  # forces overflow re-runs
    pipe.close()
    want = ora.scan_table(w.table, w.filter, w.proj, w.proj_is_string, w.proj_elem_len)
    assert sum(o.selected_rows for o in outs) == want["selected"]
    for c in range(len(w.proj)):
        got = np.concatenate([o.cols[c] for o in outs])
        assert np.array_equal(got, want["data"][c])
    assert [o.block_begin for o in outs] == list(range(0, w.table.n_blocks, 7))


def test_pipeline_with_ramped_batches():
    from oceanbase_b200.pipeline import HostScanPipeline, batch_bounds
    from oceanbase_b200.synth import make_config2_like
    w = make_config2_like(rows=120_000, rows_per_block=1400, seed=12)
    pipe = HostScanPipeline(0, n_workers=2)
    outs = pipe.scan(w.table, w.filter, w.proj, blocks_per_batch=16, selectivity_hint=0.3, ramp=2).batches
    pipe.close()
    want = ora.scan_table(w.table, w.filter, w.proj, w.proj_is_string, w.proj_elem_len)
    assert sum(o.selected_rows for o in outs) == want["selected"]
    for c in range(len(w.proj)):
        assert np.array_equal(np.concatenate([o.cols[c] for o in outs]), want["data"][c])
    bounds = batch_bounds(w.table.n_blocks, 16, 2)
    assert [o.block_begin for o in outs] == bounds[:-1] and [o.block_end for o in outs] == bounds[1:]


def test_pipeline_with_skip_index():
    # every page batch carries its slice of the aggregate rows: same rows as the unpruned oracle scan
    import oceanbase_b200 as ob
    from oceanbase_b200.pipeline import HostScanPipeline
    rng = np.random.default_rng(3)
    n = 90_000
    k = np.sort(rng.integers(0, 1 << 30, size=n, dtype=np.int64))
    v = rng.integers(0, 100, size=n, dtype=np.int64)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_INTEGER_BASE_DIFF, k), ob.Column(ob.OBJ_INT, ob.ENC_RAW, v)]
    table = ob.encode_table(cols, 700)
    rows, offs = ob.table_agg_rows(cols, [0, 1], 700)
    flt = ob.And([ob.White(0, ob.WHITE_OP_BT, (int(k[20_000]), int(k[45_000]))), ob.White(1, ob.WHITE_OP_LT, (60,))])
    pipe = HostScanPipeline(0, n_workers=3)
    outs = pipe.scan(table, flt, [0, 1], blocks_per_batch=11, selectivity_hint=0.3, ramp=2, agg_rows=rows, agg_off=offs).batches
    pipe.close()
    want = ora.scan_table(table, flt, [0, 1], [False, False], [8, 8])
    assert sum(o.selected_rows for o in outs) == want["selected"]
    for c in range(2):
        assert np.array_equal(np.concatenate([o.cols[c] for o in outs]), want["data"][c])


def test_pipeline_strings_and_pushed_down_aggregates():
    """VARCHAR columns come back as (pointer into the caller's image, length) per batch; aggregates are folded on the
    device per page batch and summed in the library: with no_row_output only 16 bytes per aggregate and batch return."""
    import oceanbase_b200 as ob
    from oceanbase_b200.pipeline import HostScanPipeline
    from oceanbase_b200.synth import make_config3_like
    w = make_config3_like(rows=40_000, rows_per_block=133, seed=3)
    base = w.table.image.ctypes.data
    pipe = HostScanPipeline(0, n_workers=3)
    out = pipe.scan(w.table, w.filter, w.proj, blocks_per_batch=40, selectivity_hint=0.14, ramp=2, string_base=base,
                    proj_is_string=w.proj_is_string, proj_elem_len=w.proj_elem_len,
                    aggs=[(ob.AGG_COUNT, 0, -1), (ob.AGG_SUM, 1, -1), (ob.AGG_MAX, 2, -1), (ob.AGG_SUM_PRODUCT, 0, 3)])
    want = ora.scan_table(w.table, w.filter, w.proj, w.proj_is_string, w.proj_elem_len, string_base=base)
    assert out.selected_rows == want["selected"] and out.total_rows == w.table.total_rows
    for c in range(len(w.proj)):
        got = np.concatenate([b.cols[c] for b in out.batches])
        assert np.array_equal(got, want["data"][c]), c
        if w.proj_is_string[c]:
            assert np.array_equal(np.concatenate([b.lens[c] for b in out.batches]), want["lens"][c])
    v = [want["data"][c].view(np.int64) for c in range(4)]
    assert out.aggregates[0] == want["selected"]
    assert out.aggregates[1] == int(sum(int(x) for x in v[1]))
    assert out.aggregates[2] == int(v[2].max())
    assert out.aggregates[3] == int(sum(int(a) * int(b) for a, b in zip(v[0], v[3])))
    only = pipe.scan(w.table, w.filter, w.proj, blocks_per_batch=40, selectivity_hint=0.01, no_row_output=True,
                     aggs=[(ob.AGG_SUM_PRODUCT, 0, 3)])
    assert only.aggregates[0] == out.aggregates[3] and only.selected_rows == out.selected_rows
    assert only.d2h_bytes == 16 * len(only.batches)
    pipe.close()


def test_pipeline_zero_copy_reads_pinned_host_memory():
    """zero_copy: the table image stays in pinned host memory and the kernels read what they reference over PCIe; same rows
    as the staged pipeline and as the oracle, and the library reports no host->device copy of its own."""
    import torch
    import oceanbase_b200 as ob
    from oceanbase_b200.pipeline import HostScanPipeline
    from oceanbase_b200.synth import make_config3_like
    w = make_config3_like(rows=30_000, rows_per_block=133, seed=4)
    pinned = torch.empty(w.table.image.size + 256, dtype=torch.uint8).pin_memory()
    img = pinned.numpy()
    img[:w.table.image.size] = w.table.image
    img[w.table.image.size:] = 0
    table = ob.TableImage(img[:w.table.image.size], w.table.offsets, w.table.sizes, w.table.total_rows, w.table.n_cols)
    base = table.image.ctypes.data
    pipe = HostScanPipeline(0, n_workers=3)
    kw = dict(blocks_per_batch=50, selectivity_hint=0.14, ramp=2, string_base=base, proj_is_string=w.proj_is_string, proj_elem_len=w.proj_elem_len)
    zc = pipe.scan(table, w.filter, w.proj, zero_copy=True, **kw)
    st = pipe.scan(table, w.filter, w.proj, zero_copy=False, **kw)
    want = ora.scan_table(table, w.filter, w.proj, w.proj_is_string, w.proj_elem_len, string_base=base)
    assert zc.selected_rows == st.selected_rows == want["selected"]
    assert zc.h2d_bytes == 0 and st.h2d_bytes >= int(w.table.sizes.sum())
    for c in range(len(w.proj)):
        got = np.concatenate([b.cols[c] for b in zc.batches])
        assert np.array_equal(got, want["data"][c]), c
        assert np.array_equal(got, np.concatenate([b.cols[c] for b in st.batches]))
        if w.proj_is_string[c]:
            assert np.array_equal(np.concatenate([b.lens[c] for b in zc.batches]), want["lens"][c])
    pipe.close()
