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
    outs = pipe.scan(w.table, w.filter, w.proj, blocks_per_batch=7, selectivity_hint=0.05)  # forces overflow re-runs
    pipe.close()
    want = ora.scan_table(w.table, w.filter, w.proj, w.proj_is_string, w.proj_elem_len)
    assert sum(o.selected_rows for o in outs) == want["selected"]
    assert sum(o.total_rows for o in outs) == w.table.total_rows
    for c in range(len(w.proj)):
        got = np.concatenate([o.cols[c] for o in outs])
        assert np.array_equal(got, want["data"][c])
    assert [o.block_begin for o in outs] == list(range(0, w.table.n_blocks, 7))


def test_pipeline_with_ramped_batches():
    from oceanbase_b200.pipeline import HostScanPipeline, batch_bounds
    from oceanbase_b200.synth import make_config2_like
    w = make_config2_like(rows=120_000, rows_per_block=1400, seed=12)
    pipe = HostScanPipeline(0, n_workers=2)
    outs = pipe.scan(w.table, w.filter, w.proj, blocks_per_batch=16, selectivity_hint=0.3, ramp=2)
    pipe.close()
    want = ora.scan_table(w.table, w.filter, w.proj, w.proj_is_string, w.proj_elem_len)
    assert sum(o.selected_rows for o in outs) == want["selected"]
    for c in range(len(w.proj)):
        assert np.array_equal(np.concatenate([o.cols[c] for o in outs]), want["data"][c])
    bounds = batch_bounds(w.table.n_blocks, 16, 2)
    assert [o.block_begin for o in outs] == bounds[:-1] and [o.block_end for o in outs] == bounds[1:]
