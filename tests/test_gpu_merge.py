"""Major-compaction merge on the device vs the oracle (bit-exact row stream): the reference's row-fuse
expectations as single-row runs, config-5 runs decoded from their PAX SSTables by the device
(obgpu_batch_decode_column incl. NULL / NOP ext values), odd run counts, empty runs, default rows."""
import numpy as np
import pytest

import oracle_binding as ora
from test_major_merge_kat import FUSE_CASES, case_runs, expected_row

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import oceanbase_b200
    return oceanbase_b200


@pytest.fixture(scope="module")
def env(ob):
    import torch
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    c = ob.ScanContext(0, stream=stream.cuda_stream)
    yield c, torch
    c.close()


def to_dev(torch, run):
    from oceanbase_b200.compaction import DecodedRun
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    return DecodedRun(t(run["key"], np.int64), None if run.get("flag") is None else t(run["flag"], np.uint8),
                      [t(v, np.int64) for v in run["vals"]], [t(e, np.uint8) for e in run["ext"]])


def assert_merge_equal(res, want, n_cols):
    info = res.info()
    assert info.out_rows == len(want["key"])
    assert info.dropped_deletes == want["dropped"] and info.fused_rows == want["fused"]
    k, _ = res.fetch(-1)
    assert np.array_equal(k, want["key"])
    for c in range(n_cols):
        v, nl = res.fetch(c)
        assert np.array_equal(nl, want["null"][c]), f"null bytes of column {c}"
        assert np.array_equal(v, want["vals"][c]), f"values of column {c}"


@pytest.mark.parametrize("name", sorted(FUSE_CASES))
@pytest.mark.parametrize("defaults", [None, ([70, 71, 72, 73, 74], [0, 1, 0, 0, 1])])
def test_reference_fuse_expectations_on_device(env, name, defaults):
    from oceanbase_b200.compaction import merge_decoded
    ctx, torch = env
    rows, expect = FUSE_CASES[name]
    dv, dn = (None, None) if defaults is None else defaults
    runs = case_runs(rows)
    res = merge_decoded(ctx, [to_dev(torch, r) for r in runs], dv, dn)
    assert_merge_equal(res, ora.major_merge(runs, 5, dv, dn), 5)
    if expect is not None:
        want = expected_row(expect, dv, dn)
        assert [(int(res.fetch(c)[1][0]), int(res.fetch(c)[0][0])) for c in range(5)] == want
    res.free()


@pytest.mark.parametrize("n_runs,window", [(1, 5000), (2, 30000), (3, 20000), (5, 9000), (8, 40000)])
def test_config5_runs_decode_and_merge(ob, env, n_runs, window):
    from oceanbase_b200.compaction import decode_run, merge_decoded
    from oceanbase_b200.synth import make_config5_runs
    ctx, torch = env
    runs = make_config5_runs(n_runs=n_runs, window=window, seed=31, rows_per_block=900)
    dec = []
    for r in runs:
        d = decode_run(ctx, r["table"], 0, 1, [2, 3, 4])
        torch.cuda.synchronize()
        assert np.array_equal(d.key.cpu().numpy(), r["key"])
        assert np.array_equal(d.flag.cpu().numpy(), r["flag"])
        for c in range(3):
            assert np.array_equal(d.ext[c].cpu().numpy(), r["ext"][c])
            assert np.array_equal(d.vals[c].cpu().numpy(), r["vals"][c])
        dec.append(d)
    res = merge_decoded(ctx, dec)
    assert_merge_equal(res, ora.major_merge(runs, 3), 3)
    res.free()


def test_empty_runs_and_no_payload(ob, env):
    from oceanbase_b200.compaction import merge_decoded
    ctx, torch = env
    e = {"key": np.zeros(0, dtype=np.int64), "flag": np.zeros(0, dtype=np.uint8), "vals": [], "ext": []}
    a = {"key": np.arange(0, 5000, 2, dtype=np.int64), "flag": None, "vals": [], "ext": []}
    b = {"key": np.arange(0, 5000, 3, dtype=np.int64), "flag": None, "vals": [], "ext": []}
    for runs in ([e, a, e, b, e], [a], [e, e], [b, a, b, a]):
        res = merge_decoded(ctx, [to_dev(torch, r) for r in runs])
        assert_merge_equal(res, ora.major_merge(runs, 0), 0)
        res.free()


def test_every_key_in_every_run(ob, env):
    # worst case for the fuse: K rows per rowkey, deep NOP chains
    from oceanbase_b200.compaction import merge_decoded
    ctx, torch = env
    rng = np.random.default_rng(5)
    n, K = 20000, 7
    key = np.cumsum(rng.integers(1, 9, size=n)).astype(np.int64)
    runs = []
    for r in range(K):
        ext = [rng.choice(np.array([0, 1, 2], dtype=np.uint8), size=n, p=[0.3, 0.1, 0.6]) for _ in range(4)]
        if r == 0:
            ext = [np.where(e == 2, 0, e).astype(np.uint8) for e in ext]
        vals = [np.where(e == 0, rng.integers(-10 ** 15, 10 ** 15, size=n), 0).astype(np.int64) for e in ext]
        flag = rng.choice(np.array([ob.DF_INSERT, ob.DF_UPDATE, ob.DF_DELETE, ob.DF_NOT_EXIST], dtype=np.uint8), size=n,
                          p=[0.45, 0.45, 0.05, 0.05])
        runs.append({"key": key, "flag": flag, "vals": vals, "ext": ext})
    res = merge_decoded(ctx, [to_dev(torch, r) for r in runs], [1, 2, 3, 4], [0, 0, 1, 0])
    assert_merge_equal(res, ora.major_merge(runs, 4, [1, 2, 3, 4], [0, 0, 1, 0]), 4)
    res.free()


def test_merged_stream_written_as_sstable_scans_back(ob, env):
    # compaction output -> reference-format blocks -> the scan path reads them back bit-exactly
    from oceanbase_b200.compaction import decode_run, merge_decoded, write_merged_sstable
    from oceanbase_b200.synth import make_config5_runs
    ctx, torch = env
    runs = make_config5_runs(n_runs=4, window=25000, seed=41)
    dec = [decode_run(ctx, r["table"], 0, 1, [2, 3, 4]) for r in runs]
    res = merge_decoded(ctx, dec)
    want = ora.major_merge(runs, 3)
    table = write_merged_sstable(res, rows_per_block=1000)
    assert table.total_rows == len(want["key"])
    k, e = ora.decode_column_ext(table, 0)
    assert np.array_equal(k, want["key"]) and not e.any()
    for c in range(3):
        v, e = ora.decode_column_ext(table, 1 + c)
        assert np.array_equal(e, want["null"][c]) and np.array_equal(v, want["vals"][c])
    d2 = decode_run(ctx, table, 0, None, [1, 2, 3])
    torch.cuda.synchronize()
    assert np.array_equal(d2.key.cpu().numpy(), want["key"])
    for c in range(3):
        assert np.array_equal(d2.ext[c].cpu().numpy(), want["null"][c])
    res.free()
