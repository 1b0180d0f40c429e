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


def _string_runs(ob, rng, n_runs, n_keys, cs):
    """Runs with an INT64 rowkey, a DML flag, one INT64 and two VARCHAR payload columns (NULL and NOP cells)."""
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(0, 24)), dtype=np.uint8)) for _ in range(200)]
    runs = []
    for r in range(n_runs):
        keys = np.sort(rng.choice(n_keys, size=int(n_keys * 0.45), replace=False)).astype(np.int64)
        n = len(keys)
        flag = np.where(rng.random(n) < 0.06, ob.DF_DELETE, ob.DF_INSERT if r == 0 else ob.DF_UPDATE).astype(np.int64)
        iv = rng.integers(-1000, 1000, size=n, dtype=np.int64)
        s1 = [words[i] for i in rng.integers(0, 200, size=n)]
        s2 = [words[i] + b"#%d" % r for i in rng.integers(0, 200, size=n)]
        ext = lambda: rng.choice([0, 0, 0, 1, 2] if r else [0, 0, 0, 1], size=n).astype(np.uint8)   # NOP only in increments
        runs.append(dict(key=keys, flag=flag, iv=iv, s1=s1, s2=s2, e_iv=ext(), e_s1=ext(), e_s2=ext()))
    return runs


@pytest.mark.parametrize("block_format", ["pax", "cs"])
def test_merge_with_string_payload_columns(env, ob, block_format):
    # VARCHAR payload cells travel through the merge as references into their run's page batch and are materialised
    # once, for the merged rows only (obgpu_merge_runs + obgpu_merge_result_fetch_strings)
    from oceanbase_b200.compaction import merge_batches, write_merged_sstable
    ctx, torch = env
    rng = np.random.default_rng(77)
    cs = block_format == "cs"
    runs = _string_runs(ob, rng, 5, 30_000, cs)
    batches = []
    for r in runs:
        if cs:   # CS blocks have no NOP: incremental cells that are NOP become NULL for this variant
            for k in ("e_iv", "e_s1", "e_s2"):
                r[k] = np.minimum(r[k], 1).astype(np.uint8)
            enc_i, enc_s1, enc_s2 = ob.ENC_CS_INTEGER, ob.ENC_CS_STRING, ob.ENC_CS_STR_DICT
        else:
            enc_i, enc_s1, enc_s2 = ob.ENC_RAW, ob.ENC_RAW, ob.ENC_DICT
        cols = [ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER if cs else ob.ENC_INTEGER_BASE_DIFF, r["key"]),
                ob.Column(ob.OBJ_INT, enc_i, r["flag"]),
                ob.Column(ob.OBJ_INT, enc_i, r["iv"], nulls=r["e_iv"]),
                ob.Column(ob.OBJ_VARCHAR, enc_s1, r["s1"], nulls=r["e_s1"]),
                ob.Column(ob.OBJ_VARCHAR, enc_s2, r["s2"], nulls=r["e_s2"])]
        batches.append(ctx.open_batch(ob.encode_table(cols, 900, rowkey_cnt=1)))
    res = merge_batches(ctx, batches, 0, 1, [2, 3, 4])
    # model: per rowkey the rows newest first; first non-NOP cell wins, a delete stops the fuse, delete first drops the key
    rows = {}
    for ri, r in enumerate(runs):
        for i, k in enumerate(r["key"].tolist()):
            rows.setdefault(k, []).append((ri, i))
    exp_key, exp = [], [[], [], []]
    for k in sorted(rows):
        cells = [None, None, None]          # None = still NOP
        first, deleted = True, False
        for ri, i in sorted(rows[k], reverse=True):
            r = runs[ri]
            if r["flag"][i] == ob.DF_DELETE:
                deleted = first
                break
            first = False
            for c, (vals, ext) in enumerate(((r["iv"], r["e_iv"]), (r["s1"], r["e_s1"]), (r["s2"], r["e_s2"]))):
                if cells[c] is None and ext[i] != 2:
                    cells[c] = ("null",) if ext[i] == 1 else ("v", vals[i])
        if deleted or first:
            continue
        exp_key.append(k)
        for c in range(3):
            exp[c].append(None if cells[c] in (None, ("null",)) else cells[c][1])   # open NOPs take the NULL default
    assert res.info().out_rows == len(exp_key)
    assert np.array_equal(res.fetch(-1)[0], np.array(exp_key, dtype=np.int64))
    v, nl = res.fetch(0)
    assert [None if z else int(x) for x, z in zip(v, nl)] == exp[0]
    with pytest.raises(ob.ObGpuError):
        res.fetch_strings(0)                                         # an integer column holds no references
    for c in (1, 2):
        heap, off, nl = res.fetch_strings(c)
        got = [None if nl[i] else bytes(heap[off[i]:off[i + 1]]) for i in range(len(exp_key))]
        assert got == exp[c], f"string column {c}"
        h2, o2, n2 = res.fetch_strings(c, 1000, 500)                 # a window
        assert [None if n2[i] else bytes(h2[o2[i]:o2[i + 1]]) for i in range(500)] == exp[c][1000:1500]
    # the merged stream as a new SSTable, read back through the oracle
    out = write_merged_sstable(res, 1000, string_cols=(1, 2))
    blk = ora.Block(out.block(3))
    for r in range(0, blk.row_count, 37):
        g = 3000 + r
        assert blk.cell(0, r) == exp_key[g] % (1 << 64) and blk.cell(2, r) == exp[1][g] and blk.cell(3, r) == exp[2][g]
    res.free()
    for b in batches:
        b.close()


def test_decode_columns_tagged_and_named_images(env, ob):
    # the two-step form: decode runs into caller-owned arrays (string cells = tagged references), merge_decoded, then name
    # the images the tags stand for
    import ctypes as C
    from oceanbase_b200.capi import lib, check
    from oceanbase_b200.compaction import DecodedRun, merge_decoded
    ctx, torch = env
    rng = np.random.default_rng(5)
    runs, images, sizes, keep = [], [], [], []
    for tag in range(3):
        n = 4000
        key = (np.arange(n, dtype=np.int64) * 3 + tag)                   # disjoint rowkeys: nothing fuses
        s = [b"r%d-" % tag + bytes(rng.integers(97, 123, size=int(rng.integers(0, 9)), dtype=np.uint8)) for _ in range(n)]
        nulls = (rng.random(n) < 0.1).astype(np.uint8)
        table = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_INTEGER_BASE_DIFF, key), ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, s, nulls=nulls)], 700)
        d_img = torch.zeros(table.image.size + 64, dtype=torch.uint8, device="cuda")
        d_img[:table.image.size].copy_(torch.from_numpy(table.image))
        torch.cuda.synchronize()
        batch = ctx.open_batch(table, device_image_ptr=d_img.data_ptr())
        vs = [torch.empty(n, dtype=torch.int64, device="cuda") for _ in range(2)]
        es = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
        ci = (C.c_int32 * 2)(0, 1)
        vp = (C.c_void_p * 2)(*[v.data_ptr() for v in vs])
        ep = (C.c_void_p * 2)(*[e.data_ptr() for e in es])
        check(lib.obgpu_batch_decode_columns_tagged(batch._h, 2, ci, tag, vp, ep), "obgpu_batch_decode_columns_tagged", ctx._h)
        refs = vs[1].cpu().numpy().view(np.uint64)
        live = nulls == 0
        assert np.all((refs[live] >> np.uint64(58)) == tag) and np.all(refs[~live] == 0)
        assert np.array_equal((refs[live] & np.uint64((1 << 22) - 1)).astype(np.int64), np.array([len(x) for x, z in zip(s, nulls) if not z]))
        runs.append(DecodedRun(vs[0], None, [vs[1]], [es[1]]))
        images.append(d_img.data_ptr())
        sizes.append(table.image.size)
        keep += [d_img, batch, s, nulls]
    res = merge_decoded(ctx, runs)
    res.set_string_images(images, [8] * 3)                      # images declared too small: references are refused, not followed
    with pytest.raises(ob.ObGpuError):
        res.fetch_strings(0)
    res.set_string_images(images, sizes)
    heap, off, nl = res.fetch_strings(0)
    k, _ = res.fetch(-1)
    for i in range(0, len(k), 53):
        tag, idx = int(k[i]) % 3, int(k[i]) // 3
        s, nulls = keep[tag * 4 + 2], keep[tag * 4 + 3]
        assert (None if nl[i] else bytes(heap[off[i]:off[i + 1]])) == (None if nulls[idx] else s[idx])
    res.free()
    for t in range(3):
        keep[t * 4 + 1].close()


@pytest.mark.parametrize("n_more", [1, 2, 3])
def test_composite_rowkeys_on_device(env, n_more):
    from test_major_merge_kat import composite_runs
    from oceanbase_b200.compaction import DecodedRun, merge_decoded
    ctx, torch = env
    rng = np.random.default_rng(90 + n_more)
    runs = composite_runs(rng, 7, 40_000, n_more)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    dev = [DecodedRun(t(r["key"], np.int64), t(r["flag"], np.uint8), [t(r["vals"][0], np.int64)], [t(r["ext"][0], np.uint8)],
                      [t(k, np.int64) for k in r["more_keys"]]) for r in runs]
    res = merge_decoded(ctx, dev)
    want = ora.major_merge(runs, 1)
    assert_merge_equal(res, want, 1)
    for c in range(n_more):
        assert np.array_equal(res.fetch(-2 - c)[0], want["more_keys"][c]), f"rowkey column {1 + c}"
    res.free()


def test_composite_rowkeys_from_page_batches(env, ob):
    # obgpu_merge_runs_keys: (tenant, order) rowkey decoded from the runs' SSTables, VARCHAR payload along
    from test_major_merge_kat import composite_runs
    from oceanbase_b200.compaction import merge_batches
    ctx, torch = env
    rng = np.random.default_rng(123)
    runs = composite_runs(rng, 4, 30_000, 1)
    batches = []
    for ri, r in enumerate(runs):
        n = len(r["key"])
        r["s"] = [b"t%d-%d-%d" % (ri, int(a), int(b)) for a, b in zip(r["key"], r["more_keys"][0])]
        cols = [ob.Column(ob.OBJ_INT, ob.ENC_RLE, r["key"]), ob.Column(ob.OBJ_INT, ob.ENC_RAW, r["more_keys"][0]),
                ob.Column(ob.OBJ_INT, ob.ENC_RAW, r["flag"].astype(np.int64)),
                ob.Column(ob.OBJ_INT, ob.ENC_RAW, r["vals"][0], nulls=r["ext"][0]), ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, r["s"])]
        batches.append(ctx.open_batch(ob.encode_table(cols, 800, rowkey_cnt=2)))
    res = merge_batches(ctx, batches, [0, 1], 2, [3, 4])
    want = ora.major_merge(runs, 1)
    assert_merge_equal(res, want, 1)
    assert np.array_equal(res.fetch(-2)[0], want["more_keys"][0])
    heap, off, nl = res.fetch_strings(1)
    k0, k1 = res.fetch(-1)[0], res.fetch(-2)[0]
    for i in range(0, len(k0), 41):   # the newest existing row of the group wrote the string
        s = bytes(heap[off[i]:off[i + 1]])
        assert not nl[i] and s.startswith(b"t") and s.endswith(b"-%d-%d" % (int(k0[i]), int(k1[i])))
    res.free()
    for b in batches:
        b.close()


def test_distributed_entry_on_one_rank(ob, env):
    """obgpu_merge_decoded_distributed with a one-rank communicator (NCCL bound inside the library): no splitters, every
    run stays local, the result is the plain merge. The multi-rank exchange is exercised by tools/bench_compaction.py
    --verify under torchrun (gpurun --gpus N)."""
    from oceanbase_b200.compaction import Comm, merge_decoded_distributed
    from oceanbase_b200.synth import make_config5_runs
    ctx, torch = env
    runs = make_config5_runs(n_runs=5, window=30_000, seed=9, encode=False)
    comm = Comm(ctx, Comm.unique_id(), 0, 1)
    dec = {q: to_dev(torch, runs[q]) for q in range(5)}
    res, split, recv = merge_decoded_distributed(ctx, comm, dec, 5, 3)
    assert len(split) == 0 and list(recv) == [len(r["key"]) for r in runs]
    assert_merge_equal(res, ora.major_merge(runs, 3), 3)
    res.free()
    comm.close()


def test_streamed_merge_of_runs_that_do_not_fit_together():
    """Range by range (host -> device copy, decode, cut, merge, fetch; two ranges in flight): the concatenated ranges equal the
    one-shot merge of the same runs."""
    import torch
    import oceanbase_b200 as ob
    from oceanbase_b200.synth import make_config5_runs
    from oceanbase_b200.compaction import decode_run, merge_decoded, streamed_major_merge
    runs = make_config5_runs(n_runs=5, window=60_000, seed=9, rows_per_block=700)
    ctx = ob.ScanContext(0)
    dec = [decode_run(ctx, r["table"], 0, 1, [2, 3, 4]) for r in runs]
    whole = merge_decoded(ctx, dec)
    n = whole.info().out_rows
    want = [whole.fetch(c) for c in (-1, 0, 1, 2)]
    end_keys = [r["key"][np.minimum(np.arange(699, len(r["key"]) + 699, 700), len(r["key"]) - 1)] for r in runs]
    for q, r in enumerate(runs):
        assert len(end_keys[q]) == r["table"].n_blocks
    got = {c: [] for c in (-1, 0, 1, 2)}
    order, stats = [], [0, 0]

    def sink(i, res):
        order.append(i)
        info = res.info()
        stats[0] += info.dropped_deletes
        stats[1] += info.fused_rows
        for c in got:
            got[c].append(res.fetch(c))

    n_ranges = streamed_major_merge([r["table"] for r in runs], end_keys, 0, 1, [2, 3, 4], 7, sink, device=torch.device("cuda", 0))
    assert order == list(range(n_ranges)) and n_ranges >= 5
    py_side = ({c: [(v.copy(), nl.copy()) for v, nl in got[c]] for c in got}, list(stats))
    # the same through the library's own range loop (obgpu_merge_runs_streamed: worker threads, one stream each)
    from oceanbase_b200.compaction import merge_runs_streamed
    got = {c: [] for c in (-1, 0, 1, 2)}
    order.clear()
    stats[0] = stats[1] = 0
    n_c = merge_runs_streamed(0, [r["table"] for r in runs], end_keys, 0, 1, [2, 3, 4], 7, sink, n_streams=3)
    assert n_c == n_ranges and order == list(range(n_ranges))
    assert stats == py_side[1]
    for c in got:
        assert np.array_equal(np.concatenate([x[0] for x in got[c]]), np.concatenate([x[0] for x in py_side[0][c]]))
    for k, c in enumerate((-1, 0, 1, 2)):
        v = np.concatenate([x[0] for x in got[c]])
        nl = np.concatenate([x[1] for x in got[c]])
        assert len(v) == n
        assert np.array_equal(nl, want[k][1])
        assert np.array_equal(v[nl == 0], want[k][0][want[k][1] == 0])
    assert stats[0] == whole.info().dropped_deletes and stats[1] == whole.info().fused_rows
    whole.free()
    ctx.close()


@pytest.mark.parametrize("k_runs,seed", [(2, 1), (3, 2), (5, 3), (8, 4), (13, 5), (33, 6), (64, 7)])
def test_single_pass_merge_shapes(env, k_runs, seed, monkeypatch):
    """The bucket merge (single-column rowkeys) over run counts that are not powers of two, runs of very different sizes, empty runs,
    rowkeys shared by every run, dense and sparse key ranges; every case also through the pairwise passes (OBGPU_MERGE_PAIRWISE)."""
    from oceanbase_b200.compaction import merge_decoded
    ctx, torch = env
    rng = np.random.default_rng(seed)
    runs = []
    shared = np.sort(rng.choice(1 << 20, size=300, replace=False)).astype(np.int64) * 7
    for r in range(k_runs):
        n = int(rng.choice([0, 1, 17, 900, 5000, 40000]))
        span = int(rng.choice([1 << 12, 1 << 22, 1 << 40]))
        key = np.unique(rng.integers(-span, span, size=n, dtype=np.int64))
        if r % 3 == 0 and n > 0:
            key = np.unique(np.concatenate([key, shared]))      # the same rowkeys in many runs: long groups
        m = len(key)
        flag = rng.choice([0, 2, 3, 3, 3, 4], size=m).astype(np.uint8)      # NOT_EXIST / UPDATE / INSERT / DELETE
        vals = [rng.integers(-(1 << 50), 1 << 50, size=m, dtype=np.int64) for _ in range(2)]
        ext = [rng.choice([0, 0, 0, 1, 2], size=m).astype(np.uint8) for _ in range(2)]
        for c in range(2):
            vals[c][ext[c] != 0] = 0
        runs.append({"key": key, "flag": flag, "vals": vals, "ext": ext})
    want = ora.major_merge(runs, 2, [5, 6], [0, 1])
    for pairwise in (False, True):
        if pairwise:
            monkeypatch.setenv("OBGPU_MERGE_PAIRWISE", "1")
        else:
            monkeypatch.delenv("OBGPU_MERGE_PAIRWISE", raising=False)
        res = merge_decoded(ctx, [to_dev(torch, r) for r in runs], [5, 6], [0, 1])
        assert_merge_equal(res, want, 2)
        res.free()
