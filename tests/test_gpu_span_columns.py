"""COLUMN_EQUAL / COLUMN_SUBSTR on the device (mat_codecs.cuh: rebuilt once per page batch at open, from the referenced column and the
exception list) vs the oracle: per-block white filters and projection, whole-table scans with span columns filtered AND projected
next to their referenced columns, integer spans through aggregates, page batches opened from the host and from a device image."""
import numpy as np
import pytest

import oracle_binding as ora
from test_span_columns import EQ, SUB
from test_gpu_string_codecs import heap_strings

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import oceanbase_b200
    return oceanbase_b200


@pytest.fixture(scope="module")
def ctx(ob):
    c = ob.ScanContext(0)
    yield c
    c.close()


def has_nop(*nulls):
    return any(n is not None and bool(np.any(n == 2)) for n in nulls)


@pytest.mark.parametrize("case", range(len(EQ)), ids=[c[0] for c in EQ])
def test_column_equal_blocks(ob, ctx, case):
    name, ot, renc, rv, rn, v, nl = EQ[case]
    n = len(v)
    is_str = ot == ob.OBJ_VARCHAR
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)),
            ob.Column(ot, ob.ENC_COLUMN_EQUAL, v, nulls=nl, ref_col=2),
            ob.Column(ot, renc, rv, nulls=rn)]
    table = ob.encode_table(cols, 170)
    batch = ctx.open_batch(table)
    assert bool(batch.column_materialised(1)) == is_str and not batch.column_materialised(2)
    isnull = np.zeros(n, dtype=bool) if nl is None else nl.astype(bool)
    present = sorted({(x if is_str else int(x)) for x, z in zip(v, isnull) if not z})
    lo, hi = present[len(present) // 4], present[3 * len(present) // 4]
    flts = [ob.White(1, ob.WHITE_OP_EQ, (present[2 % len(present)],)), ob.White(1, ob.WHITE_OP_NE, (present[0],)), ob.White(1, ob.WHITE_OP_GE, (lo,)),
            ob.White(1, ob.WHITE_OP_BT, (lo, hi)), ob.White(1, ob.WHITE_OP_IN, (present[0], present[-1])),
            ob.And([ob.White(0, ob.WHITE_OP_LT, (n - 30,)), ob.Or([ob.White(1, ob.WHITE_OP_LT, (lo,)), ob.White(2, ob.WHITE_OP_GT, (hi,))])])]
    if not has_nop(nl, rn):
        flts += [ob.White(1, ob.WHITE_OP_NU, ()), ob.White(1, ob.WHITE_OP_NN, ())]
    row0 = 0
    for b in range(table.n_blocks):
        blk = ora.Block(table.block(b))
        rows = blk.row_count
        for flt in flts:
            for start, count in ((0, None), (5, rows - 11)):
                assert np.array_equal(batch.filter_tree(b, flt, start, count), blk.filter_tree(flt, start, count)), (name, b, flt)
        rid = np.concatenate([np.arange(0, rows, 2), np.arange(rows - 1, 0, -7)]).astype(np.int32)
        if is_str:
            heap, off, nb = batch.project_strings(b, 1, rid)
            assert heap_strings(heap, off, nb) == [None if isnull[row0 + r] else v[row0 + r] for r in rid], (name, b)
        else:
            data, nb, _ = batch.project_fixed(b, 1, rid)
            want, wn, _ = blk.get_rows_fixed(1, rid)
            assert np.array_equal(np.asarray(nb).view(np.uint64)[:len(wn)], wn), (name, b)
            live = ~isnull[row0 + rid]
            assert np.array_equal(np.asarray(data).view(np.int64)[live], want.view(np.int64)[live]), (name, b)
        row0 += rows
    batch.close()
    ora.arena_reset()


@pytest.mark.parametrize("case", range(len(SUB)), ids=[c[0] for c in SUB])
def test_column_substr_blocks(ob, ctx, case):
    name, rv, rn, v, nl = SUB[case]
    n = len(v)
    cols = [ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT if case % 2 else ob.ENC_RAW, rv, nulls=rn),
            ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)),
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_COLUMN_SUBSTR, v, nulls=nl, ref_col=0)]
    table = ob.encode_table(cols, 190)
    batch = ctx.open_batch(table)
    assert batch.column_materialised(2) and not batch.column_materialised(0)
    isnull = np.zeros(n, dtype=bool) if nl is None else nl.astype(bool)
    present = sorted({x for x, z in zip(v, isnull) if not z})
    lo, hi = present[len(present) // 4], present[3 * len(present) // 4]
    flts = [ob.White(2, ob.WHITE_OP_EQ, (present[len(present) // 2],)), ob.White(2, ob.WHITE_OP_LT, (hi,)), ob.White(2, ob.WHITE_OP_BT, (lo, hi)),
            ob.White(2, ob.WHITE_OP_IN, (present[0], present[-1], b"nope")),
            ob.Or([ob.White(2, ob.WHITE_OP_GE, (hi,)), ob.White(0, ob.WHITE_OP_LT, (rv[0],))])]
    if not has_nop(nl, rn):
        flts += [ob.White(2, ob.WHITE_OP_NU, ()), ob.White(2, ob.WHITE_OP_NN, ())]
    row0 = 0
    for b in range(table.n_blocks):
        blk = ora.Block(table.block(b))
        rows = blk.row_count
        for flt in flts:
            assert np.array_equal(batch.filter_tree(b, flt, 0, None), blk.filter_tree(flt, 0, None)), (name, b, flt)
        rid = np.arange(rows - 1, -1, -3).astype(np.int32)
        heap, off, nb = batch.project_strings(b, 2, rid)
        assert heap_strings(heap, off, nb) == [None if isnull[row0 + r] else v[row0 + r] for r in rid], (name, b)
        row0 += rows
    batch.close()
    ora.arena_reset()


@pytest.mark.parametrize("on_device", [False, True])
def test_scan_over_span_columns(ob, ctx, on_device):
    """A table with an integer COLUMN_EQUAL, a string COLUMN_EQUAL and a COLUMN_SUBSTR column (and their referenced columns under
    DICT / RLE / RAW): filters on the span columns, all of them projected, pushed-down aggregates over the integer span column."""
    import torch
    rng = np.random.default_rng(33)
    n = 60000
    k = np.arange(n, dtype=np.int64)
    price = np.sort(rng.integers(1, 5000, size=n)).astype(np.int64)
    paid = price.copy()
    ex = rng.choice(n, n // 40, replace=False)
    paid[ex] = price[ex] - rng.integers(1, 100, size=len(ex))
    pn = np.zeros(n, dtype=np.uint8); pn[ex[:200]] = 1
    city = [[b"hangzhou", b"beijing", b"shanghai", b"shenzhen"][i] for i in rng.integers(0, 4, size=n)]
    ship = list(city)
    for i in ex[::3]: ship[i] = b"elsewhere-%d" % (i % 50)
    mail = [b"%s.user%05d@corp.example" % (city[i][:3], i % 9973) for i in range(n)]
    user = [m[4:13] for m in mail]
    for i in ex[1::5]: user[i] = b"anonymous"
    un = np.zeros(n, dtype=np.uint8); un[ex[2::7]] = 1
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_INTEGER_BASE_DIFF, k),
            ob.Column(ob.OBJ_INT, ob.ENC_RLE, price), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, paid, nulls=pn, ref_col=1),
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, city), ob.Column(ob.OBJ_VARCHAR, ob.ENC_COLUMN_EQUAL, ship, ref_col=3),
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, mail), ob.Column(ob.OBJ_VARCHAR, ob.ENC_COLUMN_SUBSTR, user, nulls=un, ref_col=5)]
    table = ob.encode_table(cols, 1500)
    base = table.image.ctypes.data
    if on_device:
        d = torch.empty(table.image.size + 64, dtype=torch.uint8, device="cuda:0")
        d[:table.image.size].copy_(torch.from_numpy(table.image))
        d[table.image.size:].zero_()
        torch.cuda.synchronize()
        batch = ctx.open_batch(table, device_image_ptr=d.data_ptr(), host_view=False, image_size=table.image.size)
    else:
        batch = ctx.open_batch(table)
    assert [bool(batch.column_materialised(c)) for c in range(7)] == [False, False, False, False, True, False, True]
    proj, is_str = [0, 2, 4, 6, 1], [False, False, True, True, False]
    for flt in (None, ob.White(2, ob.WHITE_OP_BT, (1000, 1200)),
                ob.And([ob.White(4, ob.WHITE_OP_EQ, (b"beijing",)), ob.White(6, ob.WHITE_OP_GE, (b"user05000",))]),
                ob.Or([ob.White(2, ob.WHITE_OP_NU, ()), ob.White(6, ob.WHITE_OP_EQ, (b"anonymous",)), ob.White(4, ob.WHITE_OP_GT, (b"f",))]),
                ob.And([ob.White(6, ob.WHITE_OP_NN, ()), ob.White(2, ob.WHITE_OP_LT, (40,)), ob.White(1, ob.WHITE_OP_LT, (45,))])):
        want = ora.scan_table(table, flt, proj, is_str, [8] * 5, string_base=base)
        res = batch.scan(flt, proj, string_base=base)
        assert res.selected_rows == want["selected"], flt
        for p in (0, 1, 4):
            data, _, nb = res.fetch_col(p)
            wn = want["nulls"][p]
            assert np.array_equal(nb[:len(wn)], wn), (flt, p)
            live = np.array([not ((int(wn[i // 64]) >> (i % 64)) & 1) for i in range(want["selected"])], dtype=bool)
            assert np.array_equal(np.asarray(data).view(np.int64)[live], want["data"][p].view(np.int64)[live]), (flt, p)
        for p in (2, 3):
            _, lens, nb = res.fetch_col(p)
            assert np.array_equal(nb[:len(want["nulls"][p])], want["nulls"][p]) and np.array_equal(lens, want["lens"][p]), (flt, p)
            heap, off = res.fetch_strings(p)
            assert heap_strings(heap, off, nb) == ora.scan_strings(table, want, p, base), (flt, p)
        rows = np.asarray(want["data"][0]).view(np.int64)
        live = pn[rows] == 0
        assert res.aggregate(ob.AGG_SUM, 1) == int(paid[rows][live].sum())
        assert res.aggregate(ob.AGG_COUNT, 1) == int(live.sum())
        if live.any():
            assert res.aggregate(ob.AGG_MAX, 1) == int(paid[rows][live].max())
        res.free()
        ora.arena_reset()
    batch.close()
