"""CS_ENCODING_ROW_STORE blocks on the device path vs the oracle: every ObIntegerStreamMeta shape through the
whole-table scan (filter + projection), the per-block entry points, and the Q6 shape of BASELINE config 4 with
the pushed-down SUM(l_extendedprice * l_discount) checked exactly against Python integers."""
import numpy as np
import pytest

import oracle_binding as ora
from test_gpu_scan import assert_scan_matches
from test_cs_encoding_kat import SHAPES

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


class W:
    def __init__(self, table, flt, proj, is_str, elem):
        self.table, self.filter, self.proj, self.proj_is_string, self.proj_elem_len = table, flt, proj, is_str, elem


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("with_nulls", [False, True])
def test_cs_integer_columns_scan(ob, ctx, shape, with_nulls):
    obj_type, gen, _ = SHAPES[shape]
    rng = np.random.default_rng(3)
    n = 7000
    v = gen(n).astype(np.int64) if obj_type != ob.OBJ_UINT64 else gen(n).astype(np.uint64).view(np.int64)
    nulls = (rng.random(n) < 0.15).astype(np.uint8) if with_nulls else None
    sel = rng.integers(0, 1000, size=n, dtype=np.int64)
    table = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, sel), ob.Column(obj_type, ob.ENC_CS_INTEGER, v, nulls=nulls)], 900)
    elem = 4 if obj_type == ob.OBJ_DATE else 8
    mid = int(np.median(v)) if obj_type != ob.OBJ_UINT64 else int(np.median(v.view(np.uint64)))
    for flt in (None, ob.White(0, ob.WHITE_OP_LT, (300,)), ob.White(1, ob.WHITE_OP_GE, (mid,)),
                ob.And([ob.White(0, ob.WHITE_OP_GE, (100,)), ob.White(1, ob.WHITE_OP_LT, (mid,))]),
                ob.Or([ob.White(1, ob.WHITE_OP_NU, ()), ob.White(0, ob.WHITE_OP_EQ, (5,))])):
        assert_scan_matches(ctx, W(table, flt, [0, 1], [False, False], [8, elem]))


def test_cs_block_entry_points(ob, ctx):
    n = 300
    rng = np.random.default_rng(8)
    v = rng.integers(-1000, 1000, size=n, dtype=np.int64)
    nulls = (rng.random(n) < 0.2).astype(np.uint8)
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.arange(n, dtype=np.int64)),
                             ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, v, nulls=nulls)])
    blk = ora.Block(block)
    table = ob.TableImage(np.concatenate([block, np.zeros((-len(block)) % 128 + 128, dtype=np.uint8)]),
                          np.array([0], dtype=np.int64), np.array([len(block)], dtype=np.int64), 0, 0)
    batch = ctx.open_batch(table)
    for op, params in ((ob.WHITE_OP_LT, (0,)), (ob.WHITE_OP_BT, (-10, 500)), (ob.WHITE_OP_NU, ()), (ob.WHITE_OP_NE, (int(v[3]),))):
        for start, count in ((0, None), (17, 200)):
            assert np.array_equal(batch.filter_white(0, 1, op, params, start, count),
                                  blk.filter_tree(ob.White(1, op, params), start, count))
    rid = np.arange(0, n, 3, dtype=np.int32)
    ed, en, ehn = blk.get_rows_fixed(1, rid, 8, 4)
    gd, gn, ghn = batch.project_fixed(0, 1, rid, 8, 4)
    assert np.array_equal(gd, ed) and np.array_equal(gn, en) and ghn == ehn
    batch.close()


def test_q6_scan_and_pushdown_sum(ob, ctx):
    from oceanbase_b200.synth import make_config4_like
    w = make_config4_like(rows=400_000, rows_per_block=2000, seed=4)
    n = assert_scan_matches(ctx, w)
    assert 0.012 < n / w.table.total_rows < 0.026          # Q6 selects ~1.9 %
    batch = ctx.open_batch(w.table)
    res = batch.scan(w.filter, w.proj)
    want = ora.scan_table(w.table, w.filter, w.proj, w.proj_is_string, w.proj_elem_len)
    price = want["data"][0].view(np.int64)[:n]
    disc = want["data"][1].view(np.int64)[:n]
    exact = sum(int(a) * int(b) for a, b in zip(price.tolist(), disc.tolist()))
    assert res.aggregate(ob.AGG_SUM_PRODUCT, 0, 1) == exact
    assert res.aggregate(ob.AGG_SUM, 0) == int(price.astype(object).sum())
    assert res.aggregate(ob.AGG_COUNT, 1) == n
    assert res.aggregate(ob.AGG_MIN, 0) == int(price.min()) and res.aggregate(ob.AGG_MAX, 0) == int(price.max())
    res.free()
    batch.close()


def test_aggregates_with_nulls_and_wide_values(ob, ctx):
    rng = np.random.default_rng(12)
    n = 50_000
    a = rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64)
    b = rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64)
    na = (rng.random(n) < 0.1).astype(np.uint8)
    nb = (rng.random(n) < 0.1).astype(np.uint8)
    u = rng.integers(0, 1 << 63, size=n, dtype=np.int64) * 2 + 1          # uint64 values above 2^63
    table = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_RAW, a, nulls=na), ob.Column(ob.OBJ_INT, ob.ENC_RAW, b, nulls=nb),
                             ob.Column(ob.OBJ_UINT64, ob.ENC_RAW, u)], 1000)
    batch = ctx.open_batch(table)
    res = batch.scan(None, [0, 1, 2])
    ok = (na == 0)
    both = ok & (nb == 0)
    M = 1 << 128
    exact_sum = sum(int(x) for x in a[ok].tolist())
    exact_prod = sum(int(x) * int(y) for x, y in zip(a[both].tolist(), b[both].tolist()))
    assert res.aggregate(ob.AGG_SUM, 0) % M == exact_sum % M               # 128-bit two's complement
    assert res.aggregate(ob.AGG_SUM_PRODUCT, 0, 1) % M == exact_prod % M
    assert res.aggregate(ob.AGG_COUNT, 0) == int(ok.sum())
    assert res.aggregate(ob.AGG_MIN, 0) == int(a[ok].min()) and res.aggregate(ob.AGG_MAX, 0) == int(a[ok].max())
    uu = u.view(np.uint64)
    assert res.aggregate(ob.AGG_SUM, 2) == sum(int(x) for x in uu.tolist())
    assert res.aggregate(ob.AGG_MAX, 2) == int(uu.max()) and res.aggregate(ob.AGG_MIN, 2) == int(uu.min())
    res.free()
    none = batch.scan(ob.White(0, ob.WHITE_OP_NU, ()), [0])
    assert none.aggregate(ob.AGG_MIN, 0) is None and none.aggregate(ob.AGG_COUNT, 0) == 0 and none.aggregate(ob.AGG_SUM, 0) == 0
    none.free()
    batch.close()


@pytest.mark.parametrize("with_nulls", [False, True])
def test_cs_int_dict_columns_scan(ob, ctx, with_nulls):
    rng = np.random.default_rng(21)
    n = 9000
    a = rng.integers(-40, 40, size=n, dtype=np.int64) * 1000
    b = rng.integers(8000, 8200, size=n, dtype=np.int64)
    na = (rng.random(n) < 0.2).astype(np.uint8) if with_nulls else None
    allnull = np.ones(n, dtype=np.uint8)
    table = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_CS_INT_DICT, a, nulls=na),
                             ob.Column(ob.OBJ_DATE, ob.ENC_CS_INT_DICT, b),
                             ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, rng.integers(0, 1 << 40, size=n, dtype=np.int64)),
                             ob.Column(ob.OBJ_INT, ob.ENC_CS_INT_DICT, np.zeros(n, dtype=np.int64), nulls=allnull)], 1100)
    flts = [None, ob.White(0, ob.WHITE_OP_GE, (0,)), ob.White(1, ob.WHITE_OP_IN, (8001, 8100, 9999)),
            ob.And([ob.White(0, ob.WHITE_OP_NE, (5000,)), ob.White(1, ob.WHITE_OP_BT, (8050, 8150)), ob.White(2, ob.WHITE_OP_LT, (1 << 39,))]),
            ob.Or([ob.White(0, ob.WHITE_OP_NU, ()), ob.White(3, ob.WHITE_OP_NN, ())])]
    for flt in flts:
        assert_scan_matches(ctx, W(table, flt, [0, 1, 2, 3], [False] * 4, [8, 4, 8, 8]))


@pytest.mark.parametrize("shape", ["var", "var_with_empty", "fixed", "all_empty"])
@pytest.mark.parametrize("null_frac", [0.0, 0.2, 1.0])
def test_cs_string_columns_scan(ob, ctx, shape, null_frac):
    # STRING (fixed / END offsets, NULL bitmap / zero length) and STR_DICT columns: bytes in the all-string-data area
    from test_cs_encoding_kat import STR_SHAPES
    rng = np.random.default_rng(31)
    n = 6000
    v = STR_SHAPES[shape][0](rng, n)
    nulls = (rng.random(n) < null_frac).astype(np.uint8) if null_frac else None
    k = rng.integers(0, 1000, size=n, dtype=np.int64)
    table = ob.encode_table([ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STRING, v, nulls=nulls),
                             ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, k),
                             ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STR_DICT, v, nulls=nulls),
                             ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STRING, v[::-1])], 800)
    probe = v[5]
    flts = [None, ob.White(0, ob.WHITE_OP_GE, (b"m",)), ob.White(2, ob.WHITE_OP_LT, (probe,)),
            ob.And([ob.White(1, ob.WHITE_OP_LT, (500,)), ob.White(2, ob.WHITE_OP_IN, (probe, v[77], b"zz")), ob.White(3, ob.WHITE_OP_NE, (probe,))]),
            ob.Or([ob.White(0, ob.WHITE_OP_NU, ()), ob.White(2, ob.WHITE_OP_EQ, (probe,)), ob.White(3, ob.WHITE_OP_BT, (b"c", b"f"))])]
    for flt in flts:
        assert_scan_matches(ctx, W(table, flt, [0, 1, 2, 3], [True, False, True, True], [8, 8, 8, 8]))
    assert_scan_matches(ctx, W(table, ob.White(1, ob.WHITE_OP_LT, (20,)), [2, 0], [True, True], [8, 8]))   # sparse selection


def test_cs_string_block_entry_points_and_bytes(ob, ctx):
    from test_cs_encoding_kat import _strings
    rng = np.random.default_rng(32)
    n = 500
    v = _strings(rng, n, 0, 14, 45)
    nulls = (rng.random(n) < 0.25).astype(np.uint8)
    block = ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STRING, v, nulls=nulls),
                             ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STR_DICT, v, nulls=nulls)])
    blk = ora.Block(block)
    image = np.concatenate([block, np.zeros((-len(block)) % 128 + 128, dtype=np.uint8)])
    table = ob.TableImage(image, np.array([0], dtype=np.int64), np.array([len(block)], dtype=np.int64), 0, 0)
    batch = ctx.open_batch(table)
    for col in (0, 1):
        for op, params in ((ob.WHITE_OP_LE, (v[1],)), (ob.WHITE_OP_NU, ()), (ob.WHITE_OP_IN, (v[2], v[3])), (ob.WHITE_OP_NE, (v[4],))):
            assert np.array_equal(batch.filter_white(0, col, op, params, 0, None), blk.filter_tree(ob.White(col, op, params)))
    base = 0x20_0000_0000
    res = batch.scan(None, [0, 1], string_base=base)
    for c in (0, 1):
        data, lens, nl = res.fetch_col(c)
        for r in range(n):
            is_null = (int(nl[r >> 6]) >> (r & 63)) & 1
            assert bool(is_null) == bool(nulls[r])
            if not is_null:
                off = int(data[r]) - base
                assert bytes(image[off:off + int(lens[r])]) == v[r]
    res.free()
    batch.close()


@pytest.mark.parametrize("kind", ["int", "str"])
def test_cs_dict_const_encoded_refs_scan(ob, ctx, kind):
    # a dominant value (or NULL) per block: the encoder's const-encoded ref stream -> K_CONST plan over the dictionary
    rng = np.random.default_rng(41)
    n = 12_000
    idx = np.zeros(n, dtype=np.int64)
    where = rng.choice(n, size=n // 40, replace=False)            # 2.5 % exceptions
    idx[where] = rng.integers(1, 9, size=len(where))
    nulls = np.zeros(n, dtype=np.uint8)
    nulls[6000:] = 1                                              # second half: NULL is the const ref
    nulls[where] = 0
    idx[9000:9600] = 0                                            # one block with a single distinct ref (all NULL)
    nulls[9000:9600] = 1
    if kind == "int":
        col = ob.Column(ob.OBJ_INT, ob.ENC_CS_INT_DICT, (idx * 1000 - 3000).astype(np.int64), nulls=nulls)
        probe, is_str = 1000, False
    else:
        words = [b"", b"alpha", b"be", b"gamma", b"delta!", b"e", b"zeta", b"eta", b"theta"]
        col = ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STR_DICT, [words[i] for i in idx.tolist()], nulls=nulls)
        probe, is_str = b"gamma", True
    table = ob.encode_table([col, ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.arange(n, dtype=np.int64))], 600)
    for flt in (None, ob.White(0, ob.WHITE_OP_EQ, (probe,)), ob.White(0, ob.WHITE_OP_NE, (probe,)), ob.White(0, ob.WHITE_OP_NU, ()),
                ob.And([ob.White(0, ob.WHITE_OP_GE, (probe,)), ob.White(1, ob.WHITE_OP_LT, (8000,))]),
                ob.Or([ob.White(0, ob.WHITE_OP_NN, ()), ob.White(1, ob.WHITE_OP_EQ, (9100,))])):
        assert_scan_matches(ctx, W(table, flt, [0, 1], [is_str, False], [8, 8]))


@pytest.mark.parametrize("name", ["integer", "integer_nulls", "uint", "uint_nulls", "int_dict_const", "str_dict_const", "varchar", "varchar_nulls"])
def test_reference_cs_filter_expectations_on_device(ob, ctx, name):
    # the reference's own CS pd-filter unit-test datasets and expected counts (tests/test_cs_reference_filter_kat.py)
    from test_cs_reference_filter_kat import DATASETS, OPS, build
    for enc in DATASETS[name][1]:
        block, cases = build(name, enc)
        image = np.concatenate([block, np.zeros((-len(block)) % 128 + 128, dtype=np.uint8)])
        table = ob.TableImage(image, np.array([0], dtype=np.int64), np.array([len(block)], dtype=np.int64), 0, 0)
        batch = ctx.open_batch(table)
        for op, params, expect in cases:
            assert int(batch.filter_white(0, 1, OPS[op], params).sum()) == expect, (name, enc, op, params)
            res = batch.scan(ob.White(1, OPS[op], params), [0])
            assert res.selected_rows == expect
            res.free()
        batch.close()
