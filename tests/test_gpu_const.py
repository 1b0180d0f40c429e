"""CONST codec on the device path vs the oracle: the reference's test_const_decoder.cpp layouts
through the per-block filter entry point (every operator, pd_filter_info windows), vector decode
through the per-block projection, and whole-table scans with CONST columns filtered and projected
next to the other codecs."""
import numpy as np
import pytest

import oracle_binding as ora
from test_gpu_scan import assert_scan_matches
from test_const_kat import build_block, seed_val, ROW_CNT, SHAPES

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


def _table_of(ob, block):
    return ob.TableImage(np.concatenate([block, np.zeros((-len(block)) % 128 + 128, dtype=np.uint8)]),
                         np.array([0], dtype=np.int64), np.array([len(block)], dtype=np.int64), 0, 0)


LAYOUTS = [
    [(None, ROW_CNT)],
    [(2, ROW_CNT)],
    [(1, ROW_CNT - 3), (2, 2), (None, 1)],
    [(0, ROW_CNT - 5), (2, 3), (4, 1), (None, 1)],
    [(0, 1), (1, 58), (2, 4), (3, 1)],
    [(0, ROW_CNT - 4), (1, 1), (2, 1), (3, 2)],
    [(None, ROW_CNT - 4), (1, 2), (2, 2)],
]


@pytest.mark.parametrize("kind", ["int", "str"])
def test_const_white_filters_all_ops(ob, ctx, kind):
    v = lambda s: seed_val(s, kind)
    plist = [(ob.WHITE_OP_EQ, (v(1),)), (ob.WHITE_OP_NE, (v(1),)), (ob.WHITE_OP_LT, (v(2),)),
             (ob.WHITE_OP_LE, (v(2),)), (ob.WHITE_OP_GT, (v(0),)), (ob.WHITE_OP_GE, (v(3),)),
             (ob.WHITE_OP_BT, (v(1), v(2))), (ob.WHITE_OP_BT, (v(2), v(1))),
             (ob.WHITE_OP_IN, (v(1), v(2), v(5))), (ob.WHITE_OP_IN, (v(5), None)),
             (ob.WHITE_OP_NU, ()), (ob.WHITE_OP_NN, ()), (ob.WHITE_OP_EQ, (None,))]
    for layout in LAYOUTS:
        block = build_block(layout, kind)
        blk = ora.Block(block)
        batch = ctx.open_batch(_table_of(ob, block))
        for op, params in plist:
            for start, count in ((0, None), (ROW_CNT - 32, 30), (ROW_CNT - 33, 30), (63, 1), (10, 0)):
                exp = blk.filter_tree(ob.White(1, op, params), start, count)
                got = batch.filter_white(0, 1, op, params, start, count)
                assert np.array_equal(got, exp), (kind, layout, op, params, start, count)
        batch.close()


@pytest.mark.parametrize("kind", ["int", "str"])
@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_const_project_block(ob, ctx, kind, shape):
    n, exc, const = SHAPES[shape]
    seeds = [const] * n
    for r, s in exc.items():
        seeds[r] = s
    nulls = np.array([s is None for s in seeds], dtype=np.uint8)
    vals = [seed_val(0 if s is None else s, kind) for s in seeds]
    if kind == "int":
        col = ob.Column(ob.OBJ_INT, ob.ENC_CONST, np.array(vals, dtype=np.int64), nulls=nulls if nulls.any() else None)
    else:
        col = ob.Column(ob.OBJ_VARCHAR, ob.ENC_CONST, vals, nulls=nulls if nulls.any() else None)
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)), col])
    blk = ora.Block(block)
    batch = ctx.open_batch(_table_of(ob, block))
    rid = np.unique(np.concatenate([np.arange(0, n, 3), np.array(sorted(exc), dtype=np.int64)])).astype(np.int32)
    if kind == "int":
        ed, en, ehn = blk.get_rows_fixed(1, rid, 8, 5)
        gd, gn, ghn = batch.project_fixed(0, 1, rid, 8, 5)
        assert np.array_equal(gd, ed) and np.array_equal(gn, en) and ghn == ehn
    else:
        eo, el, en, ehn = blk.get_rows_discrete(1, rid, 5)
        gp, gl, gn, ghn = batch.project_discrete(0, 1, rid, string_base=1 << 40, vec_offset=5)
        assert np.array_equal(gl, el) and np.array_equal(gn, en) and ghn == ehn
        go = np.where(gp != 0, gp - np.uint64(1 << 40), 0)
        assert np.array_equal(go, eo)
    batch.close()


class W:
    def __init__(self, table, flt, proj, is_str, elem):
        self.table, self.filter, self.proj, self.proj_is_string, self.proj_elem_len = table, flt, proj, is_str, elem


def _const_table(ob, n, rpb, seed):
    rng = np.random.default_rng(seed)
    row = np.arange(n)
    c_exc = np.full(n, 4242, dtype=np.int64)
    c_exc[row % 101 == 7] = 99
    c_exc[row % 211 == 3] = -17
    c_exc_nulls = (row % 307 == 11).astype(np.uint8)
    c_plain = np.full(n, -123456789012, dtype=np.int64)
    c_null_const = np.zeros(n, dtype=np.int64)
    c_null_const_nulls = np.ones(n, dtype=np.uint8)
    c_null_const_nulls[row % 157 == 5] = 0
    c_null_const[row % 157 == 5] = 31
    s_vals = [b"constant-string" if r % 89 != 1 else (b"exc-%d" % (r % 3)) for r in range(n)]
    cols = [
        ob.Column(ob.OBJ_INT, ob.ENC_RAW, rng.integers(0, 1 << 10, size=n, dtype=np.int64)),
        ob.Column(ob.OBJ_INT, ob.ENC_CONST, c_exc, nulls=c_exc_nulls),
        ob.Column(ob.OBJ_INT, ob.ENC_CONST, c_plain),
        ob.Column(ob.OBJ_INT, ob.ENC_CONST, c_null_const, nulls=c_null_const_nulls),
        ob.Column(ob.OBJ_VARCHAR, ob.ENC_CONST, s_vals),
        ob.Column(ob.OBJ_INT, ob.ENC_RLE, np.repeat(rng.integers(0, 9, size=n // 5 + 1), 5)[:n].astype(np.int64)),
        ob.Column(ob.OBJ_INT32, ob.ENC_CONST, np.where(row % 97 == 0, -(1 << 31), -5).astype(np.int64)),
    ]
    return ob.encode_table(cols, rpb)


PROJ = list(range(7))
IS_STR = [False, False, False, False, True, False, False]
ELEM = [8] * 7


@pytest.mark.parametrize("flt_id", range(8))
def test_const_columns_in_table_scan(ob, ctx, flt_id):
    table = _const_table(ob, 30_000, 600, 11)
    flts = [
        None,
        ob.White(1, ob.WHITE_OP_NE, (4242,)),
        ob.White(1, ob.WHITE_OP_EQ, (4242,)),
        ob.And([ob.White(0, ob.WHITE_OP_LT, (600,)), ob.White(1, ob.WHITE_OP_LT, (100,))]),
        ob.Or([ob.White(3, ob.WHITE_OP_NN, ()), ob.White(4, ob.WHITE_OP_GT, (b"d",))]),
        ob.White(2, ob.WHITE_OP_BT, (-200000000000, 0)),
        ob.White(4, ob.WHITE_OP_IN, (b"exc-0", b"exc-2", b"nothing")),
        ob.And([ob.White(6, ob.WHITE_OP_LT, (-5,)), ob.White(1, ob.WHITE_OP_NU, ())]),
    ]
    assert_scan_matches(ctx, W(table, flts[flt_id], PROJ, IS_STR, ELEM))
