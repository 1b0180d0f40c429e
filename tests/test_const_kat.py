"""CONST column codec: the reference's own expectations (unittest/storage/blocksstable/encoding/
test_const_decoder.cpp:111-770) re-expressed against the oracle, plus writer -> oracle cell round
trips for every header shape of ObConstMetaHeader (no exceptions: value / NULL; with exceptions:
1- and 2-byte row ids, NULL as the constant, NULL as an exception). ROW_CNT = 64 and the
[seedA x .. | seedB x .. | NULL] layouts, the pd_filter_info windows and the popcounts are the
reference's; a seed maps to an increasing integer / string (all the expectations depend on)."""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
from oceanbase_b200 import White

ROW_CNT = 64
KINDS = ["int", "str"]


def seed_val(seed, kind):
    return (seed * 1000 + 7) if kind == "int" else b"seed-%04d-%s" % (seed, b"x" * (seed % 5))


def build_block(layout, kind, enc=ob.ENC_CONST):
    vals, nulls = [], []
    for seed, cnt in layout:
        for _ in range(cnt):
            nulls.append(seed is None)
            vals.append(seed_val(0 if seed is None else seed, kind))
    nulls = np.array(nulls, dtype=np.uint8)
    if kind == "int":
        col = ob.Column(ob.OBJ_INT, enc, np.array(vals, dtype=np.int64), nulls=nulls)
    else:
        col = ob.Column(ob.OBJ_VARCHAR, enc, vals, nulls=nulls)
    pad = ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(len(vals), dtype=np.int64))
    return ob.encode_block([pad, col])


def build(layout, kind):
    block = build_block(layout, kind)
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0
    h = block[blk.b.header_size + 16: blk.b.header_size + 32]
    assert int(h[1]) == 3 and int(h[2]) == 0  # ObColumnHeader::CONST, no attr bits
    return blk


def pop(blk, op, params, start=0, count=None):
    return int(blk.filter_tree(White(1, op, params), start, count).sum())


@pytest.mark.parametrize("kind", KINDS)
def test_no_exception_nu_nn(kind):
    # test_const_decoder.cpp:111-156: every row NULL
    blk = build([(None, ROW_CNT)], kind)
    assert pop(blk, ob.WHITE_OP_NU, ()) == 64
    assert pop(blk, ob.WHITE_OP_NN, ()) == 0
    assert all(blk.cell(1, r) is None for r in range(ROW_CNT))


@pytest.mark.parametrize("kind", KINDS)
def test_no_exception_other(kind):
    # :158-242: every row seed_2
    blk = build([(2, ROW_CNT)], kind)
    v = lambda s: seed_val(s, kind)
    assert pop(blk, ob.WHITE_OP_EQ, (v(2),)) == 64
    assert pop(blk, ob.WHITE_OP_NE, (v(2),)) == 0
    assert pop(blk, ob.WHITE_OP_GT, (v(2),)) == 0
    assert pop(blk, ob.WHITE_OP_LT, (v(2),)) == 0
    assert pop(blk, ob.WHITE_OP_GE, (v(2),)) == 64
    assert pop(blk, ob.WHITE_OP_LE, (v(2),)) == 64
    assert pop(blk, ob.WHITE_OP_BT, (v(1), v(3))) == 64
    assert pop(blk, ob.WHITE_OP_IN, (v(1), v(2), v(3))) == 64
    assert pop(blk, ob.WHITE_OP_EQ, (v(1),)) == 0


@pytest.mark.parametrize("kind", KINDS)
def test_filter_push_down_nu_nn_eq_ne(kind):
    # :244-369: [seed1 x N-3 | seed2 x 2 | NULL x 1], window rows [N-32, N-2)
    blk = build([(1, ROW_CNT - 3), (2, 2), (None, 1)], kind)
    w = (ROW_CNT - 32, 30)
    r1, r2 = (seed_val(1, kind),), (seed_val(2, kind),)
    s1, s2 = ROW_CNT - 3, 2
    assert (pop(blk, ob.WHITE_OP_EQ, r1), pop(blk, ob.WHITE_OP_EQ, r1, *w)) == (s1, 29)
    assert (pop(blk, ob.WHITE_OP_NE, r1), pop(blk, ob.WHITE_OP_NE, r1, *w)) == (s2, 1)
    assert (pop(blk, ob.WHITE_OP_EQ, r2), pop(blk, ob.WHITE_OP_EQ, r2, *w)) == (s2, 1)
    assert (pop(blk, ob.WHITE_OP_NE, r2), pop(blk, ob.WHITE_OP_NE, r2, *w)) == (s1, 29)
    assert (pop(blk, ob.WHITE_OP_NU, ()), pop(blk, ob.WHITE_OP_NU, (), *w)) == (1, 0)
    assert (pop(blk, ob.WHITE_OP_NN, ()), pop(blk, ob.WHITE_OP_NN, (), *w)) == (s1 + s2, 30)


@pytest.mark.parametrize("kind", KINDS)
def test_filter_push_down_gt_lt_ge_le(kind):
    # :371-553: [seed0 x N-5 | seed2 x 3 | seed4 x 1 | NULL x 1], window rows [N-33, N-3)
    blk = build([(0, ROW_CNT - 5), (2, 3), (4, 1), (None, 1)], kind)
    w = (ROW_CNT - 33, 30)
    c0, c1, c2 = ROW_CNT - 5, 3, 1
    r = (seed_val(2, kind),)
    assert (pop(blk, ob.WHITE_OP_GT, r), pop(blk, ob.WHITE_OP_GT, r, *w)) == (c2, 0)
    assert (pop(blk, ob.WHITE_OP_LT, r), pop(blk, ob.WHITE_OP_LT, r, *w)) == (c0, 28)
    assert (pop(blk, ob.WHITE_OP_GE, r), pop(blk, ob.WHITE_OP_GE, r, *w)) == (c1 + c2, 1 + c2)
    assert (pop(blk, ob.WHITE_OP_LE, r), pop(blk, ob.WHITE_OP_LE, r, *w)) == (c0 + c1, 30)
    r = (seed_val(0, kind),)
    assert (pop(blk, ob.WHITE_OP_GT, r), pop(blk, ob.WHITE_OP_GT, r, *w)) == (c2 + c1, 2)
    assert (pop(blk, ob.WHITE_OP_LT, r), pop(blk, ob.WHITE_OP_LT, r, *w)) == (0, 0)
    assert (pop(blk, ob.WHITE_OP_GE, r), pop(blk, ob.WHITE_OP_GE, r, *w)) == (c0 + c1 + c2, 30)
    assert (pop(blk, ob.WHITE_OP_LE, r), pop(blk, ob.WHITE_OP_LE, r, *w)) == (c0, 28)


@pytest.mark.parametrize("kind", KINDS)
def test_filter_push_down_bt(kind):
    # :555-651: [seed0 x 1 | seed1 x N-6 | seed2 x 4 | seed3 x 1]
    blk = build([(0, ROW_CNT - 63), (1, 58), (2, 4), (3, 1)], kind)
    w = (ROW_CNT - 33, 30)
    v = lambda s: seed_val(s, kind)
    assert (pop(blk, ob.WHITE_OP_BT, (v(2), v(3))), pop(blk, ob.WHITE_OP_BT, (v(2), v(3)), *w)) == (5, 2)
    assert (pop(blk, ob.WHITE_OP_BT, (v(1), v(3))), pop(blk, ob.WHITE_OP_BT, (v(1), v(3)), *w)) == (63, 30)


@pytest.mark.parametrize("kind", KINDS)
def test_filter_push_down_in(kind):
    # :653-768: [seed0 x N-4 | seed1 | seed2 | seed3 | seed3]
    blk = build([(0, ROW_CNT - 4), (1, 1), (2, 1), (3, 2)][:4], kind)
    w = (ROW_CNT - 33, 30)
    v = lambda s: seed_val(s, kind)
    in1 = (v(1), v(2), v(5))
    assert (pop(blk, ob.WHITE_OP_IN, in1), pop(blk, ob.WHITE_OP_IN, in1, *w)) == (2, 1)
    in2 = (v(5),) * 3
    assert (pop(blk, ob.WHITE_OP_IN, in2), pop(blk, ob.WHITE_OP_IN, in2, *w)) == (0, 0)


# ---- writer -> oracle round trips over every header shape ---------------------------------------
def _expect(vals, nulls, kind):
    if kind == "int":
        return [None if n else int(np.uint64(np.int64(x))) for x, n in zip(vals, nulls)]
    return [None if n else x for x, n in zip(vals, nulls)]


def _meta_header(block, col=1):
    blk = ora.Block(block)
    hs = blk.b.header_size
    ncol = blk.column_count
    h = block[hs + 16 * col: hs + 16 * col + 16]
    off = int(h[8:12].view(np.uint32)[0])
    length = int(h[12:16].view(np.uint32)[0])
    m = block[hs + 16 * ncol + off:]
    return dict(version=int(m[0]), count=int(m[1]), const_ref=int(m[2]), row_id_byte=int(m[3]) & 7,
                offset=int(m[4]) | int(m[5]) << 8, length=length)


SHAPES = {
    # name: (n rows, {row: seed or None}, constant seed or None)
    "no_exc_value": (500, {}, 7),
    "no_exc_null": (500, {}, None),
    "exc_1byte_rowid": (200, {0: 3, 17: 9, 199: 1}, 7),
    "exc_2byte_rowid": (3000, {5: 3, 256: 9, 2999: None, 1500: 3}, 7),
    "null_const_value_exc": (400, {1: 3, 2: 4, 399: 5}, None),
    "max_exceptions": (1000, {r * 31: (r % 9) + 20 for r in range(32)}, 2),
}


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_const_roundtrip(kind, shape):
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
    assert blk.verify_checksums() == 0
    m = _meta_header(block)
    assert m["version"] == 0 and m["count"] == len(exc)
    if not exc:
        assert m["offset"] == 6 and m["const_ref"] == (1 if const is None else 0)
        if kind == "int" and const is not None:
            assert m["length"] == 6 + 8  # sizeof(ObConstMetaHeader) + type store size of ObIntType
    else:
        assert m["row_id_byte"] == (1 if max(exc) < 256 else 2)
        assert m["offset"] == 6 + len(exc) * (1 + m["row_id_byte"])
    assert [blk.cell(1, r) for r in range(n)] == _expect(vals, nulls, kind)
    # vector decode of a strided ascending row-id list == per-cell decode
    rid = np.arange(0, n, 3, dtype=np.int32)
    if kind == "int":
        data, nb, hn = blk.get_rows_fixed(1, rid)
        got = data.view(np.uint64)
        for i, r in enumerate(rid):
            is_null = bool((nb[i // 64] >> np.uint64(i % 64)) & np.uint64(1))
            assert is_null == bool(nulls[r])
            if not is_null:
                assert int(got[i]) == int(np.uint64(np.int64(vals[r])))
        assert hn == int(nulls[rid].any())


def test_const_not_suitable_is_rejected():
    # ob_const_encoder.cpp:110-112: more than 32 exceptions, or more than 10 % of the rows
    n = 1000
    v = np.full(n, 5, dtype=np.int64)
    v[::30] = 9  # 34 exceptions
    with pytest.raises(ob.ObGpuError):
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_CONST, v)])
    v = np.full(40, 5, dtype=np.int64)
    v[:5] = 9  # 5 > max(40 * 10 / 100, 1)
    with pytest.raises(ob.ObGpuError):
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_CONST, v)])


@pytest.mark.parametrize("obj_type,const,exc", [(ob.OBJ_INT32, -5, 2 ** 31 - 1), (ob.OBJ_TINYINT, -128, 127),
                                                (ob.OBJ_UINT32, 2 ** 32 - 1, 0), (ob.OBJ_DATE, -40000, 40000)])
def test_const_narrow_types_sign_extension(obj_type, const, exc):
    for with_exc in (False, True):
        v = np.full(100, const, dtype=np.int64)
        if with_exc:
            v[37] = exc
        blk = ora.Block(ob.encode_block([ob.Column(obj_type, ob.ENC_CONST, v)]))
        for r in (0, 37, 99):
            d = blk.cell_raw(0, r)
            if obj_type == ob.OBJ_DATE:
                assert d.len == 4 and np.int32(np.uint32(d.ival)) == v[r]
            elif obj_type == ob.OBJ_UINT32:
                assert d.len == 8 and d.ival == v[r]
            else:
                assert d.len == 8 and np.int64(np.uint64(d.ival)) == v[r]
