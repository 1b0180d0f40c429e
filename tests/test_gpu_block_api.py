"""Reference-granularity entry points (one block per call) vs the oracle: white filters with
pd_filter_info windows, filter trees, ObBitmap::get_row_ids, decode_vector into VEC_FIXED /
VEC_DISCRETE with vec_offset. Same layouts as the reference's filter tests (see
tests/test_oracle_filter_kat.py for the provenance of the expectations)."""
import numpy as np
import pytest

import oracle_binding as ora

pytestmark = pytest.mark.gpu

ROW_CNT = 64


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


def _cases(ob):
    return [("int", ob.ENC_RAW), ("int", ob.ENC_DICT), ("int", ob.ENC_RLE), ("int", ob.ENC_INTEGER_BASE_DIFF),
            ("str", ob.ENC_RAW), ("str", ob.ENC_DICT), ("str", ob.ENC_RLE)]


def seed_val(seed, kind):
    return (seed * 1000 + 7) if kind == "int" else b"seed-%04d-%s" % (seed, b"x" * (seed % 5))


def build(ob, layout, kind, enc):
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


def test_white_filters_all_ops_all_codecs(ob, ctx):
    layout = [(0, ROW_CNT - 40), (1, 10), (2, 10), (3, 10), (None, 10)]
    for kind, enc in _cases(ob):
        block = build(ob, layout, kind, enc)
        blk = ora.Block(block)
        batch = ctx.open_batch(_table_of(ob, block))
        v = lambda s: seed_val(s, kind)
        plist = [(ob.WHITE_OP_EQ, (v(1),)), (ob.WHITE_OP_NE, (v(1),)), (ob.WHITE_OP_LT, (v(2),)),
                 (ob.WHITE_OP_LE, (v(2),)), (ob.WHITE_OP_GT, (v(0),)), (ob.WHITE_OP_GE, (v(3),)),
                 (ob.WHITE_OP_BT, (v(1), v(2))), (ob.WHITE_OP_BT, (v(2), v(1))),
                 (ob.WHITE_OP_IN, (v(1), v(2), v(5))), (ob.WHITE_OP_IN, (v(5), None)),
                 (ob.WHITE_OP_NU, ()), (ob.WHITE_OP_NN, ()), (ob.WHITE_OP_EQ, (None,))]
        for op, params in plist:
            for start, count in ((0, None), (ROW_CNT - 35, 30), (ROW_CNT - 45, 30), (63, 1), (10, 0)):
                exp = blk.filter_tree(ob.White(1, op, params), start, count)
                got = batch.filter_white(0, 1, op, params, start, count)
                assert np.array_equal(got, exp), (kind, enc, op, params, start, count)
        batch.close()


def test_reference_popcounts_on_gpu(ob, ctx):
    # test_raw_decoder.cpp:774-860 through the device path
    for kind, enc in _cases(ob):
        block = build(ob, [(0xF, ROW_CNT - 20), (0x0, 10), (None, 10)], kind, enc)
        batch = ctx.open_batch(_table_of(ob, block))
        ref = (seed_val(0xF, kind),)
        assert batch.filter_white(0, 1, ob.WHITE_OP_EQ, ref).sum() == ROW_CNT - 20
        assert batch.filter_white(0, 1, ob.WHITE_OP_EQ, ref, ROW_CNT - 45, 30).sum() == 25
        assert batch.filter_white(0, 1, ob.WHITE_OP_NE, ref).sum() == 10
        assert batch.filter_white(0, 1, ob.WHITE_OP_NE, ref, ROW_CNT - 45, 30).sum() == 5
        batch.close()


def test_filter_tree_and_or(ob, ctx):
    n = 3000
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1000, size=n, dtype=np.int64)
    b = rng.integers(0, 100, size=n, dtype=np.int64)
    s = [b"k%03d" % (x % 37) for x in range(n)]
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, a), ob.Column(ob.OBJ_INT, ob.ENC_DICT, b),
                             ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, s)])
    blk = ora.Block(block)
    batch = ctx.open_batch(_table_of(ob, block))
    trees = [
        ob.And([ob.White(0, ob.WHITE_OP_GE, (100,)), ob.Or([ob.White(1, ob.WHITE_OP_LT, (10,)),
                                                             ob.White(2, ob.WHITE_OP_IN, (b"k001", b"k036"))])]),
        ob.Or([ob.And([ob.White(0, ob.WHITE_OP_LT, (500,)), ob.White(1, ob.WHITE_OP_NE, (3,))]),
               ob.And([ob.White(2, ob.WHITE_OP_GT, (b"k020",)), ob.White(0, ob.WHITE_OP_BT, (700, 800))]),
               ob.White(1, ob.WHITE_OP_EQ, (99,))]),
        ob.And([ob.White(0, ob.WHITE_OP_LT, (0,)), ob.White(1, ob.WHITE_OP_GE, (0,))]),
        ob.Or([ob.White(0, ob.WHITE_OP_GE, (0,)), ob.White(1, ob.WHITE_OP_LT, (0,))]),
    ]
    for t in trees:
        for start, count in ((0, None), (17, 2000)):
            assert np.array_equal(batch.filter_tree(0, t, start, count), blk.filter_tree(t, start, count))
    batch.close()


def test_bitmap_to_row_ids(ob, ctx):
    rng = np.random.default_rng(9)
    for n, p in ((1000, 0.3), (64, 1.0), (64, 0.0), (5000, 0.01), (257, 0.9)):
        bm = (rng.random(n) < p).astype(np.uint8)
        for start, to, limit, idoff in ((0, n, 256, 0), (n // 3, n, 7, n // 3), (0, n // 2, 100000, 0), (5, 5, 4, 0)):
            e_ids, e_from = ora.bitmap_get_row_ids(bm, start, to, limit, idoff)
            g_ids, g_from = ctx.bitmap_to_row_ids(bm, start, to, limit, idoff)
            assert np.array_equal(g_ids, e_ids) and g_from == e_from, (n, p, start, to, limit)
    # batch walk like ObBlockBatchedRowStore::get_row_ids
    bm = (rng.random(3000) < 0.4).astype(np.uint8)
    frm, got = 0, []
    while frm < 3000:
        ids, frm = ctx.bitmap_to_row_ids(bm, frm, 3000, 256)
        got.extend(ids.tolist())
    assert got == np.flatnonzero(bm).tolist()


def test_project_fixed_and_discrete_with_vec_offset(ob, ctx):
    n = 900
    rng = np.random.default_rng(4)
    v = rng.integers(-10 ** 12, 10 ** 12, size=n, dtype=np.int64)
    nulls = (rng.random(n) < 0.1).astype(np.uint8)
    s = [bytes(rng.integers(97, 123, size=rng.integers(0, 25), dtype=np.uint8)) for _ in range(n)]
    sn = (rng.random(n) < 0.1).astype(np.uint8)
    d32 = rng.integers(-40000, 40000, size=n, dtype=np.int64)
    for enc_i, enc_s in ((ob.ENC_RAW, ob.ENC_RAW), (ob.ENC_DICT, ob.ENC_DICT), (ob.ENC_INTEGER_BASE_DIFF, ob.ENC_RLE)):
        ss = s if enc_s != ob.ENC_RLE else [s[i // 4] for i in range(n)]
        block = ob.encode_block([ob.Column(ob.OBJ_INT, enc_i, v, nulls=nulls),
                                 ob.Column(ob.OBJ_VARCHAR, enc_s, ss, nulls=sn),
                                 ob.Column(ob.OBJ_DATE, enc_i, d32)])
        blk = ora.Block(block)
        batch = ctx.open_batch(_table_of(ob, block))
        rid = np.sort(rng.choice(n, size=256, replace=False)).astype(np.int32)
        for vec_offset in (0, 37):
            ed, en, ehn = blk.get_rows_fixed(0, rid, 8, vec_offset)
            gd, gn, ghn = batch.project_fixed(0, 0, rid, 8, vec_offset)
            assert np.array_equal(gd, ed) and np.array_equal(gn, en) and ghn == ehn
            ed, en, ehn = blk.get_rows_fixed(2, rid, 4, vec_offset)
            gd, gn, ghn = batch.project_fixed(0, 2, rid, 4, vec_offset)
            assert np.array_equal(gd, ed) and np.array_equal(gn, en) and ghn == ehn
            eo, el, en, ehn = blk.get_rows_discrete(1, rid, vec_offset)
            gp, gl, gn, ghn = batch.project_discrete(0, 1, rid, string_base=1 << 40, vec_offset=vec_offset)
            assert np.array_equal(gl, el) and np.array_equal(gn, en) and ghn == ehn
            go = np.where(gp != 0, gp - np.uint64(1 << 40), 0)
            assert np.array_equal(go, eo)
        batch.close()


def test_datum_format_get_rows(ob, ctx):
    """ObMicroBlockDecoder::get_rows, datum format (ob_micro_block_decoder.cpp:2100-2140): 12-byte ObDatum per row; integers
    are written through the datum's own pointer with the type's datum length, strings point into the caller's block,
    NULL is set_null(). Checked against the oracle's vector output, per block and for a whole batch result."""
    import ctypes as C
    rng = np.random.default_rng(91)
    n = 3000
    strs = [bytes(rng.integers(97, 123, size=rng.integers(1, 12), dtype=np.uint8)) for _ in range(40)]
    nl = (rng.random(n) < 0.2).astype(np.uint8)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_DICT, rng.integers(-9, 9, size=n) * 10 ** 10, nulls=nl),
            ob.Column(ob.OBJ_DATE, ob.ENC_RAW, rng.integers(-(1 << 31), 1 << 31, size=n), nulls=nl),   # the 4-byte datum class
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, [strs[i] for i in rng.integers(0, 40, size=n)], nulls=nl)]
    table = ob.encode_table(cols, 800)
    batch = ctx.open_batch(table)
    base = table.image.ctypes.data           # string datums address the caller's own buffer
    blk = ora.Block(table.block(1))
    rid = np.arange(1, blk.row_count, 2, dtype=np.int32)
    for col, el in ((0, 8), (1, 4)):
        slots = np.full(len(rid) + 3, 0x5a5a5a5a5a5a5a5a, dtype=np.uint64)
        datums = np.zeros(len(rid) + 3, dtype=ob.DATUM_DTYPE)
        datums["ptr"] = slots.ctypes.data + 8 * np.arange(len(rid) + 3, dtype=np.uint64)
        batch.project_datums(1, col, rid, datums, datum_offset=3)
        wd, wn, _ = blk.get_rows_fixed(col, rid, elem_len=el)
        want = wd.view({8: np.uint64, 4: np.uint32}[el])
        for i in range(len(rid)):
            d = datums[3 + i]
            isnull = bool((wn[i // 64] >> np.uint64(i % 64)) & np.uint64(1))
            assert bool(d["pack"] & ob.DATUM_NULL_BIT) == isnull
            if isnull:
                assert d["pack"] == ob.DATUM_NULL_BIT and slots[3 + i] == 0x5a5a5a5a5a5a5a5a     # slot untouched
            else:
                assert d["pack"] == el and d["ptr"] == slots.ctypes.data + 8 * (3 + i)
                assert (int(slots[3 + i]) & ((1 << (8 * el)) - 1)) == int(want[i])
        assert np.all(datums["pack"][:3] == 0)
    datums = np.zeros(len(rid), dtype=ob.DATUM_DTYPE)
    batch.project_datums(1, 2, rid, datums, string_base=base + int(table.offsets[1]) - int(table.offsets[1]))
    offs, lens, wn, _ = blk.get_rows_discrete(2, rid)
    for i in range(len(rid)):
        isnull = bool((wn[i // 64] >> np.uint64(i % 64)) & np.uint64(1))
        assert bool(datums[i]["pack"] & ob.DATUM_NULL_BIT) == isnull
        if not isnull:
            assert datums[i]["pack"] == lens[i]
            got = C.string_at(int(datums[i]["ptr"]), int(lens[i]))
            assert got == bytes(blk.buf[int(offs[i]):int(offs[i]) + int(lens[i])])
    # whole-batch result as datums
    res = batch.scan(ob.White(1, ob.WHITE_OP_GT, (0,)), [0, 1, 2], string_base=base)
    k = res.selected_rows
    for c, el in ((0, 8), (1, 4)):
        data, _, nulls = res.fetch_col(c)
        datums, slots = res.fetch_datums(c)
        isnull = ((nulls[np.arange(k) // 64] >> (np.arange(k) % 64).astype(np.uint64)) & np.uint64(1)).astype(bool)
        assert np.array_equal((datums["pack"] & ob.DATUM_NULL_BIT) != 0, isnull)
        assert np.all(datums["pack"][~isnull] == el)
        assert np.array_equal(datums["ptr"], slots.ctypes.data + 8 * np.arange(k, dtype=np.uint64))
        assert np.array_equal(slots[~isnull], data.astype(np.uint64)[~isnull])
    data, lens, nulls = res.fetch_col(2)
    datums, slots = res.fetch_datums(2)
    isnull = ((nulls[np.arange(k) // 64] >> (np.arange(k) % 64).astype(np.uint64)) & np.uint64(1)).astype(bool)
    assert slots is None and np.array_equal(datums["ptr"][~isnull], data[~isnull])
    assert np.array_equal(datums["pack"][~isnull], lens[~isnull].astype(np.uint32))
    assert np.all(datums["pack"][isnull] == ob.DATUM_NULL_BIT)
    res.free()
    batch.close()
