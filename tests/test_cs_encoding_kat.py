"""CS_ENCODING_ROW_STORE blocks (column-store encoding), INTEGER columns with RAW integer streams:
writer -> oracle round trips over every ObIntegerStreamMeta shape the encoder rules produce
(cs_encoding/ob_integer_column_encoder.cpp:177-287: no base / base for negative minima, NULL as a
replaced value below the minimum or above the maximum, NULL bitmap when the type's value range is
exhausted; widths 1/2/4/8), the block framing (ObAllColumnHeader, ObCSColumnHeader, stream end
offsets relative to the block start) and the white-filter semantics on top of it."""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
from oceanbase_b200 import White

RNG = np.random.default_rng(77)
I64_MIN, I64_MAX = -(1 << 63), (1 << 63) - 1

SHAPES = {
    # name: (obj_type, values, expect meta: (width bytes, use_base, null mode with NULLs present))
    "u8_from_zero": (ob.OBJ_INT, lambda n: RNG.integers(0, 200, size=n), (1, False, "replace_max_plus_1")),
    "u16_positive_min": (ob.OBJ_INT, lambda n: RNG.integers(300, 60000, size=n), (2, False, "replace_min_minus_1")),
    "negative_small": (ob.OBJ_INT, lambda n: RNG.integers(-100, 100, size=n), (1, True, "replace_min_minus_1")),
    "negative_wide": (ob.OBJ_INT, lambda n: RNG.integers(-(1 << 40), 1 << 40, size=n), (8, True, "replace_min_minus_1")),
    "full_range": (ob.OBJ_INT, lambda n: np.concatenate([[I64_MIN, I64_MAX], RNG.integers(-5, 5, size=n - 2)]), (8, True, "bitmap")),
    "int32_full": (ob.OBJ_INT32, lambda n: np.concatenate([[-(1 << 31), (1 << 31) - 1], RNG.integers(-9, 9, size=n - 2)]), (4, True, "bitmap")),
    "uint32_from_zero_to_max": (ob.OBJ_UINT32, lambda n: np.concatenate([[0, (1 << 32) - 1], RNG.integers(0, 9, size=n - 2)]), (4, False, "bitmap")),
    "uint64_big": (ob.OBJ_UINT64, lambda n: RNG.integers(1 << 40, 1 << 62, size=n), (8, False, "replace_min_minus_1")),
    "date": (ob.OBJ_DATE, lambda n: RNG.integers(8036, 10562, size=n), (2, False, "replace_min_minus_1")),
    "tinyint_neg": (ob.OBJ_TINYINT, lambda n: RNG.integers(-128, 127, size=n), (1, True, "replace_max_plus_1")),
}


def stream_meta_of(block, blk, col):
    """Parses the column's serialized ObIntegerStreamMeta straight from the bytes (single-stream INTEGER columns)."""
    hs, ncol = blk.b.header_size, blk.column_count
    first = hs + 12 + 4 * ncol
    ah = block[hs:hs + 12]
    offsets_len = int(ah[6:10].view(np.uint32)[0])
    so = block[len(block) - offsets_len:]
    ow = 1 << int(so[3])
    ends = so[5:].view({1: np.uint8, 2: np.uint16, 4: np.uint32}[ow])
    pos = first if col == 0 else int(ends[col - 1])
    attrs = int(block[hs + 12 + 4 * col + 2])
    if attrs & 0x02:
        pos += (blk.row_count + 7) // 8
    m = block[pos:]
    return dict(version=int(m[0]), attr=int(m[1]), type=int(m[2]), width=1 << int(m[3]), has_bitmap=bool(attrs & 0x02),
                end=int(ends[col]))


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("with_nulls", [False, True])
def test_cs_integer_roundtrip(shape, with_nulls):
    obj_type, gen, (width, use_base, null_mode) = SHAPES[shape]
    n = 500
    v = gen(n).astype(np.int64) if obj_type != ob.OBJ_UINT64 else gen(n).astype(np.uint64).view(np.int64)
    nulls = (RNG.random(n) < 0.15).astype(np.uint8) if with_nulls else None
    if nulls is not None:
        nulls[:2] = 0  # keep the extreme values of the "full range" shapes
    pad = np.arange(n, dtype=np.int64)
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, pad), ob.Column(obj_type, ob.ENC_CS_INTEGER, v, nulls=nulls)])
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0
    assert blk.b.row_store_type == 3 and blk.b.cs_stream_count == 2
    m = stream_meta_of(block, blk, 1)
    assert m["version"] == 1 and m["type"] == 1            # V2 meta, RAW stream
    assert bool(m["attr"] & 1) == use_base, m
    if with_nulls:
        assert m["has_bitmap"] == (null_mode == "bitmap")
        assert bool(m["attr"] & 2) == (null_mode != "bitmap")
    else:
        assert not m["has_bitmap"] and not (m["attr"] & 2)
        assert m["width"] == width
    for r in range(n):
        d = blk.cell_raw(1, r)
        if nulls is not None and nulls[r]:
            assert d.is_null == 1
            continue
        assert d.is_null == 0
        if obj_type == ob.OBJ_DATE:
            assert d.len == 4 and np.int32(np.uint32(d.ival)) == v[r]
        elif obj_type in (ob.OBJ_UINT32, ob.OBJ_UINT64):
            assert d.len == 8 and d.ival == int(np.uint64(v[r]))
        else:
            assert d.len == 8 and np.int64(np.uint64(d.ival)) == v[r]
    assert [blk.cell(0, r) for r in range(0, n, 37)] == list(range(0, n, 37))


def test_cs_white_filters_match_pax():
    # the same values stored as PAX RAW and as CS INTEGER give the same bitmaps for every operator
    n = 64
    v = np.array([7] * 34 + [1007] * 10 + [2007] * 10 + [0] * 10, dtype=np.int64)
    nulls = np.array([0] * 54 + [1] * 10, dtype=np.uint8)
    pad = np.arange(n, dtype=np.int64)
    pax = ora.Block(ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, pad), ob.Column(ob.OBJ_INT, ob.ENC_RAW, v, nulls=nulls)]))
    cs = ora.Block(ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, pad), ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, v, nulls=nulls)]))
    plist = [(ob.WHITE_OP_EQ, (1007,)), (ob.WHITE_OP_NE, (1007,)), (ob.WHITE_OP_LT, (2007,)), (ob.WHITE_OP_LE, (7,)),
             (ob.WHITE_OP_GT, (7,)), (ob.WHITE_OP_GE, (2007,)), (ob.WHITE_OP_BT, (8, 2007)), (ob.WHITE_OP_IN, (7, 2007, 5)),
             (ob.WHITE_OP_NU, ()), (ob.WHITE_OP_NN, ()), (ob.WHITE_OP_EQ, (None,))]
    for op, params in plist:
        for start, count in ((0, None), (19, 30)):
            a = pax.filter_tree(White(1, op, params), start, count)
            b = cs.filter_tree(White(1, op, params), start, count)
            assert np.array_equal(a, b), (op, params)
    # reference popcounts (test_raw_decoder.cpp:862-980 layout re-used by the cs decoder tests): GT seed1 -> 10
    assert int(cs.filter_tree(White(1, ob.WHITE_OP_GT, (1007,))).sum()) == 10
    assert int(cs.filter_tree(White(1, ob.WHITE_OP_NU, ())).sum()) == 10


def test_mixed_row_store_types_in_one_block_are_rejected():
    v = np.arange(10, dtype=np.int64)
    with pytest.raises(ob.ObGpuError):
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, v), ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, v)])


@pytest.mark.parametrize("obj_type,lo,hi", [(ob.OBJ_INT, -50, 50), (ob.OBJ_INT, 0, 70000), (ob.OBJ_UINT64, 1 << 40, (1 << 40) + 300),
                                            (ob.OBJ_DATE, 8000, 8100), (ob.OBJ_INT32, -(1 << 31), -(1 << 31) + 200)])
@pytest.mark.parametrize("with_nulls", [False, True])
def test_cs_int_dict_roundtrip(obj_type, lo, hi, with_nulls):
    # ObIntDictColumnDecoder::decode (cs_encoding/ob_int_dict_column_decoder.cpp:25-60): ref == distinct count is NULL
    n = 600
    v = RNG.integers(lo, hi, size=n).astype(np.int64)
    nulls = (RNG.random(n) < 0.2).astype(np.uint8) if with_nulls else None
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.arange(n, dtype=np.int64)),
                             ob.Column(obj_type, ob.ENC_CS_INT_DICT, v, nulls=nulls),
                             ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, v)])
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0 and blk.b.cs_stream_count == 4
    hs = blk.b.header_size
    assert int(block[hs + 12 + 4 + 1]) == 2                      # ObCSColumnHeader::INT_DICT
    for r in range(n):
        d = blk.cell_raw(1, r)
        if nulls is not None and nulls[r]:
            assert d.is_null == 1
        elif obj_type == ob.OBJ_DATE:
            assert d.is_null == 0 and d.len == 4 and np.int32(np.uint32(d.ival)) == v[r]
        else:
            assert d.is_null == 0 and np.int64(np.uint64(d.ival)) == v[r]
    assert [blk.cell(2, r) for r in range(0, n, 41)] == [int(np.uint64(x)) for x in v[::41]]   # column after the dict column
    mid = int(np.median(v))
    pax = ora.Block(ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)),
                                     ob.Column(obj_type, ob.ENC_DICT, v, nulls=nulls)]))
    for op, params in ((ob.WHITE_OP_LT, (mid,)), (ob.WHITE_OP_IN, (int(v[0]), int(v[1]), mid)), (ob.WHITE_OP_NU, ()), (ob.WHITE_OP_NE, (mid,))):
        assert np.array_equal(blk.filter_tree(White(1, op, params)), pax.filter_tree(White(1, op, params)))


def test_cs_int_dict_all_null_column_has_no_streams():
    n = 100
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_CS_INT_DICT, np.zeros(n, dtype=np.int64), nulls=np.ones(n, dtype=np.uint8)),
                             ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.arange(n, dtype=np.int64))])
    blk = ora.Block(block)
    assert blk.b.cs_stream_count == 1
    assert all(blk.cell(0, r) is None for r in range(n))
    assert [blk.cell(1, r) for r in range(n)] == list(range(n))


def _strings(rng, n, minl, maxl, card):
    d = [bytes(rng.integers(97, 123, size=int(rng.integers(minl, maxl + 1)), dtype=np.uint8)) for _ in range(card)]
    return [d[i] for i in rng.integers(0, card, size=n)]


STR_SHAPES = {
    # name: (generator, ObCSColumnHeader attrs expected for (no NULLs, some NULLs) with the STRING encoding)
    "var": (lambda rng, n: _strings(rng, n, 1, 12, 50), (0x00, 0x00)),                  # NULL as zero length
    "var_with_empty": (lambda rng, n: [b"" if i % 7 == 0 else x for i, x in enumerate(_strings(rng, n, 1, 9, 40))], (0x00, 0x02)),
    "fixed": (lambda rng, n: _strings(rng, n, 5, 5, 30), (0x01, None)),                # bitmap vs offsets by size
    "all_empty": (lambda rng, n: [b""] * n, (0x01, 0x03)),
}


@pytest.mark.parametrize("shape", sorted(STR_SHAPES))
@pytest.mark.parametrize("null_frac", [0.0, 0.2, 1.0])
@pytest.mark.parametrize("enc", ["string", "str_dict"])
def test_cs_string_roundtrip(shape, null_frac, enc):
    # ObStringColumnEncoder::do_init_ (cs_encoding/ob_string_column_encoder.cpp:44-99): fixed length, NULL as a zero
    # length value or a NULL bitmap; ObStrDictColumnEncoder: [dict meta][string stream (+offsets)][ref stream];
    # every string stream's bytes live in the block's all-string-data area (ob_micro_block_cs_encoder.cpp store_all_string_data_)
    rng = np.random.default_rng(5)
    n = 400
    gen, (attrs_plain, attrs_nulls) = STR_SHAPES[shape]
    v = gen(rng, n)
    nulls = (rng.random(n) < null_frac).astype(np.uint8) if null_frac else None
    e = ob.ENC_CS_STRING if enc == "string" else ob.ENC_CS_STR_DICT
    block = ob.encode_block([ob.Column(ob.OBJ_VARCHAR, e, v, nulls=nulls),
                             ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.arange(n, dtype=np.int64)),
                             ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STRING, v[::-1])])
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0
    hs = blk.b.header_size
    assert int(block[hs + 12 + 1]) == (1 if enc == "string" else 3)       # ObCSColumnHeader::STRING / STR_DICT
    if enc == "string" and 0.0 < null_frac < 1.0:
        if attrs_nulls is not None:
            assert int(block[hs + 12 + 2]) == attrs_nulls
    elif enc == "string" and null_frac == 0.0:
        assert int(block[hs + 12 + 2]) == attrs_plain
    exp = [None if (nulls is not None and nulls[r]) else v[r] for r in range(n)]
    assert [blk.cell(0, r) for r in range(n)] == exp
    assert [blk.cell(2, r) for r in range(n)] == v[::-1]
    assert [blk.cell(1, r) for r in range(0, n, 50)] == list(range(0, n, 50))
    # white filters agree with the PAX RAW encoding of the same cells
    pax = ora.Block(ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, v, nulls=nulls)]))
    probe = v[3]
    for op, params in ((ob.WHITE_OP_EQ, (probe,)), (ob.WHITE_OP_LT, (probe,)), (ob.WHITE_OP_GE, (b"m",)), (ob.WHITE_OP_NU, ()),
                       (ob.WHITE_OP_IN, (probe, v[9], b"zzzz")), (ob.WHITE_OP_NE, (probe,)), (ob.WHITE_OP_BT, (b"c", b"t"))):
        assert np.array_equal(blk.filter_tree(White(0, op, params)), pax.filter_tree(White(0, op, params))), (op, params)


@pytest.mark.parametrize("kind", ["int", "str"])
@pytest.mark.parametrize("shape", ["one_value", "dominant_value", "dominant_null", "too_many_exceptions", "ten_percent"])
def test_cs_dict_const_encoded_refs(kind, shape):
    # ObDictColumnEncoder::try_const_encoding_ref_ (cs_encoding/ob_dict_column_encoder.cpp:144-189): one ref covering
    # all rows, or all but <= 64 rows and < 10 % of them, turns the ref stream into
    # [exception count][const ref][exception row ids][exception refs] (ob_dict_column_encoder.h:65-116)
    rng = np.random.default_rng(23)
    n = 1000
    idx = np.zeros(n, dtype=np.int64)
    nulls = None
    exc = {"one_value": 0, "dominant_value": 40, "dominant_null": 30, "too_many_exceptions": 70, "ten_percent": 100}[shape]
    if shape in ("too_many_exceptions", "ten_percent"):
        n = 2000 if shape == "too_many_exceptions" else 1000    # 70 of 2000 (> 64), 100 of 1000 (not < 10 %)
        idx = np.zeros(n, dtype=np.int64)
    where = rng.choice(n, size=exc, replace=False)
    idx[where] = rng.integers(1, 9, size=exc)
    if shape == "dominant_null":
        nulls = np.ones(n, dtype=np.uint8)
        nulls[where] = 0
    if kind == "int":
        vals = (idx * 1000 - 3000).astype(np.int64)
        col = ob.Column(ob.OBJ_INT, ob.ENC_CS_INT_DICT, vals, nulls=nulls)
        expect = [None if (nulls is not None and nulls[r]) else int(np.uint64(vals[r])) for r in range(n)]
    else:
        words = [b"", b"alpha", b"be", b"gamma", b"delta!", b"e", b"zeta", b"eta", b"theta"]
        vals = [words[i] for i in idx.tolist()]
        col = ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STR_DICT, vals, nulls=nulls)
        expect = [None if (nulls is not None and nulls[r]) else vals[r] for r in range(n)]
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.arange(n, dtype=np.int64)), col,
                             ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.arange(n, dtype=np.int64) * 3)])
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0
    # ObDictEncodingMeta of column 1: right after column 0's stream
    hs, ncol = blk.b.header_size, 3
    ah = block[hs:hs + 12]
    so = block[len(block) - int(ah[6:10].view(np.uint32)[0]):]
    ends = so[5:].view({1: np.uint8, 2: np.uint16, 4: np.uint32}[1 << int(so[3])])
    dm = block[int(ends[0]):int(ends[0]) + 10]
    is_const = bool(dm[1] & 0x4)
    assert is_const == (shape in ("one_value", "dominant_value", "dominant_null"))
    ref_row_cnt = int(dm[6:10].view(np.uint32)[0])
    assert ref_row_cnt == (2 + 2 * exc if is_const else n)
    assert [blk.cell(1, r) for r in range(n)] == expect
    assert [blk.cell(2, r) for r in range(0, n, 97)] == [3 * r for r in range(0, n, 97)]
    probe = expect[int(where[0])] if exc and shape != "dominant_null" else (expect[0] if expect[0] is not None else expect[int(where[0])])
    for op, params in ((ob.WHITE_OP_EQ, (probe,)), (ob.WHITE_OP_NE, (probe,)), (ob.WHITE_OP_NU, ()), (ob.WHITE_OP_GE, (probe,))):
        if kind == "int":
            params = tuple(int(np.int64(np.uint64(x))) for x in params)
        bits = blk.filter_tree(White(1, op, params))
        want = np.array([(e is None) if op == ob.WHITE_OP_NU else (e is not None and {ob.WHITE_OP_EQ: e == probe, ob.WHITE_OP_NE: e != probe,
                         ob.WHITE_OP_GE: (np.int64(np.uint64(e)) >= np.int64(np.uint64(probe))) if kind == "int" else e >= probe}[op])
                         for e in expect], dtype=np.uint8)
        assert np.array_equal(bits, want), op
