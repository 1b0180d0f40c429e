"""Writer -> oracle round trip (the reference's own test strategy, SURVEY.md 4: encode with the
encoder, decode with the decoder, compare every cell -- test_micro_block_decoder.cpp:155-189), for
every encoding / width / NULL shape the writer emits, plus batch-vs-cell decode equivalence
(test_general_column_decoder.cpp:97-381) and header/checksum validity."""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora

RNG = np.random.default_rng(42)


def _decode_all(blk, col):
    return [blk.cell(col, r) for r in range(blk.row_count)]


def _expect_int(v, nulls, obj_type=ob.OBJ_INT):
    out = []
    for x, n in zip(v, nulls):
        out.append(None if n else int(np.uint64(np.int64(x))))
    return out


def _col_header(block, blk, col):
    h = block[blk.b.header_size + 16 * col: blk.b.header_size + 16 * (col + 1)]
    return dict(type=int(h[1]), attr=int(h[2]), obj_type=int(h[3]), offset=int(h[8:12].view(np.uint32)[0]),
                length=int(h[12:16].view(np.uint32)[0]))


INT_SHAPES = {
    "full64": lambda n: RNG.integers(-(1 << 63), (1 << 63) - 1, size=n, dtype=np.int64),
    "neg_small": lambda n: RNG.integers(-1000, 1000, size=n, dtype=np.int64),
    "w7": lambda n: RNG.integers(0, 1 << 7, size=n, dtype=np.int64),
    "w13": lambda n: RNG.integers(0, 1 << 13, size=n, dtype=np.int64),
    "w16": lambda n: RNG.integers(0, 1 << 16, size=n, dtype=np.int64),
    "w21": lambda n: RNG.integers(0, 1 << 21, size=n, dtype=np.int64),
    "w23": lambda n: RNG.integers(0, 1 << 23, size=n, dtype=np.int64),   # -> 3-byte packing
    "w33": lambda n: RNG.integers(0, 1 << 33, size=n, dtype=np.int64),
    "w47": lambda n: RNG.integers(0, 1 << 47, size=n, dtype=np.int64),
    "zeros": lambda n: np.zeros(n, dtype=np.int64),
    "const": lambda n: np.full(n, 123456789, dtype=np.int64),
    "lowcard": lambda n: RNG.integers(0, 9, size=n, dtype=np.int64) * 1_000_003,
    "sorted": lambda n: np.cumsum(RNG.integers(1, 700, size=n, dtype=np.int64)) + 10 ** 12,
    "runs": lambda n: np.repeat(RNG.integers(0, 50, size=n // 16 + 1, dtype=np.int64) * 77, 16)[:n],
}


@pytest.mark.parametrize("enc", [ob.ENC_RAW, ob.ENC_DICT, ob.ENC_RLE, ob.ENC_INTEGER_BASE_DIFF])
@pytest.mark.parametrize("shape", sorted(INT_SHAPES))
@pytest.mark.parametrize("null_frac", [0.0, 0.1])
def test_int_roundtrip(enc, shape, null_frac):
    n = 777
    v = INT_SHAPES[shape](n)
    if enc == ob.ENC_RLE and shape in ("full64", "w33", "w47", "w21", "w23", "w13", "w16", "neg_small", "sorted"):
        v = np.repeat(v[: n // 8 + 1], 8)[:n]  # keep the run table inside the reference's int16 ref_offset_
    nulls = (RNG.random(n) < null_frac).astype(np.uint8)
    if null_frac == 0:
        nulls[:] = 0
    block = ob.encode_block([ob.Column(ob.OBJ_INT, enc, v, nulls=nulls if nulls.any() else None)])
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0
    assert blk.row_count == n and blk.column_count == 1
    assert _decode_all(blk, 0) == _expect_int(v, nulls)
    # batch (vector) decode == per-cell decode, for a strided ascending row-id list
    rid = np.arange(0, n, 3, dtype=np.int32)
    data, nb, hn = blk.get_rows_fixed(0, rid)
    vals = data.view(np.uint64)
    for i, r in enumerate(rid):
        is_null = bool((nb[i // 64] >> np.uint64(i % 64)) & np.uint64(1))
        assert is_null == bool(nulls[r])
        if not is_null:
            assert int(vals[i]) == int(np.uint64(np.int64(v[r])))
    assert hn == int(nulls[rid].any())


def test_width_rules_match_get_packing_size():
    # encoding/ob_encoding_util.cpp:37-73: bits unless it saves too little vs whole bytes
    cases = {  # max value -> (bit_packing, length)
        0: (True, 1), 1: (True, 1), 127: (True, 7), 255: (False, 1), 256: (True, 9), (1 << 13) - 1: (True, 13),
        (1 << 16) - 1: (False, 2), (1 << 21) - 1: (True, 21), (1 << 23) - 1: (False, 3), (1 << 31) - 1: (False, 4),
        (1 << 33) - 1: (True, 33), (1 << 47) - 1: (False, 6), (1 << 57) - 1: (True, 57), (1 << 62) - 1: (False, 8),
    }
    for mx, (bp, length) in cases.items():
        v = np.zeros(50, dtype=np.int64)
        v[0] = mx
        block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, v)])
        h = _col_header(block, ora.Block(block), 0)
        assert bool(h["attr"] & 0x4) == bp, mx
        assert h["length"] == length, mx
        assert h["attr"] & 0x1  # FIX_LENGTH


def test_negative_value_forces_8_bytes():
    v = np.array([1, 2, -1, 3], dtype=np.int64)
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, v)])
    h = _col_header(block, ora.Block(block), 0)
    assert h["length"] == 8 and not (h["attr"] & 0x4)


@pytest.mark.parametrize("obj_type,lo,hi", [(ob.OBJ_INT32, -(1 << 31), (1 << 31) - 1), (ob.OBJ_TINYINT, -128, 127),
                                            (ob.OBJ_SMALLINT, -(1 << 15), (1 << 15) - 1),
                                            (ob.OBJ_UINT32, 0, (1 << 32) - 1), (ob.OBJ_DATE, -50000, 50000)])
@pytest.mark.parametrize("enc", [ob.ENC_RAW, ob.ENC_DICT, ob.ENC_INTEGER_BASE_DIFF])
def test_narrow_integer_types_sign_extension(obj_type, lo, hi, enc):
    # rule 8c.5: only ObIntTC is sign-extended (mask ~INTEGER_MASK_TABLE[type_size]); datum len 8
    # except the 4-byte map types
    n = 300
    v = RNG.integers(lo, hi, size=n, dtype=np.int64, endpoint=True)
    v[:2] = [lo, hi]
    block = ob.encode_block([ob.Column(obj_type, enc, v)])
    blk = ora.Block(block)
    for r in range(n):
        d = blk.cell_raw(0, r)
        if obj_type == ob.OBJ_DATE:
            assert d.len == 4 and np.int32(np.uint32(d.ival)) == v[r]
        elif obj_type == ob.OBJ_UINT32:
            assert d.len == 8 and d.ival == v[r]
        else:
            assert d.len == 8 and np.int64(np.uint64(d.ival)) == v[r]


def _strings(n, card, min_len, max_len):
    d = [bytes(RNG.integers(97, 123, size=RNG.integers(min_len, max_len + 1), dtype=np.uint8)) for _ in range(card)]
    idx = RNG.integers(0, card, size=n)
    return [d[i] for i in idx]


@pytest.mark.parametrize("enc", [ob.ENC_RAW, ob.ENC_DICT, ob.ENC_RLE])
@pytest.mark.parametrize("fixed_len", [False, True])
@pytest.mark.parametrize("with_nulls", [False, True])
def test_string_roundtrip(enc, fixed_len, with_nulls):
    n = 400
    vals = _strings(n, 37, 12 if fixed_len else 0, 12 if fixed_len else 40)
    if enc == ob.ENC_RLE:
        vals = [vals[i // 5] for i in range(n)]
    nulls = (RNG.random(n) < 0.15).astype(np.uint8) if with_nulls else np.zeros(n, dtype=np.uint8)
    other = RNG.integers(0, 1000, size=n, dtype=np.int64)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, other),
            ob.Column(ob.OBJ_VARCHAR, enc, vals, nulls=nulls if with_nulls else None)]
    block = ob.encode_block(cols)
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0
    got = _decode_all(blk, 1)
    for r in range(n):
        assert got[r] == (None if nulls[r] else vals[r])
    rid = np.arange(1, n, 2, dtype=np.int32)
    offs, lens, nb, hn = blk.get_rows_discrete(1, rid)
    for i, r in enumerate(rid):
        is_null = bool((nb[i // 64] >> np.uint64(i % 64)) & np.uint64(1))
        assert is_null == bool(nulls[r])
        if not is_null:
            assert bytes(block[int(offs[i]):int(offs[i]) + int(lens[i])]) == vals[r]


def test_two_var_columns_in_row_data():
    # fill_row_data: [ext bits][col_idx_byte][idx x (nvar-1)][cells]  (ob_micro_block_encoder.cpp:809-931)
    n = 200
    a = _strings(n, 50, 0, 30)
    b = _strings(n, 50, 1, 300)   # forces 2-byte column index for some rows
    c = _strings(n, 50, 3, 9)
    na = (RNG.random(n) < 0.2).astype(np.uint8)
    nc = (RNG.random(n) < 0.2).astype(np.uint8)
    block = ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, a, nulls=na),
                             ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)),
                             ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, b),
                             ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, c, nulls=nc)])
    blk = ora.Block(block)
    assert blk.b.var_column_count == 3 and blk.b.row_index_byte in (2, 4)
    for r in range(n):
        assert blk.cell(0, r) == (None if na[r] else a[r])
        assert blk.cell(1, r) == r
        assert blk.cell(2, r) == b[r]
        assert blk.cell(3, r) == (None if nc[r] else c[r])


def test_dict_is_sorted_and_flagged():
    # DICT sorts its dictionary (try_set_need_sort, ob_block_sstable_struct.h:518-524); the fixed
    # dict carries IS_SORTED, RLE's dictionary keeps first-seen order
    v = np.array([50, 10, 40, 10, 30, 50, 20], dtype=np.int64)
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_DICT, v), ob.Column(ob.OBJ_INT, ob.ENC_RLE, v)])
    blk = ora.Block(block)
    meta = blk.b.header_size + 32
    h0, h1 = _col_header(block, blk, 0), _col_header(block, blk, 1)
    d0 = block[meta + h0["offset"]:]
    assert d0[8] & 0x2 and d0[8] & 0x1
    cnt = int(d0[2:6].view(np.uint32)[0])
    assert list(d0[9:9 + cnt]) == [10, 20, 30, 40, 50]
    r = block[meta + h1["offset"]:]
    dict_off = int(r[6:10].view(np.uint32)[0])
    d1 = r[dict_off:]
    assert not (d1[8] & 0x2)
    assert list(d1[9:9 + 5]) == [50, 10, 40, 30, 20]
    assert _decode_all(blk, 0) == list(v) and _decode_all(blk, 1) == list(v)


def test_table_image_alignment_and_headers():
    n = 10_000
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, RNG.integers(0, 1 << 40, size=n, dtype=np.int64)),
            ob.Column(ob.OBJ_INT, ob.ENC_DICT, RNG.integers(0, 30, size=n, dtype=np.int64))]
    t = ob.encode_table(cols, rows_per_block=777, rowkey_cnt=1)
    assert t.n_blocks == (n + 776) // 777 and t.total_rows == n
    assert (t.offsets % 128 == 0).all()
    rows = 0
    for i in range(t.n_blocks):
        blk = ora.Block(t.block(i))
        assert blk.verify_checksums() == 0
        assert blk.b.rowkey_column_count == 1
        rows += blk.row_count
        end = t.offsets[i] + t.sizes[i]
        nxt = t.offsets[i + 1] if i + 1 < t.n_blocks else t.image.size
        assert not t.image[end:nxt].any()  # padding is zero
    assert rows == n


def test_page_batch_bounds_with_ramp():
    from oceanbase_b200.pipeline import batch_bounds
    assert batch_bounds(100, 16) == [0, 16, 32, 48, 64, 80, 96, 100]
    assert batch_bounds(100, 16, 2) == [0, 4, 12, 28, 44, 60, 76, 92, 100]
    assert batch_bounds(5, 16, 2) == [0, 4, 5]
    assert batch_bounds(3, 16, 3) == [0, 2, 3]
    assert batch_bounds(1, 1, 2) == [0, 1]


@pytest.mark.parametrize("obj_type,lo,hi,null_frac,expect_var", [
    (ob.OBJ_DATE, -30000, 30000, 0.7, True), (ob.OBJ_UINT32, 1065958239, 1065960342, 0.6, True),
    (ob.OBJ_INT, -5, 1 << 40, 0.5, True), (ob.OBJ_INT32, -100, 100, 0.7, True),
    (ob.OBJ_INT, 0, 1000, 0.1, False), (ob.OBJ_TINYINT, -128, 127, 0.7, False)])
def test_integer_column_turned_into_var_store(obj_type, lo, hi, null_frac, expect_var):
    # ObRawEncoder::traverse (ob_raw_encoder.cpp:106-110,150-155): when NULLs dominate, a RAW integer column is stored in
    # the row data, fix_data_size_ bytes per non-NULL cell and nothing per NULL cell (get_var_length :194-234)
    rng = np.random.default_rng(3)
    n = 3000
    v = rng.integers(lo, hi, size=n, dtype=np.int64)
    nulls = (rng.random(n) < null_frac).astype(np.uint8)
    s = [b"x%d" % i + b"y" * (i % 3) for i in range(n)]
    table = ob.encode_table([ob.Column(obj_type, ob.ENC_RAW, v, nulls=nulls), ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, s),
                             ob.Column(obj_type, ob.ENC_RAW, v[::-1].copy(), nulls=nulls)], 700)
    for b in range(table.n_blocks):
        raw = table.block(b)
        blk = ora.Block(raw)
        assert blk.verify_checksums() == 0
        assert bool(raw[64 + 2] & 1) != expect_var               # ObColumnHeader FIX_LENGTH attribute of column 0
        assert int(raw[22:24].view(np.uint16)[0]) == (3 if expect_var else 1)    # var_column_count_
        for c, vals in ((0, v), (2, v[::-1])):
            for r in range(0, blk.row_count, 7):
                g = b * 700 + r
                d = blk.cell_raw(c, r)
                if nulls[g]:
                    assert d.is_null == 1
                elif obj_type in (ob.OBJ_DATE,):
                    assert d.len == 4 and np.int32(np.uint32(d.ival)) == vals[g]
                elif obj_type == ob.OBJ_UINT32:
                    assert d.ival == vals[g]
                else:
                    assert d.is_null == 0 and np.int64(np.uint64(d.ival)) == vals[g]
        assert [blk.cell(1, r) for r in range(0, blk.row_count, 50)] == [s[b * 700 + r] for r in range(0, blk.row_count, 50)]
