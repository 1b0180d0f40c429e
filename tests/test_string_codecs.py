"""HEX_PACKING / STRING_DIFF / STRING_PREFIX: the PAX codecs whose values do not exist in the block and are rebuilt by
the decoder (encoding/ob_hex_string_decoder.cpp, ob_string_diff_decoder.cpp, ob_string_prefix_decoder.cpp). CPU half: the writer's
encoders against the oracle's decoders on every store shape (fixed / var store, with and without hex packing, NULL and NOP cells, one
or several var-stored columns in the row), the reference's own packing KAT (unittest/.../test_hex.cpp:24-47), the "not suitable"
conditions, and white filters / scans over such columns (the reference's retro path: decode each row, compare)."""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora


def pick(rng, alpha, k):
    return bytes(np.frombuffer(alpha, dtype=np.uint8)[rng.integers(0, len(alpha), size=k)])


def codec_cases(n=300, seed=1):
    """(name, encoding, values, nulls)"""
    rng = np.random.default_rng(seed)
    nl = (rng.random(n) < 0.2).astype(np.uint8)
    nop = np.where(rng.random(n) < 0.1, 2, nl).astype(np.uint8)   # NOP cells: 2-bit ext values
    many_null = (rng.random(n) < 0.8).astype(np.uint8)
    hexs = [pick(rng, b"0123456789abcdef", rng.integers(0, 20)) for _ in range(n)]
    hexfix = [pick(rng, b"0123456789", 11) for _ in range(n)]
    diffs = [b"ORDER-2024-" + pick(rng, b"0123456789", 5) + b"-X" for _ in range(n)]
    diffs2 = [b"id:" + bytes(rng.integers(33, 120, size=4, dtype=np.uint8)) + b":tail" for _ in range(n)]
    diff1 = [b"AB" + pick(rng, b"xyz", 1) + b"CD" for _ in range(n)]     # one differing byte: no hex packing (row_store_size 1)
    long_diff = [b"k" * 140 + pick(rng, b"01", 130) + b"t" * 20 for _ in range(n)]   # runs longer than the 127-count DiffDesc
    pre = [[b"http://www.example.com/", b"https://oceanbase.com/docs/", b"ftp://x/", b""][rng.integers(0, 4)] +
           bytes(rng.integers(97, 123, size=rng.integers(0, 12), dtype=np.uint8)) for _ in range(n)]
    prehex = [[b"AAAA-", b"BBBB-"][rng.integers(0, 2)] + pick(rng, b"0123456789", rng.integers(0, 9)) for _ in range(n)]
    many = [bytes([65 + i % 20]) * 3 + pick(rng, b"pq", rng.integers(1, 5)) for i in range(n)]    # 20 first bytes: 16 prefixes + ungrouped rows
    H, D, P = ob.ENC_HEX_PACKING, ob.ENC_STRING_DIFF, ob.ENC_STRING_PREFIX
    return [("hex_var", H, hexs, nl), ("hex_var_nop", H, hexs, nop), ("hex_fix", H, hexfix, None), ("hex_fix_null", H, hexfix, nl),
            ("hex_fix_to_var", H, hexfix, many_null),
            ("diff_hex", D, diffs, None), ("diff_hex_null", D, diffs, nl), ("diff_raw", D, diffs2, nl), ("diff_one_byte", D, diff1, nop),
            ("diff_var_store", D, diffs, many_null), ("diff_long_runs", D, long_diff, nl),
            ("prefix", P, pre, nl), ("prefix_hex", P, prehex, None), ("prefix_nop", P, pre, nop), ("prefix_many", P, many, nl)]


CASES = codec_cases()


@pytest.mark.parametrize("case", range(len(CASES)), ids=[c[0] for c in CASES])
@pytest.mark.parametrize("extra_var_col", [False, True])
def test_writer_to_oracle_round_trip(case, extra_var_col):
    name, enc, vals, nulls = CASES[case]
    n = len(vals)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)), ob.Column(ob.OBJ_VARCHAR, enc, vals, nulls=nulls)]
    if extra_var_col:   # a second var-stored column: the row then carries a column index array (locate_cell_data's general branch)
        cols.insert(1, ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, [b"x" * (i % 5) for i in range(n)]))
        cols.append(ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, [b"tail%d" % (i % 7) for i in range(n)]))
    col = 2 if extra_var_col else 1
    blk = ora.Block(ob.encode_block(cols))
    assert blk.verify_checksums() == 0
    for r in range(n):
        want = None if (nulls is not None and nulls[r]) else vals[r]
        assert blk.cell(col, r) == want, (name, r)
    if extra_var_col:
        assert blk.cell(1, 17) == b"x" * 2 and blk.cell(3, 17) == b"tail3"
    # batch path == cell path (test_general_column_decoder.cpp: batch_decode_to_datum_test)
    rid = np.arange(0, n, 3, dtype=np.int32)
    ptrs, lens, nl, _ = blk.get_rows_discrete(col, rid, absolute=True)
    import ctypes as C
    for i, r in enumerate(rid):
        isnull = bool((int(nl[i // 64]) >> (i % 64)) & 1)
        assert isnull == (nulls is not None and bool(nulls[r]))
        if not isnull:
            assert C.string_at(int(ptrs[i]), int(lens[i])) == vals[r]
    ora.arena_reset()


def test_hex_store_order_kat():
    """test_hex.cpp:24-47 (ObHexStringMap store_order): "0123456789" packs to the bytes 01 23 45 67 89 -- indexes follow byte order,
    the first character of a pair sits in the high nibble."""
    vals = [b"0123456789", b"9876543210", b"0123456789"]
    blk = ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_HEX_PACKING, vals)])
    raw = bytes(blk)
    assert bytes([0x01, 0x23, 0x45, 0x67, 0x89]) + bytes([0x98, 0x76, 0x54, 0x32, 0x10]) + bytes([0x01, 0x23, 0x45, 0x67, 0x89]) in raw
    assert b"0123456789" in raw     # the alphabet in the meta, in byte order
    b = ora.Block(blk)
    assert [b.cell(0, r) for r in range(3)] == vals


def test_not_suitable_inputs_are_refused():
    n = 50
    rng = np.random.default_rng(3)
    wide = [bytes(rng.integers(32, 127, size=9, dtype=np.uint8)) for _ in range(n)]     # > 16 distinct bytes
    with pytest.raises(ob.ObGpuError):
        ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_HEX_PACKING, wide)])
    with pytest.raises(ob.ObGpuError):   # different lengths
        ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_STRING_DIFF, [b"ab", b"abc"] * 10)])
    with pytest.raises(ob.ObGpuError):   # nothing differs
        ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_STRING_DIFF, [b"same"] * 10)])
    with pytest.raises(ob.ObGpuError):   # nothing in common
        ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_STRING_DIFF, [b"ab", b"cd"] * 10)])
    with pytest.raises(ob.ObGpuError):   # no shared prefix at all
        ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_STRING_PREFIX, [bytes([65 + i]) for i in range(20)])])
    with pytest.raises(ob.ObGpuError):   # integer column
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_HEX_PACKING, np.arange(n, dtype=np.int64))])


@pytest.mark.parametrize("case", [0, 3, 6, 7, 11, 12], ids=lambda i: CASES[i][0])
def test_filters_and_scan_over_rebuilt_strings(case):
    """White filters have no pushdown on these codecs: the reference decodes each row and compares (filter_pushdown_retro,
    ob_micro_block_decoder.cpp:1593-1678). Oracle vs a Python model, through the block filter and the whole-table scan."""
    name, enc, vals, nulls = CASES[case]
    n = len(vals)
    k = np.arange(n, dtype=np.int64)
    table = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_RAW, k), ob.Column(ob.OBJ_VARCHAR, enc, vals, nulls=nulls)], 90)
    isnull = np.zeros(n, dtype=bool) if nulls is None else nulls.astype(bool)
    present = sorted({v for v, z in zip(vals, isnull) if not z})
    lo, hi = present[len(present) // 4], present[3 * len(present) // 4]
    model = {
        "eq": (ob.White(1, ob.WHITE_OP_EQ, (present[3],)), lambda v: v == present[3]),
        "ne": (ob.White(1, ob.WHITE_OP_NE, (present[3],)), lambda v: v != present[3]),
        "lt": (ob.White(1, ob.WHITE_OP_LT, (hi,)), lambda v: v < hi),
        "bt": (ob.White(1, ob.WHITE_OP_BT, (lo, hi)), lambda v: lo <= v <= hi),
        "in": (ob.White(1, ob.WHITE_OP_IN, (present[0], present[-1], b"nope")), lambda v: v in (present[0], present[-1])),
    }
    for tag, (flt, fn) in model.items():
        res = ora.scan_table(table, flt, [0, 1], [False, True], [8, 8], string_base=table.image.ctypes.data)
        want_rows = [i for i in range(n) if not isnull[i] and fn(vals[i])]
        assert np.array_equal(res["data"][0].view(np.int64), k[want_rows]), (name, tag)
        assert ora.scan_strings(table, res, 1, table.image.ctypes.data) == [vals[i] for i in want_rows]
    res = ora.scan_table(table, ob.White(1, ob.WHITE_OP_NU, ()), [0], [False], [8])
    assert np.array_equal(res["data"][0].view(np.int64), k[isnull])
    ora.arena_reset()
