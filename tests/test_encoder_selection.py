"""OBGPU_ENC_AUTO: the writer picks a column's codec per micro-block the way ObMicroBlockEncoder::fast_encoder_detect /
choose_encoder do (encoding/ob_micro_block_encoder.cpp:1318-1366,1603-1823) among RAW / DICT / RLE / CONST / INTEGER_BASE_DIFF, from
the encoders' own calc_size() estimates (ob_raw_encoder.cpp:256, ob_dict_encoder.h:98-125, ob_rle_encoder.cpp:118, ob_const_encoder.cpp:133,
ob_integer_base_diff_encoder.cpp:254). Checked here: the choice on shapes where the estimates are far apart (the ordering of the
reference's candidates decides, not a tie), the round trip of every block through the oracle's decoders, and that a block written with
the chosen codec forced is byte-identical to the AUTO block (selection adds nothing to the format)."""
import numpy as np
import pytest

import oracle_binding as ora

NAMES = {0: "RAW", 1: "DICT", 2: "RLE", 3: "CONST", 4: "INTEGER_BASE_DIFF"}


def col_type(block, i):
    return int(block[64 + 16 * i + 1])


def shapes(n, rng):
    big = rng.integers(1 << 38, 1 << 39, size=8, dtype=np.int64)
    out = {}
    out["constant"] = (np.full(n, 777_777, dtype=np.int64), None, "CONST")
    v = np.full(n, 42, dtype=np.int64); v[[5, n // 2, n - 2]] = [9, 10, 11]
    out["const_with_3_exceptions"] = (v, None, "CONST")
    out["all_null"] = (np.zeros(n, dtype=np.int64), np.ones(n, dtype=np.uint8), "CONST")
    out["low_cardinality_shuffled"] = (big[rng.integers(0, 8, size=n)], None, "DICT")
    out["long_runs"] = (np.repeat(big, n // 8 + 1)[:n].copy(), None, "RLE")
    out["narrow_range_on_a_large_base"] = (np.int64(1) * (10**12) + np.sort(rng.choice(1 << 20, size=n, replace=False)).astype(np.int64), None,
                                           "INTEGER_BASE_DIFF")
    out["random_64_bit"] = (rng.integers(-2**63, 2**63 - 1, size=n, dtype=np.int64), None, "RAW")
    out["random_13_bit"] = (rng.integers(0, 1 << 13, size=n, dtype=np.int64), None, "RAW")
    nl = (rng.random(n) < 0.3).astype(np.uint8)
    out["low_cardinality_with_nulls"] = (big[rng.integers(0, 4, size=n)], nl, "DICT")
    return out


@pytest.mark.parametrize("n", [1000, 257])
def test_choice_on_clear_shapes_and_round_trip(n):
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_block
    rng = np.random.default_rng(11)
    for name, (v, nl, want) in shapes(n, rng).items():
        cols = [Column(capi.OBJ_INT, capi.ENC_AUTO, v, nulls=nl)]
        block = encode_block(cols)
        got = NAMES[col_type(block, 0)]
        assert got == want, (name, n, got, want)
        blk = ora.Block(block)
        assert blk.verify_checksums() == 0
        for r in list(range(0, n, 37)) + [n - 1]:
            cell = blk.cell(0, r)
            if nl is not None and nl[r]:
                assert cell is None, (name, r)
            else:
                assert (cell & 0xffffffffffffffff) == (int(v[r]) & 0xffffffffffffffff), (name, r, cell, int(v[r]))
        # the block equals the one written with the chosen codec forced
        forced = encode_block([Column(capi.OBJ_INT, col_type(block, 0), v, nulls=nl)])
        assert np.array_equal(block, forced), name


def test_strings_and_mixed_table():
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_table
    rng = np.random.default_rng(12)
    n = 5000
    words = [b"alpha", b"beta-beta", b"gamma", b"delta!!"]
    s_low = [words[i] for i in rng.integers(0, 4, size=n)]
    s_uniq = [b"row-%07d-%s" % (i, b"x" * int(rng.integers(0, 9))) for i in range(n)]
    s_const = [b"same"] * n
    key = np.arange(n, dtype=np.int64) * 7 + (1 << 41)
    cols = [Column(capi.OBJ_INT, capi.ENC_AUTO, key), Column(capi.OBJ_VARCHAR, capi.ENC_AUTO, s_low),
            Column(capi.OBJ_VARCHAR, capi.ENC_AUTO, s_uniq), Column(capi.OBJ_VARCHAR, capi.ENC_AUTO, s_const),
            Column(capi.OBJ_INT32, capi.ENC_AUTO, rng.integers(-5, 6, size=n, dtype=np.int64))]
    t = encode_table(cols, 800, rowkey_cnt=1)
    want = ["INTEGER_BASE_DIFF", "DICT", "RAW", "CONST", "DICT"]
    for b in range(t.n_blocks):
        block = t.block(b)
        got = [NAMES[col_type(block, i)] for i in range(5)]
        assert got == want, (b, got)
        blk = ora.Block(block)
        assert blk.verify_checksums() == 0
        r0 = b * 800
        for r in (0, 17, blk.row_count - 1):
            assert (blk.cell(0, r) & 0xffffffffffffffff) == int(key[r0 + r])
            assert blk.cell(1, r) == s_low[r0 + r]
            assert blk.cell(2, r) == s_uniq[r0 + r]
            assert blk.cell(3, r) == s_const[r0 + r]


def test_random_tables_round_trip():
    """Whatever the choice, the block decodes to the input (the decision only selects among codecs the writer already has)."""
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_block
    rng = np.random.default_rng(13)
    seen = set()
    for trial in range(120):
        n = int(rng.integers(1, 600))
        card = int(rng.choice([1, 2, 5, 50, 10_000]))
        bits = int(rng.choice([1, 7, 16, 33, 62]))
        pool = rng.integers(0, 1 << bits, size=card, dtype=np.int64) + int(rng.choice([0, 10**9, -(1 << 40)]))
        v = pool[rng.integers(0, card, size=n)]
        if rng.random() < 0.4:
            v = np.sort(v)
        nl = (rng.random(n) < rng.choice([0.0, 0.1, 0.9])).astype(np.uint8)
        nl = nl if nl.any() else None
        t = int(rng.choice([capi.OBJ_INT, capi.OBJ_UINT64])) if v.min() >= 0 else capi.OBJ_INT
        block = encode_block([Column(t, capi.ENC_AUTO, v, nulls=nl)])
        seen.add(col_type(block, 0))
        blk = ora.Block(block)
        for r in range(0, n, max(1, n // 25)):
            cell = blk.cell(0, r)
            if nl is not None and nl[r]:
                assert cell is None
            else:
                assert (cell & 0xffffffffffffffff) == (int(v[r]) & 0xffffffffffffffff), (trial, r)
    assert seen >= {0, 1, 2, 3, 4}, seen


CS_NAMES = {0: "INTEGER", 1: "STRING", 2: "INT_DICT", 3: "STR_DICT"}


def cs_col_type(block, i, ncol):
    # [micro header 64][ObAllColumnHeader 12][ObCSColumnHeader x ncol: version_, type_, attrs_, obj_type_]
    return int(block[64 + 12 + 4 * i + 1])


def test_cs_auto_integer_vs_dict_and_string_vs_dict():
    """OBGPU_ENC_CS_AUTO: ObMicroBlockCSEncoder::choose_encoder_for_integer_ / _for_string_ (ob_micro_block_cs_encoder.cpp:2289-2375) over the
    column encoders' estimate_store_size(): dictionary form below 70 % of the plain estimate, or below it with < rows / 2 distinct values."""
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_block
    rng = np.random.default_rng(14)
    n = 1200
    big = rng.integers(1 << 40, 1 << 41, size=6, dtype=np.int64)
    words = [b"north", b"south-south", b"east", b"westwestwest"]
    nl = (rng.random(n) < 0.2).astype(np.uint8)
    shapes = [
        ("low_card_wide", capi.OBJ_INT, big[rng.integers(0, 6, size=n)], None, "INT_DICT"),
        ("unique_wide", capi.OBJ_INT, rng.integers(0, 1 << 41, size=n, dtype=np.int64), None, "INTEGER"),
        ("constant", capi.OBJ_INT, np.full(n, 777_777, dtype=np.int64), None, "INT_DICT"),
        ("two_bit_values", capi.OBJ_INT, rng.integers(0, 4, size=n, dtype=np.int64), None, "INTEGER"),
        ("low_card_with_nulls", capi.OBJ_INT, big[rng.integers(0, 3, size=n)], nl, "INT_DICT"),
        ("negative_narrow_range", capi.OBJ_INT, rng.integers(-40, 40, size=n, dtype=np.int64), None, "INTEGER"),
        ("low_card_strings", capi.OBJ_VARCHAR, [words[i] for i in rng.integers(0, 4, size=n)], None, "STR_DICT"),
        ("unique_strings", capi.OBJ_VARCHAR, [b"row-%06d-%s" % (i, b"y" * int(rng.integers(0, 7))) for i in range(n)], None, "STRING"),
        ("strings_with_nulls", capi.OBJ_VARCHAR, [words[i] for i in rng.integers(0, 2, size=n)], nl, "STR_DICT"),
    ]
    forced = {"INTEGER": capi.ENC_CS_INTEGER, "INT_DICT": capi.ENC_CS_INT_DICT, "STRING": capi.ENC_CS_STRING, "STR_DICT": capi.ENC_CS_STR_DICT}
    for name, t, v, nulls, want in shapes:
        block = encode_block([Column(t, capi.ENC_CS_AUTO, v, nulls=nulls)])
        got = CS_NAMES[cs_col_type(block, 0, 1)]
        assert got == want, (name, got, want)
        assert np.array_equal(block, encode_block([Column(t, forced[got], v, nulls=nulls)])), name
        blk = ora.Block(block)
        assert blk.verify_checksums() == 0
        for r in list(range(0, n, 53)) + [n - 1]:
            cell = blk.cell(0, r)
            if nulls is not None and nulls[r]:
                assert cell is None, (name, r)
            elif isinstance(v, list):
                assert bytes(cell) == v[r], (name, r)
            else:
                assert (cell & 0xffffffffffffffff) == int(v[r]) & 0xffffffffffffffff, (name, r)
    # a whole table: every column resolved independently, per micro-block
    from oceanbase_b200.sstable import encode_table
    cols = [Column(t, capi.ENC_CS_AUTO, v, nulls=nulls) for (_n, t, v, nulls, _w) in shapes]
    table = encode_table(cols, 400)
    for b in range(table.n_blocks):
        got = [CS_NAMES[cs_col_type(table.block(b), i, len(cols))] for i in range(len(cols))]
        assert got == [w for (*_x, w) in shapes], (b, got)
