"""COLUMN_EQUAL / COLUMN_SUBSTR (the span columns): a column stored as "equals column R" / "is a substring of column R in
the same row", plus an exception list (encoding/ob_column_equal_decoder.cpp:32-133, ob_inter_column_substring_decoder.cpp:32-93,
exception rows: ObBitMapMetaReader, ob_encoding_bitset.h:574-760). CPU half: the writer's encoders against the oracle's decoders on
every layout of the exception meta (none / bit-packed / byte-packed integers, fixed / var-length strings, NULL and NOP exceptions),
every row-store shape of COLUMN_SUBSTR (same start, fixed length, both, neither, 1- and 2-byte fields), the referenced column under
each ordinary codec, the "not suitable" limits, and white filters / scans over such columns (the reference's retro path)."""
import ctypes as C

import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora

U64 = (1 << 64) - 1


def want_cell(vals, nulls, r, is_str):
    if nulls is not None and nulls[r]:
        return None
    return vals[r] if is_str else int(vals[r]) & U64


def equal_cases(n=400, seed=5):
    """(name, obj_type, ref encoding, ref values, ref nulls, values, nulls)"""
    rng = np.random.default_rng(seed)
    out = []
    base = rng.integers(-1000, 1000, size=n).astype(np.int64)
    out.append(("int_no_exc", ob.OBJ_INT, ob.ENC_RAW, base, None, base.copy(), None))
    v = base.copy(); ex = rng.choice(n, 20, replace=False); v[ex] = rng.integers(0, 1 << 20, size=20)
    out.append(("int_bitpacked_exc", ob.OBJ_INT, ob.ENC_RAW, base, None, v, None))
    v = base.copy(); v[ex] = rng.integers(-(1 << 40), 0, size=20)
    out.append(("int_negative_exc_8_bytes", ob.OBJ_INT, ob.ENC_DICT, base, None, v, None))
    v = base.copy(); v[ex] = rng.integers(0, 1 << 16, size=20) | 0x8000
    out.append(("int_byte_packed_exc", ob.OBJ_INT, ob.ENC_RAW, base, None, v, None))
    b32 = rng.integers(-50, 50, size=n).astype(np.int64)
    v = b32.copy(); v[ex] = -rng.integers(1, 1 << 20, size=20)
    out.append(("int32_negative_exc_sign_mask", ob.OBJ_INT32, ob.ENC_RAW, b32, None, v, None))
    rn = (rng.random(n) < 0.15).astype(np.uint8)
    v = base.copy(); nl = rn.copy(); nl[ex[:7]] ^= 1; v[ex[7:]] = 77
    out.append(("int_null_exceptions", ob.OBJ_INT, ob.ENC_RAW, base, rn, v, nl))
    nop = rn.copy(); nop[ex[:5]] = 2
    out.append(("int_nop_exceptions", ob.OBJ_INT, ob.ENC_RLE, np.sort(base), rn, np.sort(base), nop))
    u = rng.integers(0, 1 << 62, size=n).astype(np.int64)
    v = u.copy(); v[ex] = rng.integers(0, 5, size=20)
    out.append(("uint_small_exc", ob.OBJ_UINT64, ob.ENC_RAW, u, None, v, None))
    words = [b"alpha", b"beta", b"gamma-long-value", b"", b"delta"]
    s = [words[i] for i in rng.integers(0, len(words), size=n)]
    out.append(("str_no_exc", ob.OBJ_VARCHAR, ob.ENC_DICT, s, None, list(s), None))
    t = list(s)
    for i in ex: t[i] = b"EXC%05d" % i
    out.append(("str_fixed_exc", ob.OBJ_VARCHAR, ob.ENC_RAW, s, None, t, None))
    t = list(s)
    for i in ex: t[i] = b"x" * int(rng.integers(0, 30)) + b"!"
    out.append(("str_var_exc", ob.OBJ_VARCHAR, ob.ENC_DICT, s, None, t, None))
    t = list(s)
    for i in ex: t[i] = b"y" * int(rng.integers(10, 40))
    out.append(("str_var_exc_2_byte_index", ob.OBJ_VARCHAR, ob.ENC_RAW, s, None, t, None))
    nl = rn.copy(); nl[ex[:6]] ^= 1
    t = list(s)
    for i in ex[6:]: t[i] = b"z" * (int(i) % 4)
    out.append(("str_null_exceptions", ob.OBJ_VARCHAR, ob.ENC_RAW, s, rn, t, nl))
    nl = np.zeros(n, dtype=np.uint8); nl[ex[:3]] = 1
    out.append(("str_only_null_exceptions", ob.OBJ_VARCHAR, ob.ENC_RAW, s, None, list(s), nl))
    one = list(s); one[ex[0]] = b"q" * 300
    out.append(("str_one_long_exception", ob.OBJ_VARCHAR, ob.ENC_RAW, s, None, one, None))
    cst = [b"const"] * n
    t = list(cst); t[5] = b"other"
    out.append(("str_ref_const", ob.OBJ_VARCHAR, ob.ENC_CONST, cst, None, t, None))
    return out


EQ = equal_cases()


@pytest.mark.parametrize("case", range(len(EQ)), ids=[c[0] for c in EQ])
def test_column_equal_round_trip(case):
    name, ot, renc, rv, rn, v, nl = EQ[case]
    n = len(v)
    is_str = ot == ob.OBJ_VARCHAR
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)),
            ob.Column(ot, ob.ENC_COLUMN_EQUAL, v, nulls=nl, ref_col=2),
            ob.Column(ot, renc, rv, nulls=rn)]
    blk = ora.Block(ob.encode_block(cols))
    assert blk.verify_checksums() == 0
    for r in range(n):
        got = blk.cell(1, r)
        want = want_cell(v, nl, r, is_str)
        if not is_str and want is not None and ot == ob.OBJ_INT32:
            want = int(v[r]) & U64   # INT32 datums are 8 bytes long (sign extended)
        assert got == want, (name, r)
        assert blk.cell(2, r) == want_cell(rv, rn, r, is_str)
    rid = np.arange(0, n, 3, dtype=np.int32)
    if is_str:
        ptrs, lens, nb, _ = blk.get_rows_discrete(1, rid, absolute=True)
        for i, r in enumerate(rid):
            isnull = bool((int(nb[i // 64]) >> (i % 64)) & 1)
            assert isnull == (nl is not None and bool(nl[r]))
            if not isnull:
                assert C.string_at(int(ptrs[i]), int(lens[i])) == v[r]
    else:
        data, nb, _ = blk.get_rows_fixed(1, rid)
        for i, r in enumerate(rid):
            isnull = bool((int(nb[i // 64]) >> (i % 64)) & 1)
            assert isnull == (nl is not None and bool(nl[r]))
            if not isnull:
                assert int(data.view(np.int64)[i]) == int(v[r])
    ora.arena_reset()


def substr_cases(n=400, seed=9):
    """(name, ref values, ref nulls, values, nulls)"""
    rng = np.random.default_rng(seed)
    out = []
    ref = [b"user-%06d@host%02d.example.com" % (rng.integers(0, 10 ** 6), rng.integers(0, 100)) for _ in range(n)]
    out.append(("same_start_fixed_len", ref, None, [r[5:11] for r in ref], None))          # no row store at all
    out.append(("same_start_var_len", ref, None, [r[:5 + i % 7] for i, r in enumerate(ref)], None))
    var = [b"k" * int(rng.integers(0, 9)) + b"#" + bytes(rng.integers(97, 123, size=8, dtype=np.uint8)) + b"#tail" for _ in range(n)]
    out.append(("var_start_fixed_len", var, None, [r[r.index(b"#"):r.index(b"#") + 9] for r in var], None))
    out.append(("var_start_var_len", var, None, [r[r.index(b"#") + 1:] for r in var], None))
    ex = rng.choice(n, 15, replace=False)
    v = [r[5:11] for r in ref]
    for i in ex: v[i] = b"not-a-substring-%d" % i
    out.append(("exceptions_var", ref, None, v, None))
    v = [r[5:11] for r in ref]
    for i in ex: v[i] = b"??????"
    out.append(("exceptions_fixed", ref, None, v, None))
    rn = (rng.random(n) < 0.1).astype(np.uint8)
    nl = rn.copy(); nl[ex[:5]] ^= 1
    v = [r[3:3 + i % 5] for i, r in enumerate(var)]
    out.append(("nulls_both_and_either", var, rn, v, nl))
    nop = rn.copy(); nop[(rng.random(n) < 0.05)] = 2
    out.append(("nop_cells", var, nop, v, nop))
    long_ref = [b"p" * int(rng.integers(250, 400)) + b"|" + b"%04d" % i + b"|" for i in range(n)]
    out.append(("two_byte_start", long_ref, None, [r[r.index(b"|"):] for r in long_ref], None))
    out.append(("two_byte_length", long_ref, None, [r[1:int(rng.integers(1, 300))] for r in long_ref], None))
    out.append(("empty_values", ref, None, [b""] * n, None))
    out.append(("whole_value", ref, None, list(ref), None))
    return out


SUB = substr_cases()


@pytest.mark.parametrize("case", range(len(SUB)), ids=[c[0] for c in SUB])
@pytest.mark.parametrize("renc", [ob.ENC_RAW, ob.ENC_DICT])
def test_column_substr_round_trip(case, renc):
    name, rv, rn, v, nl = SUB[case]
    n = len(v)
    cols = [ob.Column(ob.OBJ_VARCHAR, renc, rv, nulls=rn),
            ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)),
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_COLUMN_SUBSTR, v, nulls=nl, ref_col=0)]
    blk = ora.Block(ob.encode_block(cols))
    assert blk.verify_checksums() == 0
    for r in range(n):
        assert blk.cell(2, r) == want_cell(v, nl, r, True), (name, r)
    rid = np.arange(n - 1, -1, -2, dtype=np.int32)
    ptrs, lens, nb, _ = blk.get_rows_discrete(2, rid, absolute=True)
    for i, r in enumerate(rid):
        isnull = bool((int(nb[i // 64]) >> (i % 64)) & 1)
        assert isnull == (nl is not None and bool(nl[r]))
        if not isnull:
            assert C.string_at(int(ptrs[i]), int(lens[i])) == v[r]
    ora.arena_reset()


def test_bitset_get_ref_kat():
    """unittest/storage/blocksstable/encoding/test_bitset.cpp:24-87 (BitSet, set_get): bits set at strides of 1..5 over 50 and 400
    positions; get_ref of a set bit is its index in the list of set positions, a clear bit has none. The writer's exception BitSet
    (bytes, little-endian words) is what the oracle's rank reads."""
    rng = np.random.default_rng(7)
    for cnt, words in ((50, 1), (400, 8)):
        for _ in range(200):
            buf = np.zeros(words, dtype=np.uint64)
            pos, i = [], 0
            while i < cnt:
                buf[i // 64] |= np.uint64(1) << np.uint64(i % 64)
                pos.append(i)
                i += int(rng.integers(1, 6))
            for i in range(cnt):
                got = ora.oracle().ora_bitset_get_ref(buf.ctypes.data, i)
                assert got == (pos.index(i) if i in pos else -1)
    # the writer lays the same bits down: exception rows 3, 64, 65, 130 of 200
    k = np.arange(200, dtype=np.int64)
    v = k.copy(); v[[3, 64, 65, 130]] = -7
    raw = bytes(ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, k), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, v, ref_col=0)]))
    want = np.zeros(4, dtype=np.uint64)
    for r in (3, 64, 65, 130):
        want[r // 64] |= np.uint64(1) << np.uint64(r % 64)
    assert want.tobytes() in raw


def test_substr_header_shapes():
    """ObInterColSubStrMetaHeader (ob_inter_column_substring_encoder.h:26-60): which fields go to the header, which to the rows."""
    ref = [b"abcdefghij%02d" % i for i in range(40)]

    def header(vals):
        blk = ob.encode_block([ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, ref), ob.Column(ob.OBJ_VARCHAR, ob.ENC_COLUMN_SUBSTR, vals, ref_col=0)])
        return bytes(blk), ora.Block(blk)

    raw, b = header([r[2:6] for r in ref])
    assert [b.cell(1, r) for r in range(40)] == [r[2:6] for r in ref]
    # same start 2, fixed length 4, ref column 0: attr = same | fix, no per-row bytes
    assert bytes([0, 0x30, 2, 0, 4, 0, 0, 0]) in raw
    raw, b = header([r[i % 3:6] for i, r in enumerate(ref)])
    assert bytes([0, 0x01 | (0x01 << 2), 0, 0, 0, 0, 0, 0]) in raw     # 1-byte start, 1-byte length, both per row
    assert b.cell(1, 4) == ref[4][1:6]


def test_span_limits_are_refused():
    n = 100
    k = np.arange(n, dtype=np.int64)
    with pytest.raises(ob.ObGpuError):   # more than min(100, rows / 10 + 1) exceptions
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, k), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, k + (k % 5 == 0), ref_col=0)])
    with pytest.raises(ob.ObGpuError):   # refers to itself
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, k, ref_col=0)])
    with pytest.raises(ob.ObGpuError):   # different column types
        ob.encode_block([ob.Column(ob.OBJ_INT32, ob.ENC_RAW, k), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, k, ref_col=0)])
    with pytest.raises(ob.ObGpuError):   # a span column cannot refer to a span column
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, k), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, k, ref_col=0),
                         ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, k, ref_col=1)])
    with pytest.raises(ob.ObGpuError):   # COLUMN_SUBSTR is a string codec
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, k), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_SUBSTR, k, ref_col=0)])
    big = 2100   # the exception BitSet (one bit per row, 64-bit words) + the other arrays must fit the uint8 offsets
    kb = np.arange(big, dtype=np.int64)
    v = kb.copy(); v[7] = -1
    with pytest.raises(ob.ObGpuError):
        ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, kb), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, v, ref_col=0)])
    ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, kb), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, kb, ref_col=0)])   # no exceptions: any size


@pytest.mark.parametrize("kind", ["equal_int", "equal_str", "substr"])
def test_filters_and_scan_over_span_columns(kind):
    """White filters on a span column take the retro path (decode each row, compare: filter_pushdown_retro,
    ob_micro_block_decoder.cpp:1593-1678). Oracle vs a Python model through the whole-table scan."""
    n = 900
    rng = np.random.default_rng(21)
    k = np.arange(n, dtype=np.int64)
    if kind == "equal_int":
        rv = rng.integers(0, 50, size=n).astype(np.int64)
        v = rv.copy(); v[::97] = 1000 + k[::97]
        nl = np.zeros(n, dtype=np.uint8); nl[5::131] = 1
        cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, k), ob.Column(ob.OBJ_INT, ob.ENC_DICT, rv), ob.Column(ob.OBJ_INT, ob.ENC_COLUMN_EQUAL, v, nulls=nl, ref_col=1)]
        isnull = nl.astype(bool)
        tests = {"eq": (ob.White(2, ob.WHITE_OP_EQ, (7,)), lambda x: x == 7), "gt": (ob.White(2, ob.WHITE_OP_GT, (40,)), lambda x: x > 40),
                 "bt": (ob.White(2, ob.WHITE_OP_BT, (10, 20)), lambda x: 10 <= x <= 20), "in": (ob.White(2, ob.WHITE_OP_IN, (3, 1097, 5)), lambda x: x in (3, 1097, 5))}
        table = ob.encode_table(cols, 300)
        for tag, (flt, fn) in tests.items():
            res = ora.scan_table(table, flt, [0, 2], [False, False], [8, 8])
            rows = [i for i in range(n) if not isnull[i] and fn(int(v[i]))]
            assert np.array_equal(res["data"][0].view(np.int64), k[rows]), tag
            assert np.array_equal(res["data"][1].view(np.int64), v[rows]), tag
        res = ora.scan_table(table, ob.White(2, ob.WHITE_OP_NU, ()), [0], [False], [8])
        assert np.array_equal(res["data"][0].view(np.int64), k[isnull])
        return
    words = [b"red-apple", b"green-pear", b"blue-plum", b"red-cherry"]
    rv = [words[i] for i in rng.integers(0, 4, size=n)]
    if kind == "equal_str":
        v = list(rv)
        for i in range(0, n, 89): v[i] = b"exception-%d" % i
        enc = ob.ENC_COLUMN_EQUAL
    else:
        v = [r[:r.index(b"-")] for r in rv]
        for i in range(0, n, 89): v[i] = b"zzz"
        enc = ob.ENC_COLUMN_SUBSTR
    nl = np.zeros(n, dtype=np.uint8); nl[3::113] = 1
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, k), ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, rv), ob.Column(ob.OBJ_VARCHAR, enc, v, nulls=nl, ref_col=1)]
    table = ob.encode_table(cols, 300)
    isnull = nl.astype(bool)
    c0 = v[1]
    tests = {"eq": (ob.White(2, ob.WHITE_OP_EQ, (c0,)), lambda x: x == c0), "ne": (ob.White(2, ob.WHITE_OP_NE, (c0,)), lambda x: x != c0),
             "lt": (ob.White(2, ob.WHITE_OP_LT, (b"green",)), lambda x: x < b"green"),
             "in": (ob.White(2, ob.WHITE_OP_IN, (b"zzz", b"red", b"exception-89")), lambda x: x in (b"zzz", b"red", b"exception-89"))}
    for tag, (flt, fn) in tests.items():
        res = ora.scan_table(table, flt, [0, 2], [False, True], [8, 8], string_base=table.image.ctypes.data)
        rows = [i for i in range(n) if not isnull[i] and fn(v[i])]
        assert np.array_equal(res["data"][0].view(np.int64), k[rows]), tag
        assert ora.scan_strings(table, res, 1, table.image.ctypes.data) == [v[i] for i in rows], tag
    ora.arena_reset()
