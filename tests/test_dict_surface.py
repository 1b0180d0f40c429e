"""Dictionary surface of a micro-block (distinct count / read_distinct / read_reference; pushdown GROUP BY, black filter on
one dictionary column). CPU half: the oracle's three helpers agree with its own cell decoder on every dictionary-coded
codec (entry[ref[row]] is the cell, NULL rows carry ref == distinct count), the property the reference's group-by tests
check (unittest/storage/blocksstable/encoding/test_dict_decoder.cpp group-by cases, ob_dict_decoder.cpp:1681-1830)."""
import numpy as np
import pytest

import oracle_binding as ora


def dict_cases(ob, n=700, seed=5):
    """(name, columns): column 0 is the dictionary-coded one, column 1 a value column, column 2 a second value column"""
    rng = np.random.default_rng(seed)
    nl = (rng.random(n) < 0.15).astype(np.uint8)
    ints = rng.integers(-20, 20, size=n) * 10 ** 9
    runs = np.repeat(rng.integers(0, 9, size=n // 10 + 1), 10)[:n] * 77
    strs = [b"k%03d" % i + b"z" * (i % 7) for i in range(30)]
    sv = [strs[i] for i in rng.integers(0, 30, size=n)]
    mostly = np.where(rng.random(n) < 0.97, 5, rng.integers(0, 4, size=n)).astype(np.int64)
    v1 = rng.integers(-(1 << 40), 1 << 40, size=n, dtype=np.int64)
    v2 = rng.integers(0, 1 << 31, size=n, dtype=np.int64)
    vn = (rng.random(n) < 0.1).astype(np.uint8)
    tail = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, v1, nulls=vn), ob.Column(ob.OBJ_INT32, ob.ENC_RAW, v2)]
    cs_tail = [ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, v1, nulls=vn), ob.Column(ob.OBJ_INT32, ob.ENC_CS_INTEGER, v2)]
    return [
        ("dict_int", [ob.Column(ob.OBJ_INT, ob.ENC_DICT, ints, nulls=nl)] + tail),
        ("dict_int_nonull", [ob.Column(ob.OBJ_INT, ob.ENC_DICT, ints)] + tail),
        ("dict_str", [ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, sv, nulls=nl)] + tail),
        ("rle_int", [ob.Column(ob.OBJ_INT, ob.ENC_RLE, runs, nulls=np.repeat(nl[::10], 10)[:n])] + tail),
        ("rle_str", [ob.Column(ob.OBJ_VARCHAR, ob.ENC_RLE, [strs[i // 12 % 30] for i in range(n)])] + tail),
        ("const_int", [ob.Column(ob.OBJ_INT, ob.ENC_CONST, mostly)] + tail),
        ("const_plain", [ob.Column(ob.OBJ_INT, ob.ENC_CONST, np.full(n, 42, dtype=np.int64))] + tail),
        ("cs_int_dict", [ob.Column(ob.OBJ_INT, ob.ENC_CS_INT_DICT, ints, nulls=nl)] + cs_tail),
        ("cs_str_dict", [ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STR_DICT, sv, nulls=nl)] + cs_tail),
    ]


def group_by_model(blk, rows, aggs):
    """{ref: [per agg (a, b)]} folded in Python integers from the oracle's refs and cells"""
    refs = blk.dict_refs(0, rows)
    n_groups = blk.dict_count(0) + 1
    out = np.zeros((len(aggs), n_groups, 2), dtype=np.int64)
    acc = [[None] * n_groups for _ in aggs]
    for r, ref in zip(rows, refs):
        for k, (kind, col) in enumerate(aggs):
            if kind == 0 and col < 0:
                acc[k][ref] = (acc[k][ref] or 0) + 1
                continue
            v = blk.cell(col, int(r))
            if v is None:
                continue
            if v >= 1 << 63:
                v -= 1 << 64
            cur = acc[k][ref]
            acc[k][ref] = {0: lambda: (cur or 0) + 1, 1: lambda: (cur or 0) + v, 3: lambda: v if cur is None else min(cur, v),
                           4: lambda: v if cur is None else max(cur, v)}[kind]()
    for k, (kind, col) in enumerate(aggs):
        for g in range(n_groups):
            a = acc[k][g]
            if kind in (0, 1):
                a = a or 0
                out[k, g, 0] = np.array(a & ((1 << 64) - 1), dtype=np.uint64).view(np.int64)
                out[k, g, 1] = np.array((a >> 64) & ((1 << 64) - 1), dtype=np.uint64).view(np.int64) if kind == 1 else 0
            elif a is not None:
                out[k, g] = (a, 1)
    return out


@pytest.mark.parametrize("case", range(9))
def test_oracle_dictionary_surface_agrees_with_cell_decode(case):
    import oceanbase_b200 as ob
    name, cols = dict_cases(ob)[case]
    blk = ora.Block(ob.encode_block(cols))
    n = blk.row_count
    count = blk.dict_count(0)
    refs = blk.dict_refs(0, np.arange(n, dtype=np.int32))
    entries = [blk.dict_entry(0, g)[0] for g in range(count)]
    assert len(set(entries)) == count, name   # the writer's dictionaries hold each value once
    seen_null = False
    for r in range(n):
        cell = blk.cell(0, r)
        if cell is None:
            assert refs[r] == count
            seen_null = True
        else:
            assert refs[r] < count and entries[refs[r]] == cell
    assert seen_null == (cols[0].nulls is not None and bool(np.any(cols[0].nulls)))
    # a column that is not dictionary coded is refused
    with pytest.raises(RuntimeError):
        blk.dict_count(1)
