"""Randomised CPU round trip: random schemas over every codec of both block formats (the generators of
tests/test_gpu_fuzz.py) -> writer -> oracle decode must give back the INPUT cells, and white filters evaluated by the
oracle must select exactly the rows an independent numpy / Python model selects. This closes the loop the GPU tests leave
open (device == oracle): oracle == ground truth."""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
import test_gpu_fuzz as F


def model_value(t, x):
    """The comparable value of a non-NULL cell: Python int in the column's own order, or bytes."""
    if isinstance(x, (bytes, bytearray)):
        return bytes(x)
    x = int(x)
    if t in (ob.OBJ_UINT64, ob.OBJ_UINT32):
        return x % (1 << 64)
    return x


def model_leaf(meta, leaf, row_range):
    t, is_str, v, nulls = meta[leaf.col]
    op, params = leaf.op, leaf.params
    out = np.zeros(len(row_range), dtype=bool)
    consts = [None if p is None else model_value(t, p) for p in params]
    for k, r in enumerate(row_range):
        isnull = nulls is not None and nulls[r]
        if op == ob.WHITE_OP_NU:
            out[k] = isnull
            continue
        if op == ob.WHITE_OP_NN:
            out[k] = not isnull
            continue
        if isnull:
            continue
        x = model_value(t, v[r])
        if op == ob.WHITE_OP_IN:
            out[k] = any(c is not None and x == c for c in consts)
            continue
        if any(c is None for c in consts):
            continue                                   # a NULL constant never matches
        if op == ob.WHITE_OP_BT:
            out[k] = consts[0] <= x <= consts[1]
        else:
            c = consts[0]
            out[k] = {ob.WHITE_OP_EQ: x == c, ob.WHITE_OP_NE: x != c, ob.WHITE_OP_LT: x < c, ob.WHITE_OP_LE: x <= c,
                      ob.WHITE_OP_GT: x > c, ob.WHITE_OP_GE: x >= c}[op]
    return out


def model_tree(meta, node, row_range):
    if isinstance(node, ob.White):
        return model_leaf(meta, node, row_range)
    kids = [model_tree(meta, c, row_range) for c in node.children]
    acc = kids[0].copy()
    for k in kids[1:]:
        acc = (acc & k) if isinstance(node, ob.And) else (acc | k)
    return acc


@pytest.mark.parametrize("seed", range(150))
def test_writer_oracle_roundtrip_and_filters(seed):
    rng, cs, n, rpb, cols, meta = F.random_case(ob, 5000 + seed)
    n_eff = min(n, 2500)                      # Python-level model: keep it quick
    for c in cols:
        if isinstance(c.values, list):
            c.values = c.values[:n_eff]
        else:
            c.values = c.values[:n_eff].copy()
        if c.nulls is not None:
            c.nulls = c.nulls[:n_eff].copy()
    meta = [(t, s, (v[:n_eff] if isinstance(v, list) else v[:n_eff]), None if nl is None else nl[:n_eff]) for t, s, v, nl in meta]
    table = F.encode_or_relax(ob, cols, rpb)
    if table is None:
        pytest.skip("the writer does not produce these forced encodings for this data")
    flts = [F.random_filter(ob, rng, meta, n_eff) for _ in range(3)]
    for b in range(table.n_blocks):
        blk = ora.Block(table.block(b))
        assert blk.verify_checksums() == 0
        lo = b * rpb
        rows = range(lo, lo + blk.row_count)
        for ci, (t, is_str, v, nulls) in enumerate(meta):
            for r in list(rows)[::max(1, blk.row_count // 40)]:
                d = blk.cell_raw(ci, r - lo)
                if nulls is not None and nulls[r]:
                    assert d.is_null == 1, (seed, ci, r)
                    continue
                assert d.is_null == 0, (seed, ci, r)
                if is_str:
                    assert blk.cell(ci, r - lo) == v[r], (seed, ci, r)
                elif t == ob.OBJ_DATE:
                    assert d.len == 4 and int(np.int32(np.uint32(d.ival & 0xffffffff))) == int(v[r]), (seed, ci, r)
                else:
                    assert int(np.int64(np.uint64(d.ival))) == int(v[r]), (seed, ci, r)
        for flt in flts:
            got = blk.filter_tree(flt).astype(bool)
            assert np.array_equal(got, model_tree(meta, flt, rows)), (seed, b, flt)
