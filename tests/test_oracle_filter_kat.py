"""The reference's own pushed-down-filter expectations, re-expressed against the oracle.

Source: unittest/storage/blocksstable/encoding/test_raw_decoder.cpp:774-1200 (and the same
pattern in test_const_decoder.cpp / test_general_column_decoder.cpp for the other codecs): blocks
of ROW_CNT = 64 rows laid out as [seedA x .. | seedB x 10 | .. | NULL x 10], popcounts of the
result bitmap for every white operator, over the whole block and over a 30-row window
(pd_filter_info start/count). Row values come from ObRowGenerate(seed) in the reference; here a
seed maps to an increasing integer / string, which is all the expectations depend on.
"""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
from oceanbase_b200 import White

ROW_CNT = 64


def seed_val(seed, kind):
    if kind == "fix":   # equal-length strings (STRING_DIFF needs them; HEX_PACKING then takes its fixed store)
        return b"seed-%04d-fix" % seed
    return (seed * 1000 + 7) if kind == "int" else b"seed-%04d-%s" % (seed, b"x" * (seed % 5))


def build(layout, kind, enc):
    """layout: list of (seed or None, count)."""
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
    try:
        return ora.Block(ob.encode_block([pad, col]))
    except ob.ObGpuError as e:
        if e.code == ob.OB_NOT_SUPPORTED and enc in (ob.ENC_STRING_DIFF, ob.ENC_HEX_PACKING, ob.ENC_STRING_PREFIX):
            pytest.skip("this layout does not suit the encoder (the reference's traverse() says 'not suitable' too)")
        raise


CASES = [("int", ob.ENC_RAW), ("int", ob.ENC_DICT), ("int", ob.ENC_RLE), ("int", ob.ENC_INTEGER_BASE_DIFF),
         ("str", ob.ENC_RAW), ("str", ob.ENC_DICT), ("str", ob.ENC_RLE),
         # the codecs that rebuild their strings (test_general_column_decoder.cpp: TestHexDecoder / string diff / prefix fixtures)
         ("str", ob.ENC_HEX_PACKING), ("str", ob.ENC_STRING_PREFIX), ("fix", ob.ENC_RAW), ("fix", ob.ENC_HEX_PACKING),
         ("fix", ob.ENC_STRING_DIFF), ("fix", ob.ENC_STRING_PREFIX)]


def pop(blk, op, params, start=0, count=None):
    return int(blk.filter_tree(White(1, op, params), start, count).sum())


@pytest.mark.parametrize("kind,enc", CASES)
def test_filter_pushdown_all_eq_ne(kind, enc):
    # test_raw_decoder.cpp:774-860: [seed1 x N-20 | seed2 x 10 | NULL x 10]
    s1, s2 = 0xF, 0x0
    blk = build([(s1, ROW_CNT - 20), (s2, 10), (None, 10)], kind, enc)
    ref = (seed_val(s1, kind),)
    assert pop(blk, ob.WHITE_OP_EQ, ref) == ROW_CNT - 20
    assert pop(blk, ob.WHITE_OP_EQ, ref, ROW_CNT - 45, 30) == 25
    assert pop(blk, ob.WHITE_OP_NE, ref) == 10
    assert pop(blk, ob.WHITE_OP_NE, ref, ROW_CNT - 45, 30) == 5


@pytest.mark.parametrize("kind,enc", CASES)
def test_filter_push_down_gt_lt_ge_le(kind, enc):
    # :862-980: [seed0 x N-30 | seed1 x 10 | seed2 x 10 | NULL x 10], constant = seed1
    blk = build([(0, ROW_CNT - 30), (1, 10), (2, 10), (None, 10)], kind, enc)
    ref = (seed_val(1, kind),)
    s0, s1, s2 = ROW_CNT - 30, 10, 10
    w = (ROW_CNT - 45, 30)
    assert pop(blk, ob.WHITE_OP_GT, ref) == s2
    assert pop(blk, ob.WHITE_OP_GT, ref, *w) == 5
    assert pop(blk, ob.WHITE_OP_LT, ref) == s0
    assert pop(blk, ob.WHITE_OP_LT, ref, *w) == 15
    assert pop(blk, ob.WHITE_OP_GE, ref) == s1 + s2
    assert pop(blk, ob.WHITE_OP_GE, ref, *w) == s1 + 5
    assert pop(blk, ob.WHITE_OP_LE, ref) == s0 + s1
    assert pop(blk, ob.WHITE_OP_LE, ref, *w) == 15 + s1


@pytest.mark.parametrize("kind,enc", CASES)
def test_filter_push_down_bt(kind, enc):
    # :982-1066: [seed0 x N-10 | seed1 x 10], BETWEEN seed0 AND seed2; reversed bounds match nothing
    blk = build([(0, ROW_CNT - 10), (1, 10)], kind, enc)
    lo, hi = seed_val(0, kind), seed_val(2, kind)
    assert pop(blk, ob.WHITE_OP_BT, (lo, hi)) == ROW_CNT
    assert pop(blk, ob.WHITE_OP_BT, (lo, hi), ROW_CNT - 35, 30) == 30
    assert pop(blk, ob.WHITE_OP_BT, (hi, lo)) == 0
    assert pop(blk, ob.WHITE_OP_BT, (hi, lo), ROW_CNT - 35, 30) == 0


@pytest.mark.parametrize("kind,enc", CASES)
def test_filter_push_down_in_nu(kind, enc):
    # :1068-1198: [seed0 x N-40 | seed1 x 10 | seed2 x 10 | seed3 x 10 | NULL x 10]
    blk = build([(0, ROW_CNT - 40), (1, 10), (2, 10), (3, 10), (None, 10)], kind, enc)
    w = (ROW_CNT - 35, 30)
    in1 = (seed_val(1, kind), seed_val(2, kind), seed_val(5, kind))
    assert pop(blk, ob.WHITE_OP_IN, in1) == 20
    assert pop(blk, ob.WHITE_OP_IN, in1, *w) == 5 + 10
    in2 = (seed_val(5, kind),) * 3
    assert pop(blk, ob.WHITE_OP_IN, in2) == 0
    assert pop(blk, ob.WHITE_OP_IN, in2, *w) == 0
    assert pop(blk, ob.WHITE_OP_NU, ()) == 10
    assert pop(blk, ob.WHITE_OP_NU, (), *w) == 5
    assert pop(blk, ob.WHITE_OP_NN, ()) == ROW_CNT - 10
    assert pop(blk, ob.WHITE_OP_NN, (), *w) == 25


@pytest.mark.parametrize("kind,enc", CASES)
def test_null_constant_matches_nothing(kind, enc):
    # ob_micro_block_decoder.cpp:1713-1715
    blk = build([(0, 40), (None, 24)], kind, enc)
    for op in (ob.WHITE_OP_EQ, ob.WHITE_OP_NE, ob.WHITE_OP_LT, ob.WHITE_OP_GE):
        assert pop(blk, op, (None,)) == 0


def test_and_or_tree_with_early_out():
    # ObPushdownFilterExecutor::execute (ob_pushdown_filter.cpp:1551-1624)
    from oceanbase_b200 import And, Or
    n = 500
    a = np.arange(n, dtype=np.int64)
    b = (np.arange(n, dtype=np.int64) * 7) % 100
    blk = ora.Block(ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, a), ob.Column(ob.OBJ_INT, ob.ENC_DICT, b)]))
    t = And([White(0, ob.WHITE_OP_GE, (100,)), Or([White(1, ob.WHITE_OP_LT, (10,)), White(1, ob.WHITE_OP_EQ, (50,))])])
    got = blk.filter_tree(t)
    exp = ((a >= 100) & ((b < 10) | (b == 50))).astype(np.uint8)
    assert np.array_equal(got, exp)
    none = And([White(0, ob.WHITE_OP_LT, (0,)), White(1, ob.WHITE_OP_GE, (0,))])
    assert blk.filter_tree(none).sum() == 0
    allt = Or([White(0, ob.WHITE_OP_GE, (0,)), White(1, ob.WHITE_OP_LT, (0,))])
    assert blk.filter_tree(allt).sum() == n
