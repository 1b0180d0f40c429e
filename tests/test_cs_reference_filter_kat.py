"""CS (column-store encoding) white filters pinned to the reference's own unit-test expectations: the datasets and the
expected selected-row counts of
  unittest/storage/blocksstable/cs_encoding/test_integer_pd_filter.cpp:23-112   test_integer_decoder_filter
                                            test_integer_pd_filter.cpp:114-189  test_integer_decoder_uint_type
                                            test_int_dict_pd_filter.cpp:186-270 test_int_dict_const_decoder (const-encoded refs)
                                            test_string_pd_filter.cpp:25-123    test_string_decoder_filter_varchar
                                            test_str_dict_pd_filter.cpp:276-372 test_var_string_dict_const_decoder (const-encoded refs)
are rebuilt with this repo's CS writer and evaluated by the oracle (and, in tests/test_gpu_cs.py, by the device)."""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
from oceanbase_b200 import White

OPS = {"EQ": ob.WHITE_OP_EQ, "NE": ob.WHITE_OP_NE, "LT": ob.WHITE_OP_LT, "LE": ob.WHITE_OP_LE, "GT": ob.WHITE_OP_GT,
       "GE": ob.WHITE_OP_GE, "IN": ob.WHITE_OP_IN, "BT": ob.WHITE_OP_BT, "NU": ob.WHITE_OP_NU, "NN": ob.WHITE_OP_NN}


def integer_dataset(has_null):
    n = 120 if has_null else 100
    v = np.array([i - 50 if i < 100 else 0 for i in range(n)], dtype=np.int64)
    nulls = np.array([0 if i < 100 else 1 for i in range(n)], dtype=np.uint8) if has_null else None
    cases = [("NU", (), 20 if has_null else 0), ("NN", (), 100)]
    cases += [("EQ", (r,), e) for r, e in zip((-55, -50, 40, 55), (0, 1, 1, 0))]
    cases += [("NE", (r,), e) for r, e in zip((-55, -50, 40, 55), (100, 99, 99, 100))]
    cases += [("LE", (r,), e) for r, e in zip((-55, -50, -40, 55), (0, 1, 11, 100))]
    cases += [("LT", (r,), e) for r, e in zip((-55, -50, -40, 55), (0, 0, 10, 100))]
    cases += [("GE", (r,), e) for r, e in zip((-55, 40, 49, 55), (100, 10, 1, 0))]
    cases += [("GT", (r,), e) for r, e in zip((-55, 40, 49, 55), (100, 9, 0, 0))]
    cases += [("IN", (-55, -27, 0, 10, 100), 3)]
    cases += [("BT", p, e) for p, e in zip(((-100, -90), (-55, -47), (-4, 4), (47, 55), (90, 100)), (0, 4, 9, 3, 0))]
    return ob.OBJ_INT, v, nulls, cases


def uint_dataset(has_null):
    n = 120 if has_null else 100
    v = np.array([100 + i if i < 100 else 0 for i in range(n)], dtype=np.int64)
    nulls = np.array([0 if i < 100 else 1 for i in range(n)], dtype=np.uint8) if has_null else None
    refs = (100, 199, 219, (1 << 32) - 1)
    cases = [("NU", (), 20 if has_null else 0), ("NN", (), 100)]
    for op, exp in (("EQ", (1, 1, 0, 0)), ("NE", (99, 99, 100, 100)), ("GT", (99, 0, 0, 0)), ("GE", (100, 1, 0, 0)),
                    ("LT", (0, 99, 100, 100)), ("LE", (1, 100, 100, 100))):
        cases += [(op, (r,), e) for r, e in zip(refs, exp)]
    return ob.OBJ_UINT32, v, nulls, cases


def const_dict_dataset():
    n = 120
    v = np.array([30] * 115 + [100, 200, 300, 400, 0], dtype=np.int64)
    nulls = np.array([0] * 119 + [1], dtype=np.uint8)
    cases = [("NU", (), 1), ("NN", (), 119)]
    cases += [("EQ", (r,), e) for r, e in zip((-100, 30, 100, 101), (0, 115, 1, 0))]
    cases += [("NE", (r,), e) for r, e in zip((-100, 30, 100, 101), (119, 4, 118, 119))]
    cases += [("LT", (r,), e) for r, e in zip((-100, 30, 31, 300), (0, 0, 115, 117))]
    cases += [("LE", (r,), e) for r, e in zip((-100, 30, 31, 300), (0, 115, 115, 118))]
    cases += [("GT", (r,), e) for r, e in zip((30, 100, 400, 500), (4, 3, 0, 0))]
    cases += [("GE", (r,), e) for r, e in zip((30, 100, 400, 500), (119, 4, 1, 0))]
    cases += [("IN", (-100, 30, 105, 300), 116)]
    cases += [("BT", p, e) for p, e in zip(((-100, -50), (-10, 40), (-1, 100), (31, 105), (50, 500)), (0, 115, 116, 1, 4))]
    return ob.OBJ_INT, v, nulls, cases


def string_dataset(has_null):
    n = 120 if has_null else 100
    s = lambda idx, ln: bytes([ord("a") + idx]) * ln
    v = [s(i // 25, i % 25 + 1) if i < 100 else b"" for i in range(n)]
    nulls = np.array([0 if i < 100 else 1 for i in range(n)], dtype=np.uint8) if has_null else None
    cases = [("NU", (), 20 if has_null else 0), ("NN", (), 100)]
    cases += [("EQ", (s(*r),), e) for r, e in zip(((0, 1), (0, 10), (1, 2), (1, 100), (3, 10)), (1, 1, 1, 0, 1))]
    cases += [("NE", (s(*r),), e) for r, e in zip(((0, 1), (0, 10), (1, 2), (1, 100), (3, 10)), (99, 99, 99, 100, 99))]
    cases += [("LT", (s(*r),), e) for r, e in zip(((0, 1), (0, 10), (1, 2), (1, 10), (3, 10)), (0, 9, 26, 34, 84))]
    cases += [("LE", (s(*r),), e) for r, e in zip(((0, 1), (0, 10), (1, 2), (1, 10), (3, 10)), (1, 10, 27, 35, 85))]
    cases += [("GT", (s(*r),), e) for r, e in zip(((3, 25), (2, 25), (1, 10)), (0, 25, 65))]
    cases += [("GE", (s(*r),), e) for r, e in zip(((3, 25), (2, 25), (1, 10)), (1, 26, 66))]
    cases += [("IN", tuple(s(*r) for r in ((0, 5), (1, 1), (1, 40), (2, 100), (3, 20))), 3)]
    cases += [("BT", (s(*a), s(*b)), e) for (a, b), e in zip((((0, 1), (0, 10)), ((0, 1), (1, 10)), ((1, 10), (3, 1)), ((3, 10), (4, 20))),
                                                             (10, 35, 42, 16))]
    return ob.OBJ_VARCHAR, v, nulls, cases


def str_dict_const_dataset():
    # test_str_dict_pd_filter.cpp:276-372 test_var_string_dict_const_decoder: one dominant string, 4 exceptions, one NULL
    s = lambda idx, ln: bytes([ord("a") + idx]) * ln
    v = [s(0, 50)] * 115 + [s(1, 50), s(2, 50), s(3, 50), s(4, 50), b""]
    nulls = np.array([0] * 119 + [1], dtype=np.uint8)
    cases = [("NU", (), 1), ("NN", (), 119)]
    cases += [("EQ", (s(*r),), e) for r, e in zip(((0, 50), (1, 50), (0, 100)), (115, 1, 0))]
    cases += [("NE", (s(*r),), e) for r, e in zip(((0, 50), (1, 50), (0, 100)), (4, 118, 119))]
    cases += [("LT", (s(*r),), e) for r, e in zip(((0, 50), (0, 51), (2, 50)), (0, 115, 116))]
    cases += [("LE", (s(*r),), e) for r, e in zip(((0, 50), (0, 51), (2, 50)), (115, 115, 117))]
    cases += [("GT", (s(*r),), e) for r, e in zip(((0, 49), (0, 50), (2, 50)), (119, 4, 2))]
    cases += [("GE", (s(*r),), e) for r, e in zip(((0, 49), (0, 50), (2, 50)), (119, 119, 3))]
    cases += [("IN", tuple(s(*r) for r in ((0, 50), (0, 100), (1, 40), (2, 100), (3, 50))), 116)]
    cases += [("BT", (s(*a), s(*b)), e) for (a, b), e in zip((((0, 10), (0, 49)), ((0, 10), (1, 50)), ((1, 50), (4, 50))), (0, 116, 4))]
    return ob.OBJ_VARCHAR, v, nulls, cases


DATASETS = {
    "integer": (lambda: integer_dataset(False), [ob.ENC_CS_INTEGER, ob.ENC_CS_INT_DICT, ob.ENC_RAW, ob.ENC_DICT]),
    "integer_nulls": (lambda: integer_dataset(True), [ob.ENC_CS_INTEGER, ob.ENC_CS_INT_DICT, ob.ENC_RAW, ob.ENC_DICT]),
    "uint": (lambda: uint_dataset(False), [ob.ENC_CS_INTEGER, ob.ENC_CS_INT_DICT]),
    "uint_nulls": (lambda: uint_dataset(True), [ob.ENC_CS_INTEGER, ob.ENC_CS_INT_DICT]),
    "int_dict_const": (const_dict_dataset, [ob.ENC_CS_INT_DICT, ob.ENC_CS_INTEGER]),
    "str_dict_const": (str_dict_const_dataset, [ob.ENC_CS_STR_DICT, ob.ENC_CS_STRING, ob.ENC_DICT]),
    "varchar": (lambda: string_dataset(False), [ob.ENC_CS_STRING, ob.ENC_CS_STR_DICT, ob.ENC_RAW, ob.ENC_DICT]),
    "varchar_nulls": (lambda: string_dataset(True), [ob.ENC_CS_STRING, ob.ENC_CS_STR_DICT, ob.ENC_RAW, ob.ENC_DICT]),
}


def build(name, enc):
    obj_type, v, nulls, cases = DATASETS[name][0]()
    n = len(v)
    first = ob.ENC_CS_INTEGER if enc >= ob.ENC_CS_INTEGER else ob.ENC_RAW
    block = ob.encode_block([ob.Column(ob.OBJ_INT32, first, np.arange(n, dtype=np.int64)), ob.Column(obj_type, int(enc), v, nulls=nulls)])
    return block, cases


@pytest.mark.parametrize("name", sorted(DATASETS))
def test_reference_cs_filter_expectations(name):
    for enc in DATASETS[name][1]:
        block, cases = build(name, enc)
        blk = ora.Block(block)
        assert blk.verify_checksums() == 0
        for op, params, expect in cases:
            got = int(blk.filter_tree(White(1, OPS[op], params)).sum())
            assert got == expect, (name, enc, op, params)
    if name == "int_dict_const":      # the shape the reference stores with const-encoded refs
        block, _ = build(name, ob.ENC_CS_INT_DICT)
        hs = 64
        ah = block[hs:hs + 12]
        so = block[len(block) - int(ah[6:10].view(np.uint32)[0]):]
        ends = so[5:].view({1: np.uint8, 2: np.uint16, 4: np.uint32}[1 << int(so[3])])
        dm = block[int(ends[0]):int(ends[0]) + 10]
        assert dm[1] & 0x4 and int(dm[6:10].view(np.uint32)[0]) == 2 + 2 * 5
