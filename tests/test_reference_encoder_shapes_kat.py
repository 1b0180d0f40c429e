"""Shapes taken from the reference's own encoder unit tests (unittest/storage/blocksstable/encoding/test_micro_block_encoder.cpp),
replayed on this repo's writer and oracle (the device reads the same codecs in tests/test_gpu_string_codecs.py):
  * test_string_diff_null_length (:358-418): RAW rowkey + a STRING_DIFF column holding three 81-byte strings that share a prefix and a
    suffix, followed by 52 400 NULL rows -- NULLs dominate, so the codec switches to var-stored cells with a null length; every row must
    read back (three strings, then NULLs);
  * test_dict_large_varchar (:191-249 shape): DICT over a few very long strings."""
import numpy as np

import oracle_binding as ora

STR1 = b"this is a meaningless string for a test case to verify fixed string diff encoding"
STR2 = b"this is 7777777777777777777777777777777777777777777777777777777777777777 encoding"
STR3 = b"this is 8888888888888888888888888888888888888888888888888888888888888888 encoding"
N_NULL = 52_400


def string_diff_block():
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_block
    n = 3 + N_NULL
    strs = [STR1, STR2, STR3] + [b""] * N_NULL
    nulls = np.ones(n, dtype=np.uint8)
    nulls[:3] = 0
    return encode_block([Column(capi.OBJ_INT, capi.ENC_RAW, np.arange(n, dtype=np.int64)),
                         Column(capi.OBJ_VARCHAR, capi.ENC_STRING_DIFF, strs, nulls=nulls)]), n


def test_string_diff_null_length_shape_round_trips():
    assert len(STR1) == len(STR2) == len(STR3) == 81
    block, n = string_diff_block()
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0
    assert blk.row_count == n
    assert int(block[64 + 16 * 1 + 1]) == 5          # the column really is STRING_DIFF
    assert blk.cell(1, 0) == STR1 and blk.cell(1, 1) == STR2 and blk.cell(1, 2) == STR3
    for r in list(range(3, n, 997)) + [n - 1]:
        assert blk.cell(1, r) is None, r
        assert blk.cell(0, r) == r
    # a NULL-dominated column costs (almost) nothing per NULL row: three cells + a row index, not 52 403 x 81 bytes
    assert block.size < 3 * 81 + n * 12


def test_dict_large_varchar_shape_round_trips():
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_block
    rng = np.random.default_rng(5)
    big = [bytes(rng.integers(32, 127, size=ln, dtype=np.uint8)) for ln in (70_000, 66_000, 65_536, 9)]
    strs = [big[i % 4] for i in range(40)]
    block = encode_block([Column(capi.OBJ_INT, capi.ENC_RAW, np.arange(40, dtype=np.int64)), Column(capi.OBJ_VARCHAR, capi.ENC_DICT, strs)])
    blk = ora.Block(block)
    assert blk.verify_checksums() == 0
    assert blk.dict_count(1) == 4
    for r in range(40):
        assert blk.cell(1, r) == strs[r], r
