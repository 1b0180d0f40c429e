"""Phase B of the compaction on the device (include/obgpu_compaction.h: obgpu_encode_columns, obgpu_merge_result_encode,
obgpu_column_checksums): the device-built micro-blocks must be byte for byte the blocks the host writer builds for the same
rows (ObMicroBlockEncoder::build_block with ObRawEncoder on every column, encoding/ob_micro_block_encoder.cpp:561-721,
ob_raw_encoder.cpp:95-155), verify under the oracle's checksum checks (ob_micro_block_header.cpp:236-285), decode back to
the input cells, and the column checksums must equal the oracle's (ob_micro_block_checksum_helper.cpp:127-257)."""
import numpy as np
import pytest

import oracle_binding as ora

pytestmark = pytest.mark.gpu


def _ctx():
    import oceanbase_b200 as ob
    return ob.ScanContext(0)


def _datum_len(t):
    from oceanbase_b200 import capi
    return 1 if t == capi.OBJ_YEAR else (4 if t == capi.OBJ_DATE else 8)


def make_columns(rng, n, spec):
    """spec: list of (obj_type, magnitude bits, null fraction, byte_packing_only) -> (values int64, nulls uint8 or None)"""
    from oceanbase_b200 import capi
    store = {capi.OBJ_TINYINT: 1, capi.OBJ_SMALLINT: 2, capi.OBJ_INT32: 4, capi.OBJ_INT: 8, capi.OBJ_UINT64: 8, capi.OBJ_UINT32: 4,
             capi.OBJ_DATE: 4, capi.OBJ_YEAR: 1, capi.OBJ_DATETIME: 8, capi.OBJ_UTINYINT: 1}
    cols = []
    for (t, bits, nullf, bo) in spec:
        bits = min(bits, store[t] * 8)
        signed = t in (capi.OBJ_TINYINT, capi.OBJ_SMALLINT, capi.OBJ_INT32, capi.OBJ_INT, capi.OBJ_DATETIME)
        if bits >= 64:
            v = rng.integers(-2**63, 2**63 - 1, size=n, dtype=np.int64)
        else:
            v = rng.integers(0, 1 << bits, size=n, dtype=np.int64)
            if signed and bits == store[t] * 8 and bits < 64:   # the full width of a narrow signed type: negative values too
                v = v - (1 << (bits - 1))
        nl = (rng.random(n) < nullf).astype(np.uint8) if nullf > 0 else None
        cols.append((t, v, nl, bo))
    return cols


def host_image(cols, rows_per_block, rowkey_cnt, align=128):
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_table
    cs = [Column(t, capi.ENC_RAW, v, nulls=nl, byte_packing_only=bo) for (t, v, nl, bo) in cols]
    return encode_table(cs, rows_per_block, rowkey_cnt=rowkey_cnt, align=align)


def device_encode(ctx, cols, rows_per_block, rowkey_cnt, align=128):
    import torch
    from oceanbase_b200 import compaction
    keep, dcols = [], []
    for (t, v, nl, bo) in cols:
        dv = torch.from_numpy(v).cuda()
        dn = torch.from_numpy(nl).cuda() if nl is not None else None
        keep += [dv, dn]
        dcols.append((dv.data_ptr(), dn.data_ptr() if dn is not None else None, t, bo))
    n = len(cols[0][1])
    return compaction.encode_columns(ctx, dcols, n, rows_per_block, rowkey_cnt=rowkey_cnt, align=align, keep=keep), dcols


def assert_same_image(enc, table):
    img, off, sz = enc.fetch()
    info = enc.info()
    assert info.n_blocks == table.n_blocks
    assert info.n_host_blocks == 0
    assert np.array_equal(off, np.asarray(table.offsets)), "block offsets differ"
    assert np.array_equal(sz, np.asarray(table.sizes)), "block sizes differ"
    want = np.asarray(table.image)
    assert img.size == want.size
    if not np.array_equal(img, want):
        bad = int(np.nonzero(img != want)[0][0])
        blk = int(np.searchsorted(off, bad, side="right") - 1)
        raise AssertionError(f"first differing byte {bad} (block {blk}, byte {bad - off[blk]} of {sz[blk]})")


SHAPES = [
    # cfg5 shape: INT64 rowkey + 3 INT64 payload columns
    ("cfg5", [(5, 40, 0.0, False), (5, 33, 0.0, False), (5, 64, 0.05, False), (5, 13, 0.3, False)], 20_000, 500, 1),
    # every packing rule: bit packing widths, byte widths 1..8, byte_packing_only
    ("widths", [(5, b, 0.0, False) for b in (1, 3, 7, 8, 9, 12, 15, 16, 17, 21, 24, 31, 32, 33, 40, 47, 48, 55, 56, 57, 63, 64)], 5_000, 333, 0),
    ("byte_only", [(5, b, 0.1, True) for b in (1, 7, 9, 17, 33, 64)], 4_000, 257, 0),
    # narrow types, signed full range (the store image is masked to the type's store size), NULLs in some columns only
    ("types", [(1, 8, 0.1, False), (2, 16, 0.0, False), (4, 32, 0.2, False), (10, 64, 0.0, False), (9, 32, 0.0, False), (19, 20, 0.1, False),
               (21, 8, 0.0, False), (17, 50, 0.2, False), (6, 5, 0.0, False)], 7_000, 1_000, 2),
    # one row per block, and a last short block
    ("tiny", [(5, 20, 0.5, False), (5, 3, 0.0, False)], 37, 1, 1),
    ("ragged", [(5, 20, 0.2, False), (5, 64, 0.0, False)], 1_001, 100, 1),
]


@pytest.mark.parametrize("name,spec,n,rpb,rk", SHAPES, ids=[s[0] for s in SHAPES])
def test_device_blocks_equal_the_host_writer(name, spec, n, rpb, rk):
    ctx = _ctx()
    rng = np.random.default_rng(abs(hash(name)) % 1000 + 7)
    cols = make_columns(rng, n, spec)
    table = host_image(cols, rpb, rk)
    enc, _ = device_encode(ctx, cols, rpb, rk)
    assert_same_image(enc, table)
    # the oracle accepts the device's blocks (header + payload checksums) and decodes the input back
    img, off, sz = enc.fetch()
    for b in (0, len(off) // 2, len(off) - 1):
        blk = ora.Block(img[off[b]:off[b] + sz[b]].copy())
        assert blk.verify_checksums() == 0
    # column checksums
    o = ora.oracle()
    got = enc.column_checksums()
    for c, (t, v, nl, _bo) in enumerate(cols):
        want = o.ora_column_checksum(v.ctypes.data, nl.ctypes.data if nl is not None else None, n, _datum_len(t))
        assert int(got[c]) == want, (c, t)
    enc.free()


def test_alignments_and_big_blocks():
    ctx = _ctx()
    rng = np.random.default_rng(5)
    cols = make_columns(rng, 30_000, [(5, 45, 0.0, False), (5, 64, 0.1, False), (5, 10, 0.0, False), (5, 27, 0.02, False)])
    for align, rpb in ((16, 700), (512, 1400), (4096, 2048)):
        table = host_image(cols, rpb, 1, align=align)
        enc, _ = device_encode(ctx, cols, rpb, 1, align=align)
        assert_same_image(enc, table)
        enc.free()


def test_null_dominated_column_is_left_to_the_host():
    """ObRawEncoder::traverse stores a NULL-dominated column as var-length cells (ob_raw_encoder.cpp:106-110): such a block
    reads size 0 and the blocks around it keep their bytes."""
    ctx = _ctx()
    rng = np.random.default_rng(6)
    n, rpb = 4_000, 500
    cols = make_columns(rng, n, [(5, 30, 0.0, False), (5, 60, 0.0, False)])
    nl = np.zeros(n, dtype=np.uint8)
    nl[1000:1500] = 1          # block 2: every cell of column 1 NULL -> max 0 -> 1-bit packing: 1 * 500 <= 500 * 16, stays fixed
    nl[1500:2000:2] = 1        # block 3: half of the 60-bit cells NULL: 60 * 250 > 500 * 16 -> var-stored in the reference
    cols[1] = (cols[1][0], cols[1][1], nl, False)
    enc, _ = device_encode(ctx, cols, rpb, 1)
    img, off, sz = enc.fetch()
    assert enc.info().n_host_blocks == 1
    assert sz[3] == 0
    table = host_image(cols, rpb, 1)
    assert np.asarray(table.block(3))[22] == 1 and sz[2] == table.sizes[2]   # opt2_: the host writer stored one var column there
    pos = 0
    for b in range(8):
        assert off[b] == pos
        if b == 3:
            continue
        want = np.asarray(table.block(b))
        assert np.array_equal(img[off[b]:off[b] + sz[b]], want), b
        pos += (sz[b] + 127) // 128 * 128
    enc.free()


def test_column_checksums_alone():
    import torch
    from oceanbase_b200 import compaction, capi
    ctx = _ctx()
    rng = np.random.default_rng(8)
    n = 1_000_003
    v = rng.integers(-2**63, 2**63 - 1, size=n, dtype=np.int64)
    nl = (rng.random(n) < 0.1).astype(np.uint8)
    d = (rng.integers(0, 1 << 20, size=n, dtype=np.int64))
    dv, dn, dd = torch.from_numpy(v).cuda(), torch.from_numpy(nl).cuda(), torch.from_numpy(d).cuda()
    got = compaction.column_checksums(ctx, [(dv.data_ptr(), dn.data_ptr(), capi.OBJ_INT, False), (dd.data_ptr(), None, capi.OBJ_DATE, False)], n)
    o = ora.oracle()
    assert int(got[0]) == o.ora_column_checksum(v.ctypes.data, nl.ctypes.data, n, 8)
    assert int(got[1]) == o.ora_column_checksum(d.ctypes.data, None, n, 4)


def test_column_groups_of_a_merge_result():
    """Column-oriented merge, writer side: the merged stream replayed into one writer per column group
    (ObCOMergeLogReplayer / ObWriteHelper::project, column_store/ob_co_merge_writer.cpp:67-117). Every group's blocks equal
    the host writer's over the merged rows the CPU oracle merge produces, and re-open as page batches that scan."""
    import oceanbase_b200 as ob
    from oceanbase_b200 import capi, compaction
    from oceanbase_b200.sstable import Column, encode_table
    ctx = _ctx()
    rng = np.random.default_rng(9)
    runs = []
    for r in range(3):
        n = 6_000
        key = np.sort(rng.choice(40_000, size=n, replace=False)).astype(np.int64)
        cols = [Column(capi.OBJ_INT, capi.ENC_INTEGER_BASE_DIFF, key)]
        for c in range(3):
            v = rng.integers(0, 1 << (10 + 17 * c), size=n, dtype=np.int64)
            nl = (rng.random(n) < 0.1).astype(np.uint8) if c == 1 else None
            cols.append(Column(capi.OBJ_INT, capi.ENC_RAW, v, nulls=nl))
        runs.append(encode_table(cols, 700, rowkey_cnt=1))
    batches = [ob.PageBatch(ctx, t) for t in runs]
    res = compaction.merge_batches(ctx, batches, 0, None, [1, 2, 3])
    n_out = res.info().out_rows
    key, _ = res.fetch(-1)
    payload = [res.fetch(c) for c in range(3)]
    groups = [[-1, 0, 1, 2], [0], [1], [2, 0]]   # all-column group with the rowkey first, then pure column groups
    types = {-1: capi.OBJ_INT, 0: capi.OBJ_INT, 1: capi.OBJ_INT, 2: capi.OBJ_INT}
    encs = compaction.co_merge_write(res, groups, types, rows_per_block=900)
    for cg, enc in zip(groups, encs):
        host_cols = []
        for c in cg:
            if c == -1:
                host_cols.append(Column(capi.OBJ_INT, capi.ENC_RAW, key))
            else:
                v, nl = payload[c]
                host_cols.append(Column(capi.OBJ_INT, capi.ENC_RAW, v, nulls=nl if nl.any() else None))
        table = encode_table(host_cols, 900, rowkey_cnt=1 if cg[0] == -1 else 0)
        assert_same_image(enc, table)
        assert enc.info().total_rows == n_out
        enc.free()
