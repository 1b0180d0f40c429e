"""Edge cases of the fused scan: NULLs in every codec, narrow / unsigned / date types, var-length
RAW strings, empty and full selections, ragged tail blocks, 1-row and max-row blocks, no
projection, unsupported encodings and corrupt headers (error behaviour of the reference:
OB_NOT_SUPPORTED => caller falls back, OB_INVALID_DATA)."""
import numpy as np
import pytest

import oracle_binding as ora
from test_gpu_scan import assert_scan_matches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import oceanbase_b200
    return oceanbase_b200


@pytest.fixture(scope="module")
def ctx(ob):
    c = ob.ScanContext(0)
    yield c
    c.close()


class W:
    def __init__(self, table, flt, proj, is_str, elem):
        self.table, self.filter, self.proj, self.proj_is_string, self.proj_elem_len = table, flt, proj, is_str, elem


def _mixed_table(ob, n, rows_per_block, seed, null_frac=0.12):
    rng = np.random.default_rng(seed)
    nl = lambda: (rng.random(n) < null_frac).astype(np.uint8)
    strs = [bytes(rng.integers(97, 123, size=rng.integers(0, 20), dtype=np.uint8)) for _ in range(97)]
    pick = lambda: [strs[i] for i in rng.integers(0, 97, size=n)]
    runs = lambda x: [x[i // 6] for i in range(n)]
    cols = [
        ob.Column(ob.OBJ_INT, ob.ENC_RAW, rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64), nulls=nl()),
        ob.Column(ob.OBJ_INT, ob.ENC_RAW, rng.integers(0, 1 << 11, size=n, dtype=np.int64), nulls=nl()),
        ob.Column(ob.OBJ_INT, ob.ENC_DICT, rng.integers(-50, 50, size=n, dtype=np.int64) * 10 ** 9, nulls=nl()),
        ob.Column(ob.OBJ_INT, ob.ENC_RLE, np.repeat(rng.integers(0, 9, size=n // 5 + 1), 5)[:n].astype(np.int64), nulls=nl()),
        ob.Column(ob.OBJ_INT, ob.ENC_INTEGER_BASE_DIFF, rng.integers(10 ** 15, 10 ** 15 + 5000, size=n, dtype=np.int64), nulls=nl()),
        ob.Column(ob.OBJ_INT32, ob.ENC_RAW, rng.integers(-(1 << 31), 1 << 31, size=n, dtype=np.int64)),
        ob.Column(ob.OBJ_DATE, ob.ENC_DICT, rng.integers(-20000, 20000, size=n, dtype=np.int64), nulls=nl()),
        ob.Column(ob.OBJ_UINT64, ob.ENC_RAW, rng.integers(0, 1 << 63, size=n, dtype=np.int64) * 2 + 1),
        ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, pick(), nulls=nl()),
        ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, pick(), nulls=nl()),
        ob.Column(ob.OBJ_VARCHAR, ob.ENC_RLE, runs(pick()), nulls=nl()),
        ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, pick()),
        ob.Column(ob.OBJ_TINYINT, ob.ENC_INTEGER_BASE_DIFF, rng.integers(-128, 128, size=n, dtype=np.int64)),
    ]
    return ob.encode_table(cols, rows_per_block), cols


PROJ = list(range(13))
IS_STR = [False] * 8 + [True] * 4 + [False]
ELEM = [8, 8, 8, 8, 8, 8, 4, 8, 8, 8, 8, 8, 8]


def test_nulls_everywhere_all_types(ob, ctx):
    table, _ = _mixed_table(ob, 20_000, 700, 1)
    flt = ob.And([ob.White(1, ob.WHITE_OP_LT, (1500,)), ob.Or([ob.White(9, ob.WHITE_OP_NN, ()), ob.White(6, ob.WHITE_OP_NU, ())])])
    assert_scan_matches(ctx, W(table, flt, PROJ, IS_STR, ELEM))


@pytest.mark.parametrize("flt_id", range(8))
def test_filters_on_each_codec(ob, ctx, flt_id):
    table, _ = _mixed_table(ob, 9_000, 512, 2)
    flts = [
        ob.White(0, ob.WHITE_OP_GT, (0,)),
        ob.White(2, ob.WHITE_OP_IN, (10 ** 9, -3 * 10 ** 9, 49 * 10 ** 9, 12345)),
        ob.White(3, ob.WHITE_OP_NE, (4,)),
        ob.White(4, ob.WHITE_OP_BT, (10 ** 15 + 100, 10 ** 15 + 2500)),
        ob.White(7, ob.WHITE_OP_GE, ((1 << 63) + 12345,)),      # unsigned compare
        ob.White(8, ob.WHITE_OP_LE, (b"mmmm",)),                # RAW var-length string compare
        ob.Or([ob.White(10, ob.WHITE_OP_EQ, (b"",)), ob.White(11, ob.WHITE_OP_GT, (b"x",)), ob.White(5, ob.WHITE_OP_LT, (-(1 << 30),))]),
        ob.And([ob.White(6, ob.WHITE_OP_LT, (0,)), ob.White(12, ob.WHITE_OP_GE, (-5,))]),
    ]
    assert_scan_matches(ctx, W(table, flts[flt_id], PROJ, IS_STR, ELEM))


def test_empty_and_full_selection_and_no_projection(ob, ctx):
    table, _ = _mixed_table(ob, 5_000, 333, 3, null_frac=0.0)
    none = ob.White(1, ob.WHITE_OP_LT, (0,))
    assert assert_scan_matches(ctx, W(table, none, PROJ, IS_STR, ELEM)) == 0
    allr = ob.White(1, ob.WHITE_OP_GE, (0,))
    assert assert_scan_matches(ctx, W(table, allr, PROJ, IS_STR, ELEM)) == 5_000
    assert assert_scan_matches(ctx, W(table, None, [], [], []), want_row_ids=False) == 5_000
    # filter only: count + bitmap, nothing projected
    batch = ctx.open_batch(table)
    res = batch.scan(ob.White(1, ob.WHITE_OP_LT, (1024,)), [])
    want = ora.scan_table(table, ob.White(1, ob.WHITE_OP_LT, (1024,)), [], [], [])
    assert res.selected_rows == want["selected"]
    assert np.array_equal(res.fetch_sel_offsets(), want["sel_offset"])
    res.free()
    batch.close()


@pytest.mark.parametrize("rows,rpb", [(1, 1), (33, 32), (40000, 40000), (70001, 35000), (257, 1), (65535, 65535), (131070, 65535)])
def test_block_shapes(ob, ctx, rows, rpb):
    rng = np.random.default_rng(rows)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, rng.integers(0, 1 << 9, size=rows, dtype=np.int64)),
            ob.Column(ob.OBJ_INT, ob.ENC_DICT, rng.integers(0, 7, size=rows, dtype=np.int64))]
    table = ob.encode_table(cols, rpb)
    flt = ob.White(0, ob.WHITE_OP_LT, (200,))
    assert_scan_matches(ctx, W(table, flt, [0, 1], [False, False], [8, 8]))


def test_unsupported_encoding_reports_not_supported(ob, ctx):
    v = np.arange(100, dtype=np.int64)
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, v), ob.Column(ob.OBJ_INT, ob.ENC_RAW, v)])
    block = block.copy()
    hs = ora.Block(block).b.header_size
    block[hs + 16 + 1] = 5  # column 1 pretends to be STRING_DIFF: not handled by the device path
    table = ob.TableImage(np.concatenate([block, np.zeros(256, dtype=np.uint8)]), np.array([0], dtype=np.int64),
                          np.array([len(block)], dtype=np.int64), 100, 2)
    batch = ctx.open_batch(table)
    ok = batch.scan(None, [0])
    assert ok.selected_rows == 100          # untouched columns do not matter
    ok.free()
    res = batch.scan(None, [1])
    with pytest.raises(ob.ObGpuError) as ei:
        res.info()
    assert ei.value.code == ob.OB_NOT_SUPPORTED
    res.free()
    with pytest.raises(ob.ObGpuError) as ei:
        batch.filter_white(0, 1, ob.WHITE_OP_EQ, (5,))
    assert ei.value.code == ob.OB_NOT_SUPPORTED
    batch.close()


def test_corrupt_header_is_rejected_at_open(ob, ctx):
    v = np.arange(100, dtype=np.int64)
    block = ob.encode_block([ob.Column(ob.OBJ_INT, ob.ENC_RAW, v)]).copy()
    bad = block.copy()
    bad[0] ^= 0xFF  # magic
    for img in (bad,):
        table = ob.TableImage(np.concatenate([img, np.zeros(256, dtype=np.uint8)]), np.array([0], dtype=np.int64),
                              np.array([len(img)], dtype=np.int64), 100, 1)
        with pytest.raises(ob.ObGpuError) as ei:
            ctx.open_batch(table)
        assert ei.value.code == ob.OB_INVALID_DATA
    table = ob.TableImage(np.concatenate([block, np.zeros(256, dtype=np.uint8)]), np.array([8], dtype=np.int64),
                          np.array([len(block)], dtype=np.int64), 100, 1)
    with pytest.raises(ob.ObGpuError) as ei:
        ctx.open_batch(table)  # misaligned offset
    assert ei.value.code == ob.OB_INVALID_ARGUMENT


def test_many_small_blocks_lookback_chain(ob, ctx):
    # 20k tiny blocks: exercises the decoupled look-back across many waves
    rows = 200_000
    rng = np.random.default_rng(77)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, rng.integers(0, 100, size=rows, dtype=np.int64)),
            ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(rows, dtype=np.int64))]
    table = ob.encode_table(cols, 10)
    flt = ob.White(0, ob.WHITE_OP_LT, (37,))
    assert_scan_matches(ctx, W(table, flt, [1], [False], [8]))


@pytest.mark.parametrize("compact", ["0", "1"])
@pytest.mark.parametrize("proj", [[0, 3, 8, 11], [8, 9, 10, 11], [12], list(range(13))])
def test_projection_staging_modes(ob, ctx, compact, proj, monkeypatch):
    # whole-block TMA vs packed per-column regions (several RAW var-length columns share the row data)
    monkeypatch.setenv("OBGPU_PROJECT_COMPACT", compact)
    table, _ = _mixed_table(ob, 12_000, 640, 5)
    flt = ob.White(1, ob.WHITE_OP_LT, (700,))
    assert_scan_matches(ctx, W(table, flt, proj, [IS_STR[i] for i in proj], [ELEM[i] for i in proj]))


@pytest.mark.parametrize("limit", [0, 3, 60, 200])
@pytest.mark.parametrize("split", ["0", "1"])
def test_sparse_selection_decodes_from_global_memory(ob, ctx, limit, split, monkeypatch):
    # <= 1/16 of a block selected: the projection skips the shared-memory staging (every codec, NULLs, strings);
    # split = 1 routes those blocks to the warp-per-block kernel (what a low selectivity hint does)
    monkeypatch.setenv("OBGPU_SPARSE_SPLIT", split)
    table, _ = _mixed_table(ob, 15_000, 750, 9)
    flt = ob.White(1, ob.WHITE_OP_LT, (limit,))
    assert_scan_matches(ctx, W(table, flt, PROJ, IS_STR, ELEM))


def test_blocks_larger_than_shared_memory(ob, ctx):
    # ~1 MB micro blocks (the reference allows up to 1 MB): nothing is staged, every column is decoded from
    # global memory; filters on bit-packed, dictionary and string columns, NULLs, all rows / some rows / few rows
    table, _ = _mixed_table(ob, 60_000, 30_000, 13)
    assert int(table.sizes.max()) > 400_000
    for flt in (None, ob.White(1, ob.WHITE_OP_LT, (900,)), ob.White(1, ob.WHITE_OP_LT, (20,)),
                ob.And([ob.White(2, ob.WHITE_OP_GE, (0,)), ob.White(9, ob.WHITE_OP_NN, ())])):
        assert_scan_matches(ctx, W(table, flt, PROJ, IS_STR, ELEM))


@pytest.mark.parametrize("obj_type,lo,hi,elem", [("OBJ_DATE", -30000, 30000, 4), ("OBJ_UINT32", 1065958239, 1065960342, 8),
                                                  ("OBJ_INT", -5, 1 << 40, 8), ("OBJ_INT32", -100, 100, 8)])
def test_var_stored_integer_columns(ob, ctx, obj_type, lo, hi, elem):
    # NULL-dominated RAW integer columns live in the row data next to the var-length strings (ObRawEncoder::traverse)
    t = getattr(ob, obj_type)
    rng = np.random.default_rng(14)
    n = 9000
    v = rng.integers(lo, hi, size=n, dtype=np.int64)
    nulls = (rng.random(n) < 0.7).astype(np.uint8)
    s = [bytes(rng.integers(97, 123, size=int(rng.integers(0, 9)), dtype=np.uint8)) for _ in range(n)]
    k = rng.integers(0, 1000, size=n, dtype=np.int64)
    table = ob.encode_table([ob.Column(t, ob.ENC_RAW, v, nulls=nulls), ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, s),
                             ob.Column(t, ob.ENC_RAW, v[::-1].copy(), nulls=nulls), ob.Column(ob.OBJ_INT, ob.ENC_RAW, k)], 800)
    assert not (table.block(0)[64 + 2] & 1)                                     # column 0 is var-stored
    mid = int(np.median(v))
    for flt in (None, ob.White(0, ob.WHITE_OP_GE, (mid,)), ob.White(2, ob.WHITE_OP_NU, ()), ob.White(3, ob.WHITE_OP_LT, (30,)),
                ob.And([ob.White(0, ob.WHITE_OP_LT, (mid,)), ob.White(2, ob.WHITE_OP_NE, (int(v[5]),)), ob.White(3, ob.WHITE_OP_GE, (100,))]),
                ob.Or([ob.White(0, ob.WHITE_OP_IN, (int(v[0]), int(v[1]), int(v[2]))), ob.White(1, ob.WHITE_OP_EQ, (s[9],))])):
        assert_scan_matches(ctx, W(table, flt, [0, 1, 2, 3], [False, True, False, False], [elem, 8, elem, 8]))


def test_device_image_without_host_view(ob, ctx):
    """A device-resident image opened with header_view == NULL is surveyed on the device: same batch facts, same scan
    result as the host-validated open; a corrupt header is still refused with OB_INVALID_DATA."""
    import torch
    table, _ = _mixed_table(ob, 5000, 700, seed=41)
    d_img = torch.zeros(table.image.size + 64, dtype=torch.uint8, device="cuda")
    d_img[:table.image.size].copy_(torch.from_numpy(table.image))
    torch.cuda.synchronize()
    flt = ob.And([ob.White(1, ob.WHITE_OP_LT, (1500,)), ob.White(9, ob.WHITE_OP_NN, ())])
    a = ctx.open_batch(table, device_image_ptr=d_img.data_ptr())
    b = ctx.open_batch(table, device_image_ptr=d_img.data_ptr(), host_view=False)
    assert a.total_rows == b.total_rows == 5000
    for i in (0, table.n_blocks - 1):
        assert a.block_info(i) == b.block_info(i)
    ra, rb = a.scan(flt, PROJ, want_row_ids=True), b.scan(flt, PROJ, want_row_ids=True)
    assert ra.selected_rows == rb.selected_rows > 0
    assert np.array_equal(ra.fetch_row_ids(), rb.fetch_row_ids())
    for c in range(len(PROJ)):
        da, la, na = ra.fetch_col(c)
        db, lb, nb = rb.fetch_col(c)
        assert np.array_equal(na, nb)
        if IS_STR[c]:
            assert np.array_equal(la, lb)
        else:
            assert np.array_equal(da, db)
    ra.free(); rb.free(); a.close(); b.close()
    bad = d_img.clone()
    bad[int(table.offsets[1])] ^= 0x55     # magic of block 1
    torch.cuda.synchronize()
    with pytest.raises(ob.ObGpuError) as ei:
        ctx.open_batch(table, device_image_ptr=bad.data_ptr(), host_view=False)
    assert ei.value.code == ob.OB_INVALID_DATA
