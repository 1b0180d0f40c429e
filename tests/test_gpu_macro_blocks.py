"""Macro blocks (disk format) opened on the device (obgpu_batch_open_macro_blocks): the page batch built from the macro image must
scan exactly like the page batch built from the plain micro-block image -- selection, row ids, projected columns bit for bit --
for PAX and CS blocks, host and device resident macro images, several macro block sizes; broken macro blocks are refused with
the error the oracle's parser gives."""
import numpy as np
import pytest

import oracle_binding as ora

pytestmark = pytest.mark.gpu


def make_table(cs=False, n=120_000, rpb=900, seed=3):
    from oceanbase_b200 import capi
    from oceanbase_b200.sstable import Column, encode_table
    rng = np.random.default_rng(seed)
    key = np.arange(n, dtype=np.int64) * 2 + 5
    a = rng.integers(0, 1 << 33, size=n, dtype=np.int64)
    b = rng.integers(0, 40, size=n, dtype=np.int64)
    nl = (rng.random(n) < 0.07).astype(np.uint8)
    strs = [b"name-%04d" % (i % 211) for i in range(n)]
    if cs:
        cols = [Column(capi.OBJ_INT, capi.ENC_CS_INTEGER, key), Column(capi.OBJ_INT, capi.ENC_CS_INTEGER, a, nulls=nl),
                Column(capi.OBJ_INT, capi.ENC_CS_INT_DICT, b), Column(capi.OBJ_VARCHAR, capi.ENC_CS_STR_DICT, strs)]
    else:
        cols = [Column(capi.OBJ_INT, capi.ENC_INTEGER_BASE_DIFF, key), Column(capi.OBJ_INT, capi.ENC_RAW, a, nulls=nl),
                Column(capi.OBJ_INT, capi.ENC_DICT, b), Column(capi.OBJ_VARCHAR, capi.ENC_DICT, strs)]
    types = [capi.OBJ_INT, capi.OBJ_INT, capi.OBJ_INT, capi.OBJ_VARCHAR]
    return encode_table(cols, rpb, rowkey_cnt=1), types


def scans_equal(r1, r2, n_proj, str_cols=()):
    assert r1.selected_rows == r2.selected_rows
    assert np.array_equal(r1.fetch_sel_offsets(), r2.fetch_sel_offsets())
    assert np.array_equal(r1.fetch_row_ids(), r2.fetch_row_ids())
    for i in range(n_proj):
        if i in str_cols:
            h1, o1 = r1.fetch_strings(i)
            h2, o2 = r2.fetch_strings(i)
            assert np.array_equal(o1, o2) and np.array_equal(h1, h2)
        else:
            d1, _, n1 = r1.fetch_col(i)
            d2, _, n2 = r2.fetch_col(i)
            assert np.array_equal(d1, d2) and np.array_equal(n1, n2)


@pytest.mark.parametrize("cs", [False, True])
@pytest.mark.parametrize("macro_size,on_device", [(2 << 20, False), (256 << 10, True), (64 << 10, False)])
def test_macro_image_scans_like_the_plain_image(cs, macro_size, on_device):
    import torch
    import oceanbase_b200 as ob
    from oceanbase_b200.sstable import build_macro_blocks
    table, types = make_table(cs=cs)
    mi = build_macro_blocks(table, types, 1, macro_block_size=macro_size)
    ctx = ob.ScanContext(0)
    plain = ob.PageBatch(ctx, table)
    keep = None
    if on_device:
        keep = torch.from_numpy(mi.image).cuda()
        mb = ob.PageBatch.from_macro_blocks(ctx, None, macro_size, mi.n_macro, device_ptr=keep.data_ptr())
    else:
        mb = ob.PageBatch.from_macro_blocks(ctx, mi.image, macro_size, mi.n_macro)
    assert mb.n_blocks == table.n_blocks and mb.total_rows == table.total_rows
    for b in (0, table.n_blocks // 2, table.n_blocks - 1):
        assert mb.block_info(b) == plain.block_info(b)
    f = ob.And([ob.White(1, ob.WHITE_OP_LT, [1 << 32]), ob.White(2, ob.WHITE_OP_IN, [3, 7, 11, 30])])
    r1 = plain.scan(f, [0, 1, 2, 3], want_row_ids=True)
    r2 = mb.scan(f, [0, 1, 2, 3], want_row_ids=True)
    scans_equal(r1, r2, 4, str_cols=(3,))
    r1 = plain.scan(None, [1, 0], want_row_ids=True)
    r2 = mb.scan(None, [1, 0], want_row_ids=True)
    scans_equal(r1, r2, 2)
    mb.close()
    plain.close()
    ctx.close()


def test_broken_macro_blocks_are_refused():
    import oceanbase_b200 as ob
    from oceanbase_b200.sstable import build_macro_blocks
    table, types = make_table(n=30_000)
    mi = build_macro_blocks(table, types, 1, macro_block_size=128 << 10)
    ctx = ob.ScanContext(0)
    O = ora.oracle()
    import ctypes as C
    O.ora_macro_block_micro_blocks.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32]
    cases = []
    bad = mi.image.copy(); bad[8] = 0x55; cases.append(bad)                                 # common header magic
    bad = mi.image.copy(); bad[(128 << 10) + 24 + 6] ^= 1; cases.append(bad)                 # second macro block: macro header magic
    bad = mi.image.copy(); bad[24 + 52] += 1; cases.append(bad)                              # micro_block_count_ one too many
    first_micro = 24 + 128 + 4 * 16 + 1
    bad = mi.image.copy(); bad[first_micro + 44] ^= 0x10; cases.append(bad)                  # data_zlength_ of the first micro-block
    bad = mi.image.copy(); bad[24 + 104] = 5; cases.append(bad)                              # compressor_type_ = zstd
    for i, bad in enumerate(cases):
        with pytest.raises(ob.ObGpuError) as e:
            ob.PageBatch.from_macro_blocks(ctx, bad, 128 << 10, mi.n_macro)
        assert e.value.code in (ob.OB_INVALID_DATA, ob.OB_NOT_SUPPORTED), i
        if i != 4:   # the oracle's walk refuses the same blocks
            m = 1 if i == 1 else 0
            blk = np.ascontiguousarray(bad[m * (128 << 10):(m + 1) * (128 << 10)])
            offs, szs, cnt = np.zeros(4096, dtype=np.int64), np.zeros(4096, dtype=np.int64), C.c_int32(0)
            assert O.ora_macro_block_micro_blocks(blk.ctypes.data, blk.size, offs.ctypes.data, szs.ctypes.data, 4096, C.byref(cnt), 0) != 0, i
    ctx.close()
