"""HEX_PACKING / STRING_DIFF / STRING_PREFIX on the device (mat_codecs.cuh: rebuilt once per page batch at open) vs the oracle:
per-block white filters and string projection, whole-table scans with such columns filtered AND projected (bytes through
obgpu_result_fetch_strings), next to ordinary columns, page batches opened from the host and from a device image."""
import numpy as np
import pytest

import oracle_binding as ora
from test_string_codecs import CASES

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


def heap_strings(heap, off, nulls):
    out = []
    for k in range(len(off) - 1):
        if (int(nulls[k // 64]) >> (k % 64)) & 1:
            out.append(None)
        else:
            out.append(bytes(heap[off[k]:off[k + 1]]))
    return out


@pytest.mark.parametrize("case", range(len(CASES)), ids=[c[0] for c in CASES])
def test_block_filters_and_projection(ob, ctx, case):
    name, enc, vals, nulls = CASES[case]
    n = len(vals)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, np.arange(n, dtype=np.int64)),
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, [b"x" * (i % 5) for i in range(n)]),
            ob.Column(ob.OBJ_VARCHAR, enc, vals, nulls=nulls), ob.Column(ob.OBJ_INT, ob.ENC_DICT, np.arange(n, dtype=np.int64) % 7)]
    table = ob.encode_table(cols, 110)
    batch = ctx.open_batch(table)
    assert batch.column_materialised(2) and not batch.column_materialised(1) and not batch.column_materialised(0)
    isnull = np.zeros(n, dtype=bool) if nulls is None else nulls.astype(bool)
    present = sorted({v for v, z in zip(vals, isnull) if not z})
    lo, hi = present[len(present) // 4], present[3 * len(present) // 4]
    row0 = 0
    # NOP cells (ext value 2) only occur in incremental SSTables, which a scan never sees un-fused; IS [NOT] NULL over them is the
    # one place the device (NOP reads as NULL) and the reference's retro path (a NOP datum is not NULL) differ: not compared
    has_nop = nulls is not None and bool(np.any(nulls == 2))
    for b in range(table.n_blocks):
        blk = ora.Block(table.block(b))
        rows = blk.row_count
        for flt in (ob.White(2, ob.WHITE_OP_EQ, (present[2],)), ob.White(2, ob.WHITE_OP_NE, (present[2],)), ob.White(2, ob.WHITE_OP_GE, (lo,)),
                    ob.White(2, ob.WHITE_OP_BT, (lo, hi)), ob.White(2, ob.WHITE_OP_IN, (present[0], present[-1], b"nope")),
                    ob.White(2, ob.WHITE_OP_NU, ()), ob.White(2, ob.WHITE_OP_NN, ()),
                    ob.And([ob.White(3, ob.WHITE_OP_LT, (5,)), ob.Or([ob.White(2, ob.WHITE_OP_LT, (lo,)), ob.White(1, ob.WHITE_OP_EQ, (b"xx",))])])):
            if has_nop and getattr(flt, "op", None) in (ob.WHITE_OP_NU, ob.WHITE_OP_NN):
                continue
            for start, count in ((0, None), (7, rows - 20)):
                assert np.array_equal(batch.filter_tree(b, flt, start, count), blk.filter_tree(flt, start, count)), (name, b, flt)
        rid = np.concatenate([np.arange(0, rows, 2), np.arange(rows - 1, 0, -9)]).astype(np.int32)
        heap, off, nl = batch.project_strings(b, 2, rid)
        want = [None if isnull[row0 + r] else vals[row0 + r] for r in rid]
        assert heap_strings(heap, off, nl) == want, (name, b)
        heap, off, nl = batch.project_strings(b, 1, rid)     # an ordinary string column through the same call
        assert heap_strings(heap, off, nl) == [b"x" * ((row0 + r) % 5) for r in rid]
        row0 += rows
    batch.close()
    ora.arena_reset()


@pytest.mark.parametrize("case", [0, 2, 5, 7, 10, 11, 12, 14], ids=lambda i: CASES[i][0])
@pytest.mark.parametrize("on_device", [False, True])
def test_scan_filters_and_projects_rebuilt_strings(ob, ctx, case, on_device):
    import torch
    name, enc, vals, nulls = CASES[case]
    reps = 40
    vals = vals * reps
    nulls = None if nulls is None else np.tile(nulls, reps)
    n = len(vals)
    k = np.arange(n, dtype=np.int64)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_INTEGER_BASE_DIFF, k), ob.Column(ob.OBJ_VARCHAR, enc, vals, nulls=nulls),
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, [b"d%d" % (i % 11) for i in range(n)])]
    table = ob.encode_table(cols, 700)
    base = table.image.ctypes.data
    if on_device:
        d = torch.empty(table.image.size + 64, dtype=torch.uint8, device="cuda:0")
        d[:table.image.size].copy_(torch.from_numpy(table.image))
        d[table.image.size:].zero_()
        torch.cuda.synchronize()
        batch = ctx.open_batch(table, device_image_ptr=d.data_ptr(), host_view=False, image_size=table.image.size)
    else:
        batch = ctx.open_batch(table)
    isnull = np.zeros(n, dtype=bool) if nulls is None else nulls.astype(bool)
    present = sorted({v for v, z in zip(vals, isnull) if not z})
    lo, hi = present[len(present) // 3], present[2 * len(present) // 3]
    for flt in (None, ob.White(1, ob.WHITE_OP_BT, (lo, hi)), ob.And([ob.White(0, ob.WHITE_OP_GE, (n // 5,)), ob.White(1, ob.WHITE_OP_NE, (present[1],))]),
                ob.Or([ob.White(1, ob.WHITE_OP_NU, ()), ob.White(2, ob.WHITE_OP_EQ, (b"d3",))])):
        want = ora.scan_table(table, flt, [0, 1, 2], [False, True, True], [8, 8, 8], string_base=base)
        res = batch.scan(flt, [0, 1, 2], string_base=base)
        assert res.selected_rows == want["selected"], (name, flt)
        data, _, nl0 = res.fetch_col(0)
        assert np.array_equal(data, want["data"][0])
        _, lens, nl1 = res.fetch_col(1)
        assert np.array_equal(nl1[:len(want["nulls"][1])], want["nulls"][1]) and np.array_equal(lens, want["lens"][1])
        heap, off = res.fetch_strings(1)
        assert heap_strings(heap, off, nl1) == ora.scan_strings(table, want, 1, base), (name, flt)
        # the ordinary dictionary column: pointers into the caller's image, and the same bytes through fetch_strings
        p2, l2, nl2 = res.fetch_col(2)
        assert np.array_equal(p2, want["data"][2]) and np.array_equal(l2, want["lens"][2])
        heap, off = res.fetch_strings(2)
        assert heap_strings(heap, off, nl2) == ora.scan_strings(table, want, 2, base)
        # a window of the rows
        if res.selected_rows > 50:
            heap, off = res.fetch_strings(1, 17, 30)
            full = ora.scan_strings(table, want, 1, base)[17:47]
            assert [bytes(heap[off[i]:off[i + 1]]) for i in range(30)] == [b"" if v is None else v for v in full]
        res.free()
        ora.arena_reset()
    batch.close()


def test_batch_with_cs_coded_streams_and_rebuilt_pax_strings(ob, ctx):
    """One page batch holding CS blocks whose integer streams are codec-coded (restated at open) AND PAX blocks with HEX_PACKING /
    STRING_PREFIX columns (materialised at open): the two passes compose -- string pointers of the untouched columns keep addressing
    the caller's image, rebuilt strings come back as bytes."""
    import ctypes as C
    rng = np.random.default_rng(21)
    n = 2600
    k = np.arange(n, dtype=np.int64) * 3
    hexs = [bytes(np.frombuffer(b"0123456789abcdef", dtype=np.uint8)[rng.integers(0, 16, size=rng.integers(1, 14))]) for _ in range(n)]
    plain = [b"p%d" % (i % 13) for i in range(n)]
    nl = (rng.random(n) < 0.1).astype(np.uint8)
    pax = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_INTEGER_BASE_DIFF, k), ob.Column(ob.OBJ_VARCHAR, ob.ENC_HEX_PACKING, hexs, nulls=nl),
                           ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, plain)], 400)
    cs_cols = [ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, k), ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STRING, hexs, nulls=nl),
               ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STR_DICT, plain)]
    cs_raw = ob.encode_table(cs_cols, 400)                  # what the oracle reads: the same blocks with RAW streams
    ob.capi.lib.obgpu_writer_set_cs_stream_encoding(0)      # detect: the sorted key column takes a delta codec
    try:
        cs = ob.encode_table(cs_cols, 400)
    finally:
        ob.capi.lib.obgpu_writer_set_cs_stream_encoding(1)
    assert cs.image.size < cs_raw.image.size                # some stream really is coded

    def interleave(cs_t):   # the two tables' blocks alternate in one image
        image = np.concatenate([pax.image, cs_t.image])
        offs = np.concatenate([pax.offsets, cs_t.offsets + pax.image.size])
        sizes = np.concatenate([pax.sizes, cs_t.sizes])
        order = np.argsort(np.concatenate([np.arange(pax.n_blocks) * 2, np.arange(cs_t.n_blocks) * 2 + 1]), kind="stable")
        return ob.TableImage(image, offs[order], sizes[order], pax.total_rows + cs_t.total_rows, 3)

    table, otable = interleave(cs), interleave(cs_raw)
    base, obase = table.image.ctypes.data, otable.image.ctypes.data
    batch = ctx.open_batch(table)
    assert batch.column_materialised(1)
    for flt in (None, ob.And([ob.White(0, ob.WHITE_OP_GE, (900,)), ob.White(1, ob.WHITE_OP_NE, (hexs[5],))]), ob.White(2, ob.WHITE_OP_IN, (b"p3", b"p7"))):
        want = ora.scan_table(otable, flt, [0, 1, 2], [False, True, True], [8, 8, 8], string_base=obase)
        res = batch.scan(flt, [0, 1, 2], string_base=base)
        assert res.selected_rows == want["selected"]
        assert np.array_equal(res.fetch_col(0)[0], want["data"][0])
        _, l1, n1 = res.fetch_col(1)
        heap, off = res.fetch_strings(1)
        assert heap_strings(heap, off, n1) == ora.scan_strings(otable, want, 1, obase)
        p2, l2, n2 = res.fetch_col(2)
        assert np.array_equal(l2, want["lens"][2])
        got2 = [C.string_at(int(p), int(l)) for p, l in zip(p2, l2)]        # still pointers into the caller's image
        assert got2 == ora.scan_strings(otable, want, 2, obase)
        res.free()
        ora.arena_reset()
    batch.close()
