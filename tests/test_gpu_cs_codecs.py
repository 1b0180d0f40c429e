"""Device path on CS blocks whose integer streams carry the non-RAW codecs (K13 / row a16): the batch is restated as RAW at
open (stream_codecs.cuh) and every scan result must equal the oracle's on the RAW-written twin of the same table -- rows,
offsets, row ids, integer payloads, NULL images, string lengths and the pointed-to bytes (pointers still address the
CALLER's codec image)."""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
from oceanbase_b200.capi import lib
from test_cs_stream_codecs import MODES, make_cols

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ob.ScanContext(0)
    yield c
    c.close()


def twin_tables(n, rpb, seed, mode):
    lib.obgpu_writer_set_cs_stream_encoding(1)
    raw = ob.encode_table(make_cols(n, seed), rpb)
    try:
        assert lib.obgpu_writer_set_cs_stream_encoding(mode) == 0
        enc = ob.encode_table(make_cols(n, seed), rpb)
    finally:
        lib.obgpu_writer_set_cs_stream_encoding(1)
    return raw, enc


PROJ = [0, 1, 2, 3, 4, 5, 6, 7]
IS_STR = [False, False, False, False, False, True, True, False]
ELEM = [8, 8, 8, 4, 8, 8, 8, 8]


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("rpb", [130, 900])
def test_scan_of_codec_blocks_equals_the_raw_twin(ctx, mode, rpb):
    n = 4000
    raw, enc = twin_tables(n, rpb, 300 + mode, mode)
    assert sum(enc.sizes) != sum(raw.sizes)
    flt = ob.And([ob.White(1, ob.WHITE_OP_GE, (-300,)), ob.Or([ob.White(4, ob.WHITE_OP_LT, (20 * 1_000_003,)), ob.White(2, ob.WHITE_OP_NU, ())])])
    base = 0x20_0000_0000
    for f in (None, flt, ob.White(3, ob.WHITE_OP_BT, (8500, 9500))):
        want = ora.scan_table(raw, f, PROJ, IS_STR, ELEM, string_base=base)
        batch = ctx.open_batch(enc)
        res = batch.scan(f, PROJ, want_row_ids=True, string_base=base)
        k = res.selected_rows
        assert k == want["selected"]
        assert np.array_equal(res.fetch_sel_offsets(), want["sel_offset"])
        assert np.array_equal(res.fetch_row_ids(), want["row_ids"])
        for i in range(len(PROJ)):
            data, lens, nulls = res.fetch_col(i)
            assert np.array_equal(nulls, want["nulls"][i]), (MODES[mode], i)
            if not IS_STR[i]:
                assert np.array_equal(data, want["data"][i]), (MODES[mode], i)
                continue
            assert np.array_equal(lens, want["lens"][i])
            mine = (data - np.uint64(base)).astype(np.int64)
            theirs = (want["data"][i] - np.uint64(base)).astype(np.int64)
            isnull = (nulls[np.arange(k) // 64] >> (np.arange(k) % 64).astype(np.uint64)) & np.uint64(1)
            for r in range(0, k, max(k // 200, 1)):
                if isnull[r]:
                    continue
                a = enc.image[mine[r]:mine[r] + lens[r]]
                b = raw.image[theirs[r]:theirs[r] + lens[r]]
                assert np.array_equal(a, b), (MODES[mode], i, r)
        res.free()
        # reference-granularity calls on a restated batch
        blk = ora.Block(raw.block(1))
        if f is not None:
            assert np.array_equal(batch.filter_tree(1, f), blk.filter_tree(f))
        rid = np.arange(0, blk.row_count, 3, dtype=np.int32)
        d, nl, _ = batch.project_fixed(1, 2, rid)
        wd, wn, _ = blk.get_rows_fixed(2, rid)
        assert np.array_equal(nl, wn)
        nn = ~(((wn[np.arange(len(rid)) // 64] >> (np.arange(len(rid)) % 64).astype(np.uint64)) & np.uint64(1)).astype(bool))
        assert np.array_equal(d.view(np.uint64)[:len(rid)][nn], wd.view(np.uint64)[:len(rid)][nn])
        batch.close()
