"""common::ObBitmap::get_row_ids semantics (lib/container/ob_bitmap.cpp:300-333,540-561; rule 8c.3)."""
import numpy as np

import oracle_binding as ora


def test_get_row_ids_limit_and_from_advance():
    rng = np.random.default_rng(5)
    bm = (rng.random(1000) < 0.3).astype(np.uint8)
    ids_all = np.flatnonzero(bm)
    frm, got = 0, []
    while frm < 1000:
        ids, frm = ora.bitmap_get_row_ids(bm, frm, 1000, 256)
        got.extend(ids.tolist())
        assert len(ids) <= 256
    assert got == ids_all.tolist()


def test_get_row_ids_window_and_id_offset():
    bm = np.zeros(300, dtype=np.uint8)
    bm[[5, 100, 101, 250, 299]] = 1
    ids, frm = ora.bitmap_get_row_ids(bm, 100, 260, 10, id_offset=100)
    assert ids.tolist() == [0, 1, 150] and frm == 260
    ids, frm = ora.bitmap_get_row_ids(bm, 0, 300, 2)
    assert ids.tolist() == [5, 100] and frm == 101
    ids, frm = ora.bitmap_get_row_ids(bm, 101, 300, 2)
    assert ids.tolist() == [101, 250] and frm == 251
    ids, frm = ora.bitmap_get_row_ids(np.zeros(64, dtype=np.uint8), 0, 64, 256)
    assert len(ids) == 0 and frm == 64
