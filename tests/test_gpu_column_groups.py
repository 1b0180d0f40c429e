"""Column-store tables: every column group is its own SSTable with its own micro-block boundaries. Filters run on their own group
(ObCGScanner::apply_filter), meet in a range bitmap (ObCGBitmap bit_and / bit_or / set_bitmap) and the projection groups decode the rows
it selects (ObCGRowScanner::get_next_rows(bitmap)) -- obgpu_cg_bitmap + obgpu_scan_bitmap vs the oracle scanning the same columns as
one row-store table."""
import numpy as np
import pytest

import oracle_binding as ora

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


def test_q6_over_four_column_groups(ob, ctx):
    from oceanbase_b200.synth import Q6_DATE_LO, Q6_DATE_HI
    rng = np.random.default_rng(6)
    n = 200_000
    ship = rng.integers(8036, 8036 + 2526, size=n, dtype=np.int64)
    disc = rng.integers(0, 11, size=n, dtype=np.int64)
    qty = rng.integers(1, 51, size=n, dtype=np.int64)
    price = qty * rng.integers(90_000, 200_001, size=n, dtype=np.int64)
    dn = (rng.random(n) < 0.03).astype(np.uint8)      # some NULL discounts: never selected by the BETWEEN
    # four column groups, four block sizes (row boundaries never line up), PAX and CS stores mixed
    cg_ship = ob.encode_table([ob.Column(ob.OBJ_DATE, ob.ENC_CS_INTEGER, ship)], 3001)
    cg_disc = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_DICT, disc, nulls=dn)], 1777)
    cg_qty = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_RAW, qty)], 4096)
    cg_price = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, price)], 997)
    b_ship, b_disc, b_qty, b_price = (ctx.open_batch(t) for t in (cg_ship, cg_disc, cg_qty, cg_price))
    bm = ob.CGBitmap(ctx, n, all_true=True)
    assert bm.popcnt() == n
    f_ship = b_ship.scan(ob.And([ob.White(0, ob.WHITE_OP_GE, (Q6_DATE_LO,)), ob.White(0, ob.WHITE_OP_LT, (Q6_DATE_HI,))]), [])
    bm.apply(f_ship, 0, "and")
    f_disc = b_disc.scan(ob.White(0, ob.WHITE_OP_BT, (5, 7)), [])
    bm.apply(f_disc, 0, "and")
    f_qty = b_qty.scan(ob.White(0, ob.WHITE_OP_LT, (24,)), [])
    bm.apply(f_qty, 0, "and")
    sel = (ship >= Q6_DATE_LO) & (ship < Q6_DATE_HI) & (disc >= 5) & (disc <= 7) & (dn == 0) & (qty < 24)
    assert bm.popcnt() == int(sel.sum())
    assert np.array_equal(bm.fetch().astype(bool), sel)
    assert bm.popcnt(1000, 77_777) == int(sel[1000:77_777].sum())
    # the projection groups decode only those rows
    r_price = b_price.scan_bitmap(bm, [0], want_row_ids=True)
    r_disc = b_disc.scan_bitmap(bm, [0])
    assert r_price.selected_rows == r_disc.selected_rows == int(sel.sum())
    p, _, pn = r_price.fetch_col(0)
    d, _, dnl = r_disc.fetch_col(0)
    assert np.array_equal(p.view(np.int64), price[sel]) and np.array_equal(d.view(np.int64), disc[sel])
    revenue = int(sum(int(a) * int(b) for a, b in zip(p.view(np.int64), d.view(np.int64))))
    # the same query on the four columns as ONE row-store table, by the oracle
    one = ob.encode_table([ob.Column(ob.OBJ_DATE, ob.ENC_RAW, ship), ob.Column(ob.OBJ_INT, ob.ENC_RAW, disc, nulls=dn),
                           ob.Column(ob.OBJ_INT, ob.ENC_RAW, qty), ob.Column(ob.OBJ_INT, ob.ENC_RAW, price)], 2000)
    flt = ob.And([ob.White(0, ob.WHITE_OP_GE, (Q6_DATE_LO,)), ob.White(0, ob.WHITE_OP_LT, (Q6_DATE_HI,)), ob.White(1, ob.WHITE_OP_BT, (5, 7)),
                  ob.White(2, ob.WHITE_OP_LT, (24,))])
    want = ora.scan_table(one, flt, [3, 1], [False, False], [8, 8])
    assert want["selected"] == r_price.selected_rows
    assert np.array_equal(want["data"][0], p) and np.array_equal(want["data"][1], d)
    assert revenue == int(sum(int(a) * int(b) for a, b in zip(want["data"][0].view(np.int64), want["data"][1].view(np.int64))))
    # per-block tables of a bitmap scan behave like a filter scan's: sel_offset, row ids, aggregates
    so = r_price.fetch_sel_offsets()
    assert so[-1] == r_price.selected_rows and len(so) == cg_price.n_blocks + 1
    rid = r_price.fetch_row_ids()
    glob = np.concatenate([np.arange(0, n, 997)[b] + rid[so[b]:so[b + 1]] for b in range(cg_price.n_blocks)])
    assert np.array_equal(glob, np.nonzero(sel)[0])
    assert r_price.aggregate(ob.AGG_SUM, 0, -1) == int(price[sel].sum())
    for r in (f_ship, f_disc, f_qty, r_price, r_disc):
        r.free()
    # OR / SET at an offset: a second row range appended behind the first
    bm2 = ob.CGBitmap(ctx, 2 * n + 13)
    f = b_qty.scan(ob.White(0, ob.WHITE_OP_GE, (40,)), [])
    bm2.apply(f, 0, "set")
    bm2.apply(f, n + 13, "or")
    g = b_disc.scan(ob.White(0, ob.WHITE_OP_EQ, (3,)), [])
    bm2.apply(g, n + 13, "or")
    want2 = np.zeros(2 * n + 13, dtype=bool)
    want2[:n] = qty >= 40
    want2[n + 13:] = (qty >= 40) | ((disc == 3) & (dn == 0))
    assert np.array_equal(bm2.fetch().astype(bool), want2)
    r = b_price.scan_bitmap(bm2, [0], row_offset=n + 13)
    assert np.array_equal(r.fetch_col(0)[0].view(np.int64), price[want2[n + 13:]])
    r.free(); f.free(); g.free()
    bm.free(); bm2.free()
    for b in (b_ship, b_disc, b_qty, b_price):
        b.close()
