"""Skip-index pruning on the device vs the oracle: per-block verdicts of obgpu_batch_skip_index_filter equal the
restated ObSSTableIndexFilter::check_range, scans with aggregate rows attached return exactly the rows of an
unpruned oracle scan, and pruned blocks are really not read."""
import numpy as np
import pytest

import oracle_binding as ora
from test_gpu_scan import assert_scan_matches

pytestmark = pytest.mark.gpu

U, T, F = 0, 1, 2


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


def make_table(ob, n=40_000, rpb=900, seed=5, long_strings=False):
    rng = np.random.default_rng(seed)
    k = np.sort(rng.integers(-50_000, 50_000, size=n, dtype=np.int64))          # clustered key: prunable
    v = rng.integers(-50, 50, size=n, dtype=np.int64)
    nv = (rng.random(n) < 0.05).astype(np.uint8)
    nv[:rpb] = 1                                                                 # block 0 of column 1: all NULL
    nv[rpb:2 * rpb] = 0                                                          # block 1: no NULL
    d = np.sort(rng.integers(8000, 9500, size=n, dtype=np.int64))
    u = np.sort((rng.integers(0, 1 << 62, size=n, dtype=np.int64) * 4).view(np.uint64)).view(np.int64)     # above 2^63 too
    pad = b"-" * 45 if long_strings else b""
    s = [pad + b"k%07d" % (x // 7 + 8000) for x in k.tolist()]
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_INTEGER_BASE_DIFF, k), ob.Column(ob.OBJ_INT, ob.ENC_RAW, v, nulls=nv),
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, s), ob.Column(ob.OBJ_DATE, ob.ENC_RAW, d), ob.Column(ob.OBJ_UINT64, ob.ENC_RAW, u)]
    table = ob.encode_table(cols, rpb)
    rows, offs = ob.table_agg_rows(cols, [0, 1, 2, 3, 4], rpb)
    types = [ob.OBJ_INT, ob.OBJ_INT, ob.OBJ_VARCHAR, ob.OBJ_DATE, ob.OBJ_UINT64]
    return dict(table=table, agg=(rows, offs), types=types, k=k, v=v, s=s, d=d, u=u, rpb=rpb)


def filters(ob, t):
    k, s, d, u = t["k"], t["s"], t["d"], t["u"]
    W_ = ob.White
    return [
        W_(0, ob.WHITE_OP_BT, (-10_000, 10_000)), W_(0, ob.WHITE_OP_LT, (int(k[3000]),)), W_(0, ob.WHITE_OP_GE, (int(k[-2000]),)),
        W_(0, ob.WHITE_OP_EQ, (int(k[12345]),)), W_(0, ob.WHITE_OP_NE, (int(k[0]),)),
        W_(0, ob.WHITE_OP_IN, (int(k[10]), int(k[20000]), 123456789)), W_(0, ob.WHITE_OP_IN, (999_999,)),
        W_(1, ob.WHITE_OP_NU, ()), W_(1, ob.WHITE_OP_NN, ()), W_(1, ob.WHITE_OP_GT, (-51,)), W_(1, ob.WHITE_OP_LE, (None,)),
        W_(2, ob.WHITE_OP_GT, (s[30000],)), W_(2, ob.WHITE_OP_EQ, (s[777],)), W_(2, ob.WHITE_OP_BT, (s[5000], s[9000])),
        W_(2, ob.WHITE_OP_IN, (s[100], s[39000], b"zzz")), W_(2, ob.WHITE_OP_LE, (s[2000][:-3],)), W_(2, ob.WHITE_OP_NE, (s[5],)),
        W_(3, ob.WHITE_OP_BT, (8500, 8600)), W_(3, ob.WHITE_OP_GT, (int(d[-100]),)),
        W_(4, ob.WHITE_OP_GE, (int(u[20000]),)), W_(4, ob.WHITE_OP_LT, (int(u[100]),)),
        ob.And([W_(0, ob.WHITE_OP_GE, (-20_000,)), W_(0, ob.WHITE_OP_LE, (5_000,))]),                          # merged into one range on device
        ob.And([W_(0, ob.WHITE_OP_GE, (0,)), W_(0, ob.WHITE_OP_LT, (30_000,)), W_(1, ob.WHITE_OP_NE, (3,)), W_(3, ob.WHITE_OP_LE, (9300,))]),
        ob.Or([W_(0, ob.WHITE_OP_LT, (-45_000,)), W_(1, ob.WHITE_OP_NU, ()), W_(2, ob.WHITE_OP_GT, (s[38000],))]),
        ob.And([ob.Or([W_(0, ob.WHITE_OP_GT, (40_000,)), W_(2, ob.WHITE_OP_LE, (s[1500],))]), W_(1, ob.WHITE_OP_NN, ())]),
        ob.Or([ob.And([W_(0, ob.WHITE_OP_GE, (10_000,)), W_(0, ob.WHITE_OP_LE, (12_000,))]), W_(3, ob.WHITE_OP_EQ, (int(d[50]),))]),
    ]


def oracle_verdicts(t, flt):
    table, (rows, offs) = t["table"], t["agg"]
    out = np.zeros(table.n_blocks, dtype=np.uint8)
    for b in range(table.n_blocks):
        rc = min(t["rpb"], table.total_rows - b * t["rpb"])
        out[b] = ora.skip_index_filter(rows[offs[b]:offs[b + 1]], rc, t["types"], flt)
    return out


@pytest.mark.parametrize("long_strings", [False, True])
def test_block_verdicts_match_the_oracle(ob, ctx, long_strings):
    t = make_table(ob, long_strings=long_strings)
    batch = ctx.open_batch(t["table"])
    flt0 = filters(ob, t)[0]
    assert not batch.skip_index_filter(flt0).any()               # no aggregate rows attached: all uncertain
    batch.set_agg_rows(*t["agg"])
    seen = set()
    for i, flt in enumerate(filters(ob, t)):
        want = oracle_verdicts(t, flt)
        got = batch.skip_index_filter(flt)
        assert np.array_equal(got, want), (i, np.nonzero(got != want)[0][:5], got[got != want][:5], want[got != want][:5])
        seen.update(want.tolist())
    assert seen == {U, T, F}
    batch.set_agg_rows(None)
    assert not batch.skip_index_filter(flt0).any()
    batch.close()


@pytest.mark.parametrize("long_strings", [False, True])
def test_pruned_scans_return_the_same_rows(ob, ctx, long_strings):
    t = make_table(ob, n=40_000, rpb=700, seed=6, long_strings=long_strings)
    for flt in filters(ob, t):
        assert_scan_matches(ctx, W(t["table"], flt, [0, 1, 2, 3], [False, False, True, False], [8, 8, 8, 4]), agg=t["agg"])


def test_skip_info_and_unread_blocks(ob, ctx):
    t = make_table(ob, n=50_000, rpb=1000, seed=7)
    table = t["table"]
    flt = ob.White(0, ob.WHITE_OP_BT, (-10_000, 10_000))
    verdicts = oracle_verdicts(t, flt)
    assert (verdicts == F).sum() > 10 and (verdicts == T).sum() > 3
    want = ora.scan_table(table, flt, [0, 1], [False, False], [8, 8])
    # move the key column of every always-false block into the predicate's range by zeroing its INTEGER_BASE_DIFF
    # base (the aggregate rows still describe the original data): if the count kernel read such a block, its rows
    # would now be selected
    image = table.image.copy()
    for b in np.nonzero(verdicts == F)[0]:
        lo = int(table.offsets[b])
        meta = 64 + 16 * table.n_cols + int(image[lo + 64 + 8:lo + 64 + 12].view(np.uint32)[0])
        assert image[lo + 64 + 1] == 4                         # ObColumnHeader::INTEGER_BASE_DIFF
        image[lo + meta + 2:lo + meta + 10] = 0
    broken = ob.TableImage(image, table.offsets, table.sizes, table.total_rows, table.n_cols)
    batch = ctx.open_batch(broken)
    batch.set_agg_rows(*t["agg"])
    res = batch.scan(flt, [0, 1])
    assert res.selected_rows == want["selected"]
    assert res.skip_info() == (int((verdicts == F).sum()), int((verdicts == T).sum()))
    for c in range(2):
        data, _, nulls = res.fetch_col(c)
        assert np.array_equal(data, want["data"][c]) and np.array_equal(nulls, want["nulls"][c])
    assert np.array_equal(res.fetch_sel_offsets(), want["sel_offset"])
    res.free()
    # without the index the wrecked blocks are read: the result must differ (the test would be vacuous otherwise)
    batch.set_agg_rows(None)
    res = batch.scan(flt, [0])
    assert res.selected_rows > want["selected"]
    res.free()
    batch.close()


def test_partial_and_missing_aggregate_rows(ob, ctx):
    t = make_table(ob, n=40_000, rpb=600, seed=8)
    rows, offs = t["agg"]
    # every second block loses its aggregate row (ObMicroIndexInfo::has_agg_data() false): uncertain there
    keep = np.arange(t["table"].n_blocks) % 2 == 0
    parts, new_off = [], [0]
    for b in range(t["table"].n_blocks):
        if keep[b]:
            parts.append(rows[offs[b]:offs[b + 1]])
        new_off.append(new_off[-1] + (int(offs[b + 1] - offs[b]) if keep[b] else 0))
    agg = (np.concatenate(parts), np.array(new_off, dtype=np.int64))
    batch = ctx.open_batch(t["table"])
    batch.set_agg_rows(*agg)
    flt = ob.And([ob.White(0, ob.WHITE_OP_GE, (0,)), ob.White(3, ob.WHITE_OP_LT, (9000,))])
    got = batch.skip_index_filter(flt)
    want = oracle_verdicts(t, flt)
    assert np.array_equal(got[keep], want[keep]) and not got[~keep].any()
    batch.close()
    assert_scan_matches(ctx, W(t["table"], flt, [0, 3], [False, False], [8, 4]), agg=agg)
