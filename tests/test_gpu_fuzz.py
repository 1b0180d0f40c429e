"""Randomised differential test: random schemas (every codec of both block formats, NULL fractions, value shapes that
make the writer pick different layouts), random filter trees with constants drawn from the data, random projections,
block sizes and staging modes -- device scan vs oracle scan, bit for bit, with and without the skip index."""
import numpy as np
import pytest

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


def random_int_values(rng, n, obj_type, ob):
    lo, hi = {ob.OBJ_INT: (-(1 << 62), 1 << 62), ob.OBJ_INT32: (-(1 << 31), 1 << 31), ob.OBJ_DATE: (-30000, 30000),
              ob.OBJ_TINYINT: (-128, 128), ob.OBJ_SMALLINT: (-(1 << 15), 1 << 15), ob.OBJ_UINT32: (0, 1 << 32),
              ob.OBJ_UINT64: (0, 1 << 62)}[obj_type]
    shape = rng.integers(0, 5)
    if shape == 0:      # full range
        v = rng.integers(lo, hi, size=n, dtype=np.int64)
    elif shape == 1:    # narrow band somewhere in the range
        w = min(int(rng.integers(1, 5000)), hi - lo)      # stay inside the type's domain
        a = int(rng.integers(lo, max(lo + 1, hi - w)))
        v = rng.integers(a, a + w, size=n, dtype=np.int64)
    elif shape == 2:    # few distinct values
        d = rng.integers(lo, hi, size=int(rng.integers(1, 12)), dtype=np.int64)
        v = d[rng.integers(0, len(d), size=n)]
    elif shape == 3:    # runs
        run = int(rng.integers(2, 40))
        d = rng.integers(max(lo, -1000), min(hi, 1000), size=n // run + 1, dtype=np.int64)
        v = np.repeat(d, run)[:n]
    else:               # one dominant value with a few exceptions
        v = np.full(n, int(rng.integers(max(lo, -5), min(hi, 5))), dtype=np.int64)
        k = rng.choice(n, size=max(1, n // 50), replace=False)
        v[k] = rng.integers(max(lo, -100), min(hi, 100), size=len(k))
    if obj_type == ob.OBJ_UINT64 and rng.integers(0, 2):
        v = (v.astype(np.uint64) * np.uint64(4) + np.uint64(3)).view(np.int64)   # above 2^63
    if rng.integers(0, 3) == 0:
        v = np.sort(v.view(np.uint64)).view(np.int64) if obj_type in (ob.OBJ_UINT64, ob.OBJ_UINT32) else np.sort(v)
    return v


def random_strings(rng, n):
    shape = rng.integers(0, 4)
    card = int(rng.integers(1, 60))
    if shape == 0:
        d = [bytes(rng.integers(97, 123, size=int(rng.integers(0, 18)), dtype=np.uint8)) for _ in range(card)]
    elif shape == 1:      # fixed length
        ln = int(rng.integers(1, 9))
        d = [bytes(rng.integers(97, 123, size=ln, dtype=np.uint8)) for _ in range(card)]
    elif shape == 2:      # long (beyond the 40-byte skip-index prefix), sharing prefixes
        d = [b"p" * int(rng.integers(38, 44)) + bytes(rng.integers(97, 100, size=int(rng.integers(0, 6)), dtype=np.uint8)) for _ in range(card)]
    else:
        d = [b"", b"a", b"ab", b"abc", b"b"][:max(1, card % 6)]
    idx = rng.integers(0, len(d), size=n)
    if rng.integers(0, 4) == 0:
        idx[:] = idx[0]
        k = rng.choice(n, size=max(1, n // 60), replace=False)
        idx[k] = rng.integers(0, len(d), size=len(k))
    return [d[i] for i in idx]


def random_case(ob, seed):
    rng = np.random.default_rng(seed)
    cs = bool(rng.integers(0, 2))
    n = int(rng.integers(200, 9000))
    rpb = int(rng.choice([37, 256, 700, 1400, 3000]))
    ncol = int(rng.integers(1, 7))
    int_types = [ob.OBJ_INT, ob.OBJ_INT32, ob.OBJ_DATE, ob.OBJ_TINYINT, ob.OBJ_UINT64, ob.OBJ_UINT32]
    cols, meta = [], []
    for _ in range(ncol):
        is_str = rng.integers(0, 3) == 0
        null_frac = float(rng.choice([0.0, 0.0, 0.1, 0.6, 1.0]))
        nulls = (rng.random(n) < null_frac).astype(np.uint8) if null_frac else None
        if is_str:
            v = random_strings(rng, n)
            enc = rng.choice([ob.ENC_CS_STRING, ob.ENC_CS_STR_DICT]) if cs else rng.choice([ob.ENC_RAW, ob.ENC_DICT, ob.ENC_RLE, ob.ENC_CONST])
            t = ob.OBJ_VARCHAR
        else:
            t = int(rng.choice(int_types))
            v = random_int_values(rng, n, t, ob)
            enc = rng.choice([ob.ENC_CS_INTEGER, ob.ENC_CS_INT_DICT]) if cs else \
                rng.choice([ob.ENC_RAW, ob.ENC_DICT, ob.ENC_RLE, ob.ENC_CONST, ob.ENC_INTEGER_BASE_DIFF])
        cols.append(ob.Column(t, int(enc), v, nulls=nulls))
        meta.append((t, is_str, v, nulls))
    # a span column next to some PAX tables: an integer COLUMN_EQUAL over one of the integer columns, a few exception rows (values,
    # NULL <-> value flips). Its own generator, so the tables of the seeds above stay what they were.
    srng = np.random.default_rng(770000 + seed)
    ints = [j for j, m in enumerate(meta) if not m[1]]
    if not cs and ints and srng.integers(0, 3) == 0:
        j = int(srng.choice(ints))
        t, _, v, nulls = meta[j]
        v2 = np.array(v, dtype=np.int64).copy()
        n2 = np.zeros(n, dtype=np.uint8) if nulls is None else nulls.copy()
        ex = srng.choice(n, size=max(1, n // 150), replace=False)
        v2[ex] = random_int_values(srng, len(ex), t, ob)
        flip = ex[:len(ex) // 3]
        n2[flip] ^= 1
        n2 = n2 if n2.any() else None
        try:   # the span encoder's own limits (exceptions per block, the BitSet's uint8 offsets) decide whether the column exists
            ob.encode_table([ob.Column(t, ob.ENC_RAW, v, nulls=nulls), ob.Column(t, ob.ENC_COLUMN_EQUAL, v2, nulls=n2, ref_col=0)], rpb)
            cols.append(ob.Column(t, ob.ENC_COLUMN_EQUAL, v2, nulls=n2, ref_col=j))
            meta.append((t, False, v2, n2))
        except ob.ObGpuError:
            pass
    return rng, cs, n, rpb, cols, meta


def encode_or_relax(ob, cols, rpb):
    """A forced CONST encoding needs <= 255 exception rows per block: fall back to DICT for those columns."""
    for attempt in range(3):
        try:
            return ob.encode_table(cols, rpb)
        except ob.ObGpuError as e:
            if e.code != ob.OB_NOT_SUPPORTED:
                raise
            for c in cols:
                if attempt == 0 and c.encoding == ob.ENC_CONST:
                    c.encoding = ob.ENC_DICT
                elif attempt == 1 and c.encoding in (ob.ENC_DICT, ob.ENC_RLE, ob.ENC_INTEGER_BASE_DIFF):
                    c.encoding = ob.ENC_RAW      # all-NULL dictionaries, base-diff on types it is not defined for
    return None


def random_leaf(ob, rng, meta, n):
    c = int(rng.integers(0, len(meta)))
    t, is_str, v, nulls = meta[c]
    op = int(rng.choice([ob.WHITE_OP_EQ, ob.WHITE_OP_NE, ob.WHITE_OP_LT, ob.WHITE_OP_LE, ob.WHITE_OP_GT, ob.WHITE_OP_GE, ob.WHITE_OP_BT,
                         ob.WHITE_OP_IN, ob.WHITE_OP_NU, ob.WHITE_OP_NN]))
    def const():
        x = v[int(rng.integers(0, n))]
        if is_str:
            r = rng.integers(0, 4)
            return x if r else (x + b"q" if r == 1 else x[:max(0, len(x) - 1)])
        x = int(np.uint64(x)) if t == ob.OBJ_UINT64 else int(x)
        x += int(rng.integers(-2, 3)) if rng.integers(0, 2) else 0
        if t == ob.OBJ_UINT64:
            x = min(max(x, 0), (1 << 64) - 1)
            return x - (1 << 64) if x >= (1 << 63) else x      # the C-ABI carries the 64-bit image
        return x
    if op in (ob.WHITE_OP_NU, ob.WHITE_OP_NN):
        params = ()
    elif op == ob.WHITE_OP_BT:
        a, b = const(), const()
        key = (lambda z: z) if is_str or t != ob.OBJ_UINT64 else (lambda z: z % (1 << 64))
        params = (a, b) if key(a) <= key(b) or rng.integers(0, 5) == 0 else (b, a)
    elif op == ob.WHITE_OP_IN:
        params = tuple(const() for _ in range(int(rng.integers(1, 5))))
        if rng.integers(0, 6) == 0:
            params = params + (None,)
    else:
        params = (const(),) if rng.integers(0, 12) else (None,)
    return ob.White(c, op, params)


def random_filter(ob, rng, meta, n, depth=0):
    r = rng.integers(0, 10)
    if depth >= 2 or r < 4:
        return random_leaf(ob, rng, meta, n)
    kids = [random_filter(ob, rng, meta, n, depth + 1) for _ in range(int(rng.integers(2, 4)))]   # <= 13 nodes, <= 45 constants
    return ob.And(kids) if r < 7 else ob.Or(kids)


@pytest.mark.parametrize("seed", range(60))
def test_random_tables_and_filters(ob, ctx, seed, monkeypatch):
    rng, cs, n, rpb, cols, meta = random_case(ob, 1000 + seed)
    table = encode_or_relax(ob, cols, rpb)
    if table is None:
        pytest.skip("the writer does not produce these forced encodings for this data")
    elem = [8 if s else {ob.OBJ_DATE: 4}.get(t, 8) for t, s, _, _ in meta]
    agg = ob.table_agg_rows(cols, list(range(len(cols))), rpb)
    if seed % 3 == 1:
        monkeypatch.setenv("OBGPU_PROJECT_COMPACT", str(seed % 2))
    for k in range(4):
        flt = random_filter(ob, rng, meta, n) if k else None
        proj = sorted(rng.choice(len(cols), size=int(rng.integers(1, len(cols) + 1)), replace=False).tolist())
        w = W(table, flt, proj, [meta[c][1] for c in proj], [elem[c] for c in proj])
        assert_scan_matches(ctx, w)
        if flt is not None:
            assert_scan_matches(ctx, w, agg=agg)                 # pruned by the skip index: same rows
            assert_scan_matches(ctx, w, max_selected_rows=max(16, n // 20) if k == 3 else 0, want_row_ids=k != 2) if k != 3 else None
