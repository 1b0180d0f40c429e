"""Skip index (pre-aggregated min / max / null count per micro-block): the aggregate-row format and the filter
verdicts, pinned to the reference's own unit tests.

* unittest/storage/blocksstable/test_agg_row_struct.cpp:209-293 (test_agg_row_serialize_arm): the 21 datums of
  that test (7 columns x min / max / null count, 8-byte ints, 27- / 29- / 16-byte values, NULL min / max) go
  through the writer and come back through the reader; :386-435 (test_agg_row): random (col, type) subsets with
  NULLs read back as given.
* unittest/storage/blocksstable/test_skip_index_filter.cpp:405-1420: per operator the cases a..h of the test
  (seed0 < seed1 < seed2 < seed3 < seed4 stand for ordered values, ref = seed1, BT / IN = [seed1, seed3]) with
  the expected ObBoolMask, for an integer column and a string column."""
import struct

import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
from oceanbase_b200 import White, And, Or

U, T, F = 0, 1, 2   # uncertain, always true, always false (sql::ObBoolMaskType)
ROWS = 1000

ARM_DATUMS = [  # test_agg_row_struct.cpp:271-299 -- (col, type, hex bytes or None for a NULL datum)
    (0, 0, "f2 cc e9 ce ff ff ff ff"), (0, 1, "f2 cc e9 ce ff ff ff ff"), (0, 2, "00 00 00 00 00 00 00 00"),
    (1, 0, "73 74 72 75 67 67 6c 65 73 20 61 72 6d 65 72 20 63 61 77 73 20 61 70 70 6c 79 20"),
    (1, 1, "73 74 72 75 67 67 6c 65 73 20 61 72 6d 65 72 20 63 61 77 73 20 61 70 70 6c 79 20"),
    (1, 2, "00 00 00 00 00 00 00 00"),
    (2, 0, "c5 be db 19 00 00 00 00"), (2, 1, "c5 be db 19 00 00 00 00"), (2, 2, "00 00 00 00 00 00 00 00"),
    (3, 0, "63 6f 6d 70 72 65 68 65 6e 64 69 6e 67 20 64 72 65 73 73 6d 61 6b 65 72 27 73 20 65 6c"),
    (3, 1, "63 6f 6d 70 72 65 68 65 6e 64 69 6e 67 20 64 72 65 73 73 6d 61 6b 65 72 27 73 20 65 6c"),
    (3, 2, "00 00 00 00 00 00 00 00"),
    (4, 0, "80 ac 3c 2e 72 00 00 00"), (4, 1, "80 ac 3c 2e 72 00 00 00"), (4, 2, "00 00 00 00 00 00 00 00"),
    (7, 0, "19 07 00 00 00 00 00 00 03 00 00 00 00 00 00 00"), (7, 1, "19 07 00 00 00 00 00 00 03 00 00 00 00 00 00 00"),
    (7, 2, "00 00 00 00 00 00 00 00"),
    (8, 0, None), (8, 1, None), (8, 2, "01 00 00 00 00 00 00 00"),
]


@pytest.mark.parametrize("version", [1, 2, 3])
def test_agg_row_serialize_arm_dataset(version):
    cells = [(c, t, None if h is None else bytes.fromhex(h.replace(" ", ""))) for c, t, h in ARM_DATUMS]
    row = ob.agg_row_write(cells, version)
    ver, length, cnt, pack = struct.unpack_from("<hhhH", row.tobytes(), 0)
    assert (ver, length, cnt) == (version, len(row), 7)
    # ObAggRowHeader: 1-byte column indexes, 2-byte cell positions once the row passes 255 bytes, 1-byte bitmap
    assert pack & 0x3f == 1 and (pack >> 6) & 7 == (2 if len(row) > 255 else 1) and (pack >> 12) & 0xf == 1
    for c, t, v in cells:
        got, prefix = ora.agg_row_read(row, c, t)
        assert got == v and not prefix
    for c in (5, 6, 9, 100):                              # columns without aggregates
        assert ora.agg_row_read(row, c, 0) == (None, False)
    assert ora.agg_row_read(row, 0, 3) == (None, False)   # SUM was never stored


def test_agg_row_random_subsets():
    rng = np.random.default_rng(9)
    for _ in range(200):
        want = {}
        while len(want) < 10:                             # test_agg_row: 10 distinct (col 0..4, type 0..5)
            key = (int(rng.integers(0, 5)), int(rng.integers(0, 6)))
            if key not in want:
                want[key] = None if rng.integers(0, 5) == 0 else struct.pack("<q", len(want))
        cells = [(c, t, v) for (c, t), v in want.items()]
        row = ob.agg_row_write(cells)
        for (c, t), v in want.items():
            assert ora.agg_row_read(row, c, t)[0] == v


def test_agg_row_wide_offsets_and_prefix_flags():
    big = bytes(range(40))
    cells = [(c, t, (big, t != 2) if t != 2 else struct.pack("<q", 5)) for c in (3, 300, 70000) for t in (0, 1, 2)]
    row = ob.agg_row_write(cells)
    pack = struct.unpack_from("<H", row.tobytes(), 6)[0]
    assert pack & 0x3f == 3 and (pack >> 6) & 7 == 2      # 3-byte column indexes, row longer than 255 bytes
    for c in (3, 300, 70000):
        assert ora.agg_row_read(row, c, 0) == (big, True) and ora.agg_row_read(row, c, 1) == (big, True)
        assert ora.agg_row_read(row, c, 2) == (struct.pack("<q", 5), False)
    v1 = ob.agg_row_write(cells, 1)                       # version 1 has no prefix bitmap
    assert ora.agg_row_read(v1, 300, 0) == (big, False) and len(v1) == len(row) - 3


# ---- filter verdicts ----------------------------------------------------------------------------------------
INT_SEED = {k: (k - 2) * 1000 - 7 for k in range(5)}           # seed0 < ... < seed4, negative and positive
STR_SEED = {0: b"apple", 1: b"banana", 2: b"bananas", 3: b"cherry", 4: b"d"}


def agg_row(kind, mn, mx, null_count):
    seed = INT_SEED if kind == "int" else STR_SEED
    enc = (lambda k: struct.pack("<q", seed[k])) if kind == "int" else (lambda k: seed[k])
    cells = [(1, ob.SK_IDX_MIN, None if mn is None else enc(mn)), (1, ob.SK_IDX_MAX, None if mx is None else enc(mx)),
             (1, ob.SK_IDX_NULL_COUNT, None if null_count is None else struct.pack("<q", null_count))]
    if all(v is None for _, _, v in cells):
        cells.append((0, ob.SK_IDX_NULL_COUNT, struct.pack("<q", 0)))   # a row needs one stored aggregate
    return ob.agg_row_write(cells)


ALLNULL = (None, None, ROWS)
NOAGG = (None, None, None)
CASES = {  # op: (constants as seeds, [((min seed, max seed, null count), expected mask)])
    ob.WHITE_OP_EQ: ((1,), [((2, 2, 0), F), ((0, 0, 0), F), (ALLNULL, F), ((1, 1, 0), T), ((0, 2, 0), U), (NOAGG, U)]),       # :405-494
    ob.WHITE_OP_NE: ((1,), [((1, 1, 0), F), (ALLNULL, F), ((1, 2, 0), U), ((0, 1, 0), U), ((0, 2, 0), U), ((0, 0, 0), T),
                            ((2, 2, 0), T), (NOAGG, U)]),                                                                        # :496-589
    ob.WHITE_OP_LT: ((1,), [((1, 1, 0), F), ((2, 2, 0), F), (ALLNULL, F), ((0, 2, 0), U), ((0, 0, 0), T), (NOAGG, U)]),         # :591-669
    ob.WHITE_OP_LE: ((1,), [((2, 2, 0), F), (ALLNULL, F), ((1, 1, 0), T), ((0, 0, 0), T), ((0, 2, 0), U), (NOAGG, U)]),         # :671-752
    ob.WHITE_OP_GT: ((1,), [((1, 1, 0), F), ((0, 0, 0), F), (ALLNULL, F), ((0, 2, 0), U), ((2, 2, 0), T), (NOAGG, U)]),         # :754-833
    ob.WHITE_OP_GE: ((1,), [((0, 0, 0), F), (ALLNULL, F), ((0, 2, 0), U), ((1, 1, 0), T), ((2, 2, 0), T), (NOAGG, U)]),         # :835-914
    ob.WHITE_OP_NU: ((), [((1, 1, 0), F), ((1, 1, ROWS - 1), U), ((1, 1, ROWS), T), (NOAGG, U)]),                                # :916-977
    ob.WHITE_OP_NN: ((), [(ALLNULL, F), ((1, 1, ROWS - 1), U), ((1, 1, 0), T), (NOAGG, U)]),                                     # :979-1042
    ob.WHITE_OP_BT: ((1, 3), [((4, 4, 0), F), ((0, 0, 0), F), (ALLNULL, F), ((0, 4, 0), U), ((0, 2, 0), U), ((2, 4, 0), U),
                              ((1, 1, 0), T), ((3, 3, 0), T), ((2, 2, 0), T), (NOAGG, U)]),                                      # :1044-1165
    ob.WHITE_OP_IN: ((1, 3), [((0, 0, 0), F), ((2, 2, 0), F), ((4, 4, 0), F), ((1, 1, 0), T), (ALLNULL, F), ((2, 4, 0), U),
                              ((0, 2, 0), U), ((0, 4, 0), U), (NOAGG, U)]),                                                      # :1167-1303
}
HAS_NULL = [  # test_has_null (:1305-1421): an always-true verdict is downgraded when some rows are NULL
    (ob.WHITE_OP_NU, (1,), (1, 1, ROWS - 1)), (ob.WHITE_OP_NN, (1,), (1, 1, ROWS - 1)), (ob.WHITE_OP_EQ, (1,), (1, 1, ROWS // 2)),
    (ob.WHITE_OP_NE, (1,), (0, 0, ROWS // 2)), (ob.WHITE_OP_LT, (1,), (0, 0, ROWS // 2)), (ob.WHITE_OP_LE, (1,), (0, 0, ROWS // 2)),
    (ob.WHITE_OP_GT, (1,), (2, 2, ROWS // 2)), (ob.WHITE_OP_GE, (1,), (2, 2, ROWS // 2)), (ob.WHITE_OP_BT, (1, 3), (2, 2, ROWS // 2)),
    (ob.WHITE_OP_IN, (1, 3), (1, 1, ROWS // 2)),
]
COL_TYPES = {"int": [ob.OBJ_INT, ob.OBJ_INT], "str": [ob.OBJ_INT, ob.OBJ_VARCHAR]}


def consts(kind, seeds, op):
    seed = INT_SEED if kind == "int" else STR_SEED
    return () if op in (ob.WHITE_OP_NU, ob.WHITE_OP_NN) else tuple(seed[k] for k in seeds)


@pytest.mark.parametrize("kind", ["int", "str"])
@pytest.mark.parametrize("op", sorted(CASES))
def test_reference_filter_cases(kind, op):
    seeds, cases = CASES[op]
    for (mn, mx, nc), want in cases:
        got = ora.skip_index_filter(agg_row(kind, mn, mx, nc), ROWS, COL_TYPES[kind], White(1, op, consts(kind, seeds, op)))
        assert got == want, (op, mn, mx, nc)
    if op not in (ob.WHITE_OP_NU, ob.WHITE_OP_NN):     # case g: a NULL constant never matches
        params = (None,) * len(seeds) if op != ob.WHITE_OP_IN else (None,)
        assert ora.skip_index_filter(agg_row(kind, 0, 0, 0), ROWS, COL_TYPES[kind], White(1, op, params)) == F


@pytest.mark.parametrize("kind", ["int", "str"])
def test_reference_has_null_cases(kind):
    for op, seeds, (mn, mx, nc) in HAS_NULL:
        got = ora.skip_index_filter(agg_row(kind, mn, mx, nc), ROWS, COL_TYPES[kind], White(1, op, consts(kind, seeds, op)))
        assert got == U, op


def test_missing_min_or_max_is_an_open_bound():
    # ObSkipIndexFilterExecutor::compare (:505-516): NULL min is below, NULL max above every constant
    v = INT_SEED
    row = agg_row("int", None, 1, 0)
    assert ora.skip_index_filter(row, ROWS, COL_TYPES["int"], White(1, ob.WHITE_OP_GT, (v[2],))) == F
    assert ora.skip_index_filter(row, ROWS, COL_TYPES["int"], White(1, ob.WHITE_OP_LE, (v[2],))) == T
    assert ora.skip_index_filter(row, ROWS, COL_TYPES["int"], White(1, ob.WHITE_OP_GE, (v[0],))) == U
    row = agg_row("int", 1, None, 0)
    assert ora.skip_index_filter(row, ROWS, COL_TYPES["int"], White(1, ob.WHITE_OP_LT, (v[0],))) == F
    assert ora.skip_index_filter(row, ROWS, COL_TYPES["int"], White(1, ob.WHITE_OP_GE, (v[0],))) == T
    assert ora.skip_index_filter(row, ROWS, COL_TYPES["int"], White(1, ob.WHITE_OP_IN, (v[0],))) == F
    assert ora.skip_index_filter(row, ROWS, COL_TYPES["int"], White(1, ob.WHITE_OP_IN, (v[0], v[3]))) == U


def test_tree_verdicts_follow_obboolmask():
    # ObBoolMask operator& / operator| (ob_pushdown_filter.h:133-158) through execute_skipping_filter
    v = INT_SEED
    row = agg_row("int", 1, 3, 0)
    t = White(1, ob.WHITE_OP_GE, (v[1],))   # always true
    f = White(1, ob.WHITE_OP_GT, (v[3],))   # always false
    u = White(1, ob.WHITE_OP_EQ, (v[2],))   # uncertain
    n = White(0, ob.WHITE_OP_EQ, (5,))      # column without aggregates: uncertain
    types = COL_TYPES["int"]
    for expr, want in ((And([t, t]), T), (And([t, u]), U), (And([u, f]), F), (And([f, t]), F), (And([t, n]), U),
                       (Or([f, f]), F), (Or([f, u]), U), (Or([u, t]), T), (Or([n, t]), T), (Or([f, n]), U),
                       (And([Or([f, t]), Or([f, f])]), F), (Or([And([t, t]), u]), T), (And([Or([u, f]), t]), U)):
        assert ora.skip_index_filter(row, ROWS, types, expr) == want
    assert ora.skip_index_filter(None, ROWS, types, And([t, f])) == U      # no aggregate data at all


def test_string_prefix_rules():
    # compare_for_non_pad_charset / compare_with_prefix (:415-452, :476-496), binary collation
    p40 = b"x" * 40
    row = ob.agg_row_write([(1, ob.SK_IDX_MIN, (p40, True)), (1, ob.SK_IDX_MAX, (p40, True)), (1, ob.SK_IDX_NULL_COUNT, struct.pack("<q", 0))])
    types = COL_TYPES["str"]
    chk = lambda op, c: ora.skip_index_filter(row, ROWS, types, White(1, op, (c,)))
    assert chk(ob.WHITE_OP_EQ, p40) == F                 # every value is longer than its 40-byte prefix
    assert chk(ob.WHITE_OP_GT, p40) == T
    assert chk(ob.WHITE_OP_EQ, p40 + b"yz") == U         # the prefix is a prefix of the constant: cannot tell
    assert chk(ob.WHITE_OP_LT, p40 + b"yz") == U
    assert chk(ob.WHITE_OP_LT, b"y") == T                # differs inside the prefix
    assert chk(ob.WHITE_OP_GE, b"y") == F
    assert chk(ob.WHITE_OP_GE, b"w" * 50) == T
    assert ora.skip_index_filter(row, ROWS, types, White(1, ob.WHITE_OP_IN, (p40, b"a"))) == F
    assert ora.skip_index_filter(row, ROWS, types, White(1, ob.WHITE_OP_IN, (p40 + b"q", b"a"))) == U


def test_block_aggregates_of_the_writer():
    rng = np.random.default_rng(4)
    n = 3000
    a = rng.integers(-500, 500, size=n, dtype=np.int64)
    na = (rng.random(n) < 0.1).astype(np.uint8)
    d = rng.integers(8000, 9000, size=n, dtype=np.int64)
    u = rng.integers(0, 1 << 63, size=n, dtype=np.int64) * 2 + 1
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(0, 60)), dtype=np.uint8)) for _ in range(n)]
    allnull = np.ones(n, dtype=np.uint8)
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, a, nulls=na), ob.Column(ob.OBJ_DATE, ob.ENC_RAW, d), ob.Column(ob.OBJ_UINT64, ob.ENC_RAW, u),
            ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, words), ob.Column(ob.OBJ_INT, ob.ENC_RAW, a, nulls=allnull)]
    rows, offs = ob.table_agg_rows(cols, [0, 1, 2, 3, 4], 700)
    assert len(offs) == 6 and offs[0] == 0 and offs[-1] == len(rows)
    for b in range(5):
        lo, hi = b * 700, min(n, (b + 1) * 700)
        row = rows[offs[b]:offs[b + 1]]
        assert np.array_equal(row, ob.block_agg_row(cols, [0, 1, 2, 3, 4], lo, hi - lo))
        ok = na[lo:hi] == 0
        assert ora.agg_row_read(row, 0, ob.SK_IDX_MIN) == (struct.pack("<q", int(a[lo:hi][ok].min())), False)
        assert ora.agg_row_read(row, 0, ob.SK_IDX_MAX) == (struct.pack("<q", int(a[lo:hi][ok].max())), False)
        assert ora.agg_row_read(row, 0, ob.SK_IDX_NULL_COUNT)[0] == struct.pack("<q", int(na[lo:hi].sum()))
        assert ora.agg_row_read(row, 1, ob.SK_IDX_MIN)[0] == struct.pack("<i", int(d[lo:hi].min()))        # 4-byte date datum
        uu = u[lo:hi].view(np.uint64)
        assert ora.agg_row_read(row, 2, ob.SK_IDX_MAX)[0] == struct.pack("<Q", int(uu.max()))
        assert ora.agg_row_read(row, 2, ob.SK_IDX_MIN)[0] == struct.pack("<Q", int(uu.min()))
        ws = words[lo:hi]
        assert ora.agg_row_read(row, 3, ob.SK_IDX_MIN) == (min(ws)[:40], len(min(ws)) > 40)
        assert ora.agg_row_read(row, 3, ob.SK_IDX_MAX) == (max(ws)[:40], len(max(ws)) > 40)
        assert ora.agg_row_read(row, 4, ob.SK_IDX_MIN) == (None, False) and ora.agg_row_read(row, 4, ob.SK_IDX_NULL_COUNT)[0] == struct.pack("<q", hi - lo)


def test_verdicts_never_contradict_the_rows():
    """Soundness over random tables: an always-false block selects no row, an always-true block every row."""
    rng = np.random.default_rng(17)
    n, rpb = 20_000, 500
    k = np.sort(rng.integers(0, 100_000, size=n, dtype=np.int64))           # clustered: the index can prune
    v = rng.integers(-50, 50, size=n, dtype=np.int64)
    nv = (rng.random(n) < 0.05).astype(np.uint8)
    nv[:rpb] = 1                                                             # one all-NULL block
    s = [b"k%06d" % (x // 7) for x in k.tolist()]
    cols = [ob.Column(ob.OBJ_INT, ob.ENC_RAW, k), ob.Column(ob.OBJ_INT, ob.ENC_RAW, v, nulls=nv), ob.Column(ob.OBJ_VARCHAR, ob.ENC_RAW, s)]
    table = ob.encode_table(cols, rpb)
    rows, offs = ob.table_agg_rows(cols, [0, 1, 2], rpb)
    types = [ob.OBJ_INT, ob.OBJ_INT, ob.OBJ_VARCHAR]
    flts = [White(0, ob.WHITE_OP_BT, (20_000, 40_000)), White(0, ob.WHITE_OP_IN, (int(k[10]), int(k[9000]), -4)),
            And([White(0, ob.WHITE_OP_GE, (30_000,)), White(0, ob.WHITE_OP_LT, (60_000,)), White(1, ob.WHITE_OP_NE, (3,))]),
            Or([White(0, ob.WHITE_OP_LT, (5_000,)), White(1, ob.WHITE_OP_NU, ()), White(2, ob.WHITE_OP_GT, (b"k012000",))]),
            And([Or([White(0, ob.WHITE_OP_GT, (90_000,)), White(2, ob.WHITE_OP_LE, (b"k001",))]), White(1, ob.WHITE_OP_NN, ())]),
            White(1, ob.WHITE_OP_NN, ()), White(2, ob.WHITE_OP_EQ, (s[7777],))]
    seen = set()
    for flt in flts:
        for b in range(table.n_blocks):
            blk = ora.Block(table.block(b))
            bits = blk.filter_tree(flt)
            m = ora.skip_index_filter(rows[offs[b]:offs[b + 1]], blk.row_count, types, flt)
            seen.add(m)
            if m == F:
                assert not bits.any()
            elif m == T:
                assert bits.all()
    assert seen == {U, T, F}
