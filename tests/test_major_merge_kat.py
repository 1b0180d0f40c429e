"""Major-compaction merge, oracle side. The reference's own row-fuse expectations
(unittest/storage/test_row_fuse.cpp:118-227: test_fuse_nomal, test_fuse_delete, test_fuse_empty) are
re-expressed as merges of single-row runs sharing one rowkey (rows listed newest first there = run index
descending here; strings / numbers stand in as integers); the K-way order, the delete drop of a major
merge and the default-row fill are checked against an independent dictionary model on config-5 runs."""
import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
from oceanbase_b200.synth import make_config5_runs

NOP, NULL, MAX = "NOP", "NULL", (1 << 62)
E, D, X = ob.DF_INSERT, ob.DF_DELETE, ob.DF_NOT_EXIST  # EXIST, DELETE, EMPTY of the reference's mock rows

# (rows newest first: (flag, 5 cells)), expected fused cells or None when the fused row is a delete
FUSE_CASES = {
    "fuse_normal": ([(E, [1, NOP, NOP, NOP, NOP]), (E, [2, 22, NOP, NOP, NULL]), (E, [3, 33, 333, NOP, 5555]),
                     (E, [4, 44, NOP, NOP, NOP]), (E, [5, 99, 999, MAX, NULL])], [1, 22, 333, MAX, NULL]),
    "delete_first": ([(D, [NOP] * 5), (E, [1, NOP, NOP, NOP, NOP]), (E, [2, 22, NOP, NOP, NULL]),
                      (E, [9, 99, 999, MAX, NULL])], None),
    "delete_in_the_middle": ([(E, [1, NOP, NOP, NOP, NOP]), (E, [2, 22, NOP, NOP, NULL]), (D, [NOP] * 5),
                              (E, [3, 33, 333, NOP, 5555]), (D, [NOP] * 5), (E, [4, 44, NOP, NOP, NOP]),
                              (E, [9, 99, 999, MAX, NULL])], [1, 22, NOP, NOP, NULL]),
    "delete_last": ([(E, [1, NOP, NOP, NOP, NOP]), (E, [2, 22, NOP, NOP, NULL]), (E, [3, 33, 333, NOP, 5555]),
                     (E, [4, 44, NOP, NOP, NOP]), (E, [9, 99, 999, MAX, NULL]), (D, [NOP] * 5)], [1, 22, 333, MAX, NULL]),
    "empty_then_delete": ([(X, [NOP] * 5), (X, [NOP] * 5), (D, [NOP] * 5), (E, [1, NOP, NOP, NOP, NOP]),
                           (E, [9, 99, 999, MAX, NULL])], None),
    "empty_interleaved": ([(X, [NOP] * 5), (X, [NOP] * 5), (E, [1, NOP, NOP, NOP, NOP]), (X, [NOP] * 5),
                           (E, [2, 22, NOP, NOP, NULL]), (X, [NOP] * 5), (D, [NOP] * 5), (E, [3, 33, 333, NOP, 5555]),
                           (E, [9, 99, 999, MAX, NULL])], [1, 22, NOP, NOP, NULL]),
    "empty_around": ([(X, [NOP] * 5), (E, [1, NOP, NOP, NOP, NOP]), (E, [2, 22, NOP, NOP, NULL]), (X, [NOP] * 5),
                      (X, [NOP] * 5), (E, [3, 33, 333, NOP, 5555]), (E, [4, 44, NOP, NOP, NOP]),
                      (E, [9, 99, 999, MAX, NULL]), (X, [NOP] * 5), (D, [NOP] * 5), (X, [NOP] * 5)], [1, 22, 333, MAX, NULL]),
    "only_empty": ([(X, [NOP] * 5), (X, [NOP] * 5)], None),
}


def case_runs(rows_newest_first, key=7):
    """One single-row run per listed row; the LAST listed row is the oldest table (run 0)."""
    runs = []
    for flag, cells in reversed(rows_newest_first):
        vals = [np.array([0 if c in (NOP, NULL) else c], dtype=np.int64) for c in cells]
        ext = [np.array([2 if c == NOP else (1 if c == NULL else 0)], dtype=np.uint8) for c in cells]
        runs.append({"key": np.array([key], dtype=np.int64), "flag": np.array([flag], dtype=np.uint8), "vals": vals, "ext": ext})
    return runs


def expected_row(expect, default_vals=None, default_null=None):
    """NOP left after the fuse takes the default row (ObMajorPartitionMergeFuser::end_fuse_row)."""
    out = []
    for c, x in enumerate(expect):
        if x == NOP:
            dn = 1 if default_null is None else default_null[c]
            out.append((1, 0) if dn else (0, default_vals[c]))
        elif x == NULL:
            out.append((1, 0))
        else:
            out.append((0, x))
    return out


@pytest.mark.parametrize("name", sorted(FUSE_CASES))
@pytest.mark.parametrize("defaults", [None, ([70, 71, 72, 73, 74], [0, 1, 0, 0, 1])])
def test_reference_fuse_expectations(name, defaults):
    rows, expect = FUSE_CASES[name]
    dv, dn = (None, None) if defaults is None else defaults
    m = ora.major_merge(case_runs(rows), 5, dv, dn)
    if expect is None:
        assert len(m["key"]) == 0
        assert m["dropped"] == (0 if name == "only_empty" else 1)
        return
    assert m["key"].tolist() == [7]
    want = expected_row(expect, dv, dn)
    got = [(int(m["null"][c][0]), int(m["vals"][c][0])) for c in range(5)]
    assert got == want


def model_merge(runs, n_cols):
    keys = {}
    for ri, r in enumerate(runs):
        for i, k in enumerate(r["key"].tolist()):
            keys.setdefault(k, []).append((ri, i))
    out = []
    for k in sorted(keys):
        first, dele, cells = True, False, [None] * n_cols
        for ri, i in sorted(keys[k], reverse=True):
            f = runs[ri]["flag"][i] if runs[ri].get("flag") is not None else ob.DF_INSERT
            if f == ob.DF_NOT_EXIST:
                continue
            if f == ob.DF_DELETE:
                dele = first
                break
            for c in range(n_cols):
                e = int(runs[ri]["ext"][c][i])
                if first or (cells[c][0] == 2 and e != 2):
                    cells[c] = (e, 0 if e else int(runs[ri]["vals"][c][i]))
            first = False
            if all(x[0] != 2 for x in cells):
                break
        if dele or first:
            continue
        out.append((k, [(1, 0) if x[0] == 2 else x for x in cells]))
    return out


@pytest.mark.parametrize("n_runs,window", [(1, 3000), (2, 5000), (3, 4000), (8, 6000)])
def test_config5_merge_matches_dictionary_model(n_runs, window):
    runs = make_config5_runs(n_runs=n_runs, window=window, seed=11, encode=False)
    m = ora.major_merge(runs, 3)
    want = model_merge(runs, 3)
    assert m["key"].tolist() == [k for k, _ in want]
    for c in range(3):
        assert m["null"][c].tolist() == [cells[c][0] for _, cells in want]
        assert m["vals"][c].tolist() == [cells[c][1] for _, cells in want]
    assert (np.diff(m["key"]) > 0).all()


def test_runs_roundtrip_through_the_writer_with_nop_cells():
    # NOP is an ObStoredExtValue (2 ext bits per cell once a column of the block holds a NOP)
    runs = make_config5_runs(n_runs=3, window=8000, seed=12, rows_per_block=700)
    for r in runs:
        k, e = ora.decode_column_ext(r["table"], 0)
        assert np.array_equal(k, r["key"]) and not e.any()
        f, _ = ora.decode_column_ext(r["table"], 1)
        assert np.array_equal(f, r["flag"])
        for c in range(3):
            v, e = ora.decode_column_ext(r["table"], 2 + c)
            assert np.array_equal(e, r["ext"][c]) and np.array_equal(v, r["vals"][c])
        blk = ora.Block(r["table"].block(0))
        assert blk.verify_checksums() == 0


def test_unsorted_run_is_rejected():
    r = {"key": np.array([5, 5, 9], dtype=np.int64), "flag": None, "vals": [], "ext": []}
    with pytest.raises(Exception):
        ora.major_merge([r], 0)


def composite_runs(rng, n_runs, n_keys, n_more):
    """Runs with a composite INT64 rowkey whose FIRST column ties a lot (so the later columns decide), one payload column."""
    runs = []
    universe = np.stack([rng.integers(0, 40, size=n_keys)] + [rng.integers(-5, 5, size=n_keys) for _ in range(n_more)], axis=1)
    universe = np.unique(universe, axis=0)                                  # sorted lexicographically, distinct
    for r in range(n_runs):
        pick = np.sort(rng.choice(len(universe), size=int(len(universe) * 0.5), replace=False))
        keys = universe[pick]
        n = len(keys)
        flag = np.where(rng.random(n) < 0.08, 4, 3 if r == 0 else 2).astype(np.uint8)
        runs.append(dict(key=keys[:, 0].astype(np.int64), more_keys=[keys[:, 1 + c].astype(np.int64) for c in range(n_more)],
                         flag=flag, vals=[rng.integers(0, 1000, size=n, dtype=np.int64)],
                         ext=[rng.choice([0, 0, 1, 2] if r else [0, 0, 1], size=n).astype(np.uint8)]))
    return runs


def composite_model(runs, n_more):
    rows = {}
    for ri, r in enumerate(runs):
        for i in range(len(r["key"])):
            k = (int(r["key"][i]),) + tuple(int(r["more_keys"][c][i]) for c in range(n_more))
            rows.setdefault(k, []).append((ri, i))
    keys, vals, nulls, dropped, fused = [], [], [], 0, 0
    for k in sorted(rows):
        cell, first, deleted = None, True, False
        for ri, i in sorted(rows[k], reverse=True):
            r = runs[ri]
            if r["flag"][i] == 4:
                deleted = first
                break
            first = False
            if cell is None and r["ext"][0][i] != 2:
                cell = (int(r["ext"][0][i]), 0 if r["ext"][0][i] else int(r["vals"][0][i]))
                break
        if deleted:
            dropped += 1
            continue
        if first:
            continue
        keys.append(k)
        nulls.append(1 if cell is None else cell[0])
        vals.append(0 if cell is None else cell[1])
        fused += len(rows[k]) > 1
    return keys, vals, nulls, dropped, fused


@pytest.mark.parametrize("n_more", [1, 2, 3])
def test_composite_rowkeys_in_the_oracle(n_more):
    # rowkeys compare column by column (ObStorageDatumUtils); the first column alone does not separate the groups
    rng = np.random.default_rng(60 + n_more)
    runs = composite_runs(rng, 5, 3000, n_more)
    got = ora.major_merge(runs, 1)
    keys, vals, nulls, dropped, fused = composite_model(runs, n_more)
    assert len(got["key"]) == len(keys) and got["dropped"] == dropped and got["fused"] == fused
    assert [(int(got["key"][i]),) + tuple(int(got["more_keys"][c][i]) for c in range(n_more)) for i in range(len(keys))] == keys
    assert got["null"][0].tolist() == nulls and got["vals"][0].tolist() == vals
    # a run that is not strictly ascending in the composite order is refused
    bad = composite_runs(rng, 2, 500, n_more)
    bad[1]["more_keys"][0][1] = bad[1]["more_keys"][0][0]
    bad[1]["key"][1] = bad[1]["key"][0]
    for c in range(1, n_more):
        bad[1]["more_keys"][c][1] = bad[1]["more_keys"][c][0]
    with pytest.raises(RuntimeError):
        ora.major_merge(bad, 1)
