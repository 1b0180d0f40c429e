"""Replays one case of tests/test_gpu_fuzz.py and reports where device and oracle differ (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oceanbase_b200 as ob
import oracle_binding as ora
import test_gpu_fuzz as F

seed = int(sys.argv[1])
rng, cs, n, rpb, cols, meta = F.random_case(ob, 1000 + seed)
table = F.encode_or_relax(ob, cols, rpb)
elem = [8 if s else {ob.OBJ_DATE: 4}.get(t, 8) for t, s, _, _ in meta]
ctx = ob.ScanContext(0)
base = 0x10_0000_0000
cases = []
for k in range(4):
    flt = F.random_filter(ob, rng, meta, n) if k else None
    proj = sorted(rng.choice(len(cols), size=int(rng.integers(1, len(cols) + 1)), replace=False).tolist())
    cases.append((flt, proj))
for c in range(len(cols)):
    cases.append((None, [c]))
for flt, proj in cases:
    batch = ctx.open_batch(table)
    res = batch.scan(flt, proj, string_base=base)
    want = ora.scan_table(table, flt, proj, [meta[c][1] for c in proj], [elem[c] for c in proj], string_base=base)
    nsel = res.selected_rows
    line = f"filter={flt is not None} proj={proj} selected {nsel} vs {want['selected']}"
    for i, c in enumerate(proj):
        data, lens, nulls = res.fetch_col(i)
        bad = np.nonzero(data != want["data"][i])[0]
        if len(bad):
            so = want["sel_offset"]
            blk = int(np.searchsorted(so, bad[0], side="right") - 1)
            line += f" | col {c} enc {cols[c].encoding}: {len(bad)} payload mismatches, first at dense row {bad[0]} (block {blk}, row {bad[0] - so[blk]}), dev {data[bad[0]] - (base if meta[c][1] else 0)} ora {want['data'][i][bad[0]] - (base if meta[c][1] else 0)}"
        if lens is not None and not np.array_equal(lens, want["lens"][i]):
            line += f" | col {c} lens differ"
        if not np.array_equal(nulls, want["nulls"][i]):
            line += f" | col {c} nulls differ"
    print(line)
    res.free()
    batch.close()
