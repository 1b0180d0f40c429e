"""Thread-scaling of the CPU reference leg (oracle port) on the box's host cores.
Writes gpurun_out/cpu_scaling.json. Test/bench infrastructure only."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import oracle_binding as ora

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
w, _ = bench.build_workload(rows, 0, 2)
out = {"rows": rows, "ncpu": os.cpu_count(), "runs": []}
for nt in (1, 16, 32, 64, 96, 128, 256):
    if nt == 1:
        lim = w.table.n_blocks // 16
    else:
        lim = None
    best = 0
    for _ in range(3):
        t0 = time.perf_counter()
        r, s, _c = ora.scan_table_mt(w.table, w.filter, w.proj, batch_size=256, n_threads=nt, block_limit=lim)
        dt = time.perf_counter() - t0
        best = max(best, r / dt)
    out["runs"].append({"threads": nt, "rows": r, "rows_per_s": best})
    print(nt, r, best / 1e9, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_scaling.json"), "w"), indent=1)
