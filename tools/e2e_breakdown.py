"""Where the end-to-end (host buffers in, host vectors out) time of bench.py goes.

Same workload and page batches as bench.py's e2e leg, three timed variants on the public API:
  h2d_only   every batch: open (H2D + index kernel) -> close            (no result traffic)
  d2h_only   batches opened beforehand and resident: scan -> fetch       (no input traffic)
  full       open -> scan -> fetch, with per-phase host timestamps: how much of the wall time had an
             H2D in flight, a D2H in flight, both, or neither
One JSON line; the link limits to hold it against come from tools/pcie_probe.py."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def union_ms(iv):
    iv = sorted(iv)
    tot, cur_a, cur_b = 0.0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        tot += cur_b - cur_a
    return tot * 1e3


def overlap_ms(a, b):
    # total time covered by both unions (sweep over the merged boundaries)
    pts = sorted({x for iv in a + b for x in iv})
    tot = 0.0
    for lo, hi in zip(pts[:-1], pts[1:]):
        mid = (lo + hi) / 2
        if any(x <= mid < y for x, y in a) and any(x <= mid < y for x, y in b):
            tot += hi - lo
    return tot * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--batches", type=int, default=12)
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import oceanbase_b200 as ob
    from oceanbase_b200.pipeline import split_table

    bench.bind_to_gpu_numa_node(0)
    w, _pin = bench.build_workload(args.rows, 0, 1234, pinned=True)
    table = w.table
    bpb = max(1, table.n_blocks // args.batches)
    parts = split_table(table, bpb)
    out_np, null_np, keep = [], [], []
    for part in parts:
        capp = int(int(part.n_blocks) * 1400 * 0.30) + 2048
        bufs = [torch.empty(capp, dtype=torch.int64, pin_memory=True) for _ in w.proj]
        nbufs = [torch.zeros((capp + 63) // 64, dtype=torch.int64, pin_memory=True) for _ in w.proj]
        keep.append(bufs + nbufs)
        out_np.append([t.numpy().view(np.uint64) for t in bufs])
        null_np.append([t.numpy().view(np.uint64) for t in nbufs])
    # the round-1 Python pipeline's shape rebuilt here on purpose: the phases of obgpu_pipeline_scan (open / scan / fetch per page batch,
    # one ctx = one stream per worker) timed separately
    class _Pipe:
        ctxs = [ob.ScanContext(0) for _ in range(args.workers)]
    pipe = _Pipe()

    def run(kind, resident=None, log=None):
        lock = threading.Lock()
        nxt = [0]
        sel = [0] * len(parts)

        def worker(ctx):
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(parts):
                    return
                t0 = time.perf_counter()
                batch = resident[ctx][i] if resident is not None and i in resident[ctx] else None
                own = batch is None
                if own:
                    batch = ctx.open_batch(parts[i])
                t1 = time.perf_counter()
                if kind != "h2d_only":
                    res = batch.scan(w.filter, w.proj, max_selected_rows=int(batch.total_rows * 0.30) + 1024)
                    n = res.selected_rows
                    t2 = time.perf_counter()
                    res.fetch_cols(list(range(len(w.proj))), 0, n, outs=out_np[i], out_nulls=null_np[i])
                    t3 = time.perf_counter()
                    sel[i] = n
                    res.free()
                    if log is not None:
                        log.append((i, t0, t1, t2, t3))
                if own:
                    batch.close()

        ths = [threading.Thread(target=worker, args=(c,)) for c in pipe.ctxs]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, sum(sel), t0

    out = {"rows": table.total_rows, "h2d_bytes": int(table.image.size), "batches": len(parts), "workers": args.workers}
    run("full")
    run("full")
    out["h2d_only_ms"] = min(run("h2d_only")[0] for _ in range(args.reps))
    # resident batches: worker k owns batches k, k+W, ... (a batch belongs to the ctx that opened it)
    resident = {c: {} for c in pipe.ctxs}
    for i, part in enumerate(parts):
        c = pipe.ctxs[i % len(pipe.ctxs)]
        resident[c][i] = c.open_batch(part)

    def run_resident():
        # static assignment so that every batch is scanned on its own ctx
        sel = [0] * len(parts)

        def worker(ctx):
            for i, batch in resident[ctx].items():
                res = batch.scan(w.filter, w.proj, max_selected_rows=int(batch.total_rows * 0.30) + 1024)
                n = res.selected_rows
                res.fetch_cols(list(range(len(w.proj))), 0, n, outs=out_np[i], out_nulls=null_np[i])
                sel[i] = n
                res.free()

        ths = [threading.Thread(target=worker, args=(c,)) for c in pipe.ctxs]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, sum(sel)

    run_resident()
    r = [run_resident() for _ in range(args.reps)]
    out["d2h_only_ms"] = min(x[0] for x in r)
    out["d2h_bytes"] = int(r[0][1]) * 8 * len(w.proj)
    for c in pipe.ctxs:
        for b in resident[c].values():
            b.close()
    best = None
    for _ in range(args.reps):
        log = []
        ms, n, t0 = run("full", log=log)
        if best is None or ms < best[0]:
            best = (ms, n, t0, log)
    ms, n, t0, log = best
    h2d = [(a, b) for _, a, b, _, _ in log]
    d2h = [(c, d) for _, _, _, c, d in log]
    out["full_ms"] = ms
    out["full_selected"] = n
    out["full_h2d_busy_ms"] = union_ms(h2d)
    out["full_d2h_busy_ms"] = union_ms(d2h)
    out["full_both_busy_ms"] = overlap_ms(h2d, d2h)
    out["full_open_ms_mean"] = float(np.mean([b - a for a, b in h2d]) * 1e3)
    out["full_scan_ms_mean"] = float(np.mean([c - b for _, _, b, c, _ in log]) * 1e3)
    out["full_fetch_ms_mean"] = float(np.mean([d - c for c, d in d2h]) * 1e3)
    out["first_d2h_start_ms"] = (min(c for c, _ in d2h) - t0) * 1e3
    out["last_h2d_end_ms"] = (max(b for _, b in h2d) - t0) * 1e3
    for c in pipe.ctxs:
        c.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
