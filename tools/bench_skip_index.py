"""Skip-index pruning on the config-2 table: a rowkey range predicate (BETWEEN on the sorted PK, --frac of the
rows) scanned with and without the micro-blocks' aggregate rows attached. Device resident, all 8 columns
projected. Reports the blocks the index decided, the kernel time of both variants (CUDA events on the launching
stream) and checks that both return the same rows. One JSON line; not the round's bench line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--frac", type=float, default=0.25)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    import oceanbase_b200 as ob
    from oceanbase_b200.synth import config2_pk
    seed = 1234
    w, _ = bench.build_workload(args.rows, 0, seed)
    table = w.table
    rpb = w.rows_per_block
    t0 = time.perf_counter()
    pk = config2_pk(args.rows, seed)
    # build_workload encodes the table in chunks of 4 M rows (each ends with a short block): same blocking here
    chunk = 4_000_000
    parts, offs, pos = [], [np.zeros(1, dtype=np.int64)], 0
    for s0 in range(0, args.rows, chunk):
        r, o = ob.table_agg_rows([ob.Column(ob.OBJ_INT, ob.ENC_RAW, pk[s0:s0 + chunk])], [0], rpb)
        parts.append(r)
        offs.append(o[1:] + pos)
        pos += int(o[-1])
    agg_rows, agg_off = np.concatenate(parts), np.concatenate(offs)
    assert len(agg_off) == table.n_blocks + 1
    t_agg = time.perf_counter() - t0
    lo_row = int(args.rows * (0.5 - args.frac / 2))
    hi_row = int(args.rows * (0.5 + args.frac / 2)) - 1
    flt = ob.White(0, ob.WHITE_OP_BT, (int(pk[lo_row]), int(pk[hi_row])))
    want = hi_row - lo_row + 1
    del pk
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = ob.ScanContext(0, stream=stream.cuda_stream)
    ctx.set_profiling(True)
    d_image = torch.empty(table.image.size + 64, dtype=torch.uint8, device=dev)
    d_image[:table.image.size].copy_(torch.from_numpy(table.image))
    d_image[table.image.size:].zero_()
    torch.cuda.synchronize()
    batch = ctx.open_batch(table, device_image_ptr=d_image.data_ptr())
    cap = want + 4096

    def run(label):
        for _ in range(args.warmup):
            r = batch.scan(flt, w.proj, max_selected_rows=cap)
            r.info()
            r.free()
        torch.cuda.synchronize()
        last = None
        for _ in range(args.steps):
            if last is not None:
                last.free()
            last = batch.scan(flt, w.proj, max_selected_rows=cap)
        torch.cuda.synchronize()
        sel = last.selected_rows
        skipped = last.skip_info()
        data, _, _ = last.fetch_col(0)
        chk = int(np.bitwise_xor.reduce(data.view(np.uint64)))
        last.free()
        return {"kernel_ms": float(np.mean(ctx.kernel_times_ms(args.steps))), "selected": sel, "always_false_blocks": skipped[0],
                "always_true_blocks": skipped[1], "pk_xor": chk}

    plain = run("plain")
    batch.set_agg_rows(agg_rows, agg_off)
    pruned = run("pruned")
    assert plain["selected"] == pruned["selected"] == want and plain["pk_xor"] == pruned["pk_xor"]
    peak, src = bench.measured_peak_gbs()
    print(json.dumps({"workload": f"cfg2 table, BETWEEN on the sorted PK ({args.frac:.0%} of the rows), 8 columns projected",
                      "rows": table.total_rows, "micro_blocks": table.n_blocks, "agg_row_bytes": int(agg_rows.size),
                      "agg_rows_build_s": round(t_agg, 2), "without_skip_index": plain, "with_skip_index": pruned,
                      "speedup": plain["kernel_ms"] / pruned["kernel_ms"],
                      "rows_per_s_with": table.total_rows / (pruned["kernel_ms"] * 1e-3), "hbm_peak_gbs": peak}))
    batch.close()
    ctx.close()


if __name__ == "__main__":
    main()
