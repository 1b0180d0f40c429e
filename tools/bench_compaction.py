"""Config 5 (BASELINE.json configs[4], scaled stand-in): K-way major-compaction merge of K sorted runs with
overlapping INT64 rowkey ranges. One process per GPU; run q lives on rank q % world; ranks range-partition
the rowkey space (sample -> all_gather -> splitters), exchange run slices (NCCL P2P over NVLink), merge
locally on the device. Prints one JSON line: input rows/s over the whole job (max over ranks), the phase
split and a parity check (row counts / checksums of the merged stream against the oracle at --verify size)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=8)
    ap.add_argument("--window", type=int, default=4_000_000, help="rowkey indexes covered by one run")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--verify", action="store_true", help="compare the merged stream with the oracle (small sizes)")
    ap.add_argument("--python-exchange", action="store_true", help="round-1 path: torch.distributed exchange driven from Python")
    ap.add_argument("--stream-ranges", type=int, default=0, help="N > 0: host-resident runs merged range by range (runs larger than HBM): "
                                                                 "every step copies the blocks in, merges and fetches the rows out")
    args = ap.parse_args()
    return run(args)


def run(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    rank, world, local = bench.env_int("RANK", 0), bench.env_int("WORLD_SIZE", 1), bench.env_int("LOCAL_RANK", 0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    import oceanbase_b200 as ob
    from oceanbase_b200.synth import make_config5_runs
    from oceanbase_b200.compaction import decode_run, merge_decoded, distributed_major_merge, merge_decoded_distributed, Comm, encode_merge_result
    from oceanbase_b200 import capi
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = ob.ScanContext(local, stream=stream.cuda_stream)
    t0 = time.perf_counter()
    mine = [q for q in range(args.runs) if q % world == rank]
    runs = make_config5_runs(n_runs=args.runs, window=args.window, seed=5, n_threads=max(1, bench.host_cpus() // world),
                             encode=True, only=None if (args.verify or world == 1) else set(mine))
    t_gen = time.perf_counter() - t0
    if getattr(args, "stream_ranges", 0) > 0:
        return run_streamed(args, runs, t_gen, dev, local)
    # device-resident SSTables of the local runs
    images = {}
    for q in mine:
        tb = runs[q]["table"]
        d = torch.empty(tb.image.size + 64, dtype=torch.uint8, device=dev)
        d[:tb.image.size].copy_(torch.from_numpy(tb.image))
        d[tb.image.size:].zero_()
        images[q] = d
    torch.cuda.synchronize()
    # page batches stay open across steps (the block cache); the communicator is made once
    batches = {q: ctx.open_batch(runs[q]["table"], device_image_ptr=images[q].data_ptr()) for q in mine}
    comm = None
    if world > 1 and not getattr(args, "python_exchange", False):
        comm = Comm.from_torch_distributed(ctx, dev)
    in_rows_local = sum(len(runs[q]["key"]) for q in mine)
    enc_bytes_local = sum(int(runs[q]["table"].sizes.sum()) for q in mine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ENC_ROWS_PER_BLOCK = 500
    phase_b = {}

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        dec = {q: decode_run(ctx, runs[q]["table"], 0, 1, [2, 3, 4], device=dev, batch=batches[q], check_rowkey=False) for q in mine}
        ev[1].record(stream)
        if comm is not None:
            res, _, recv_rows = merge_decoded_distributed(ctx, comm, dec, args.runs, 3)
        elif world > 1:
            res, _, recv_rows = distributed_major_merge(dec, args.runs, 3, lambda rs: merge_decoded(ctx, rs))
        else:
            res = merge_decoded(ctx, [dec[q] for q in range(args.runs)])
            recv_rows = None
        info = res.info()
        ev[2].record(stream)
        # phase B: the merged rows leave the device as SSTable bytes (PAX blocks, every column RAW) + column checksums
        enc = encode_merge_result(res, [-1, 0, 1, 2], [capi.OBJ_INT] * 4, ENC_ROWS_PER_BLOCK, rowkey_cnt=1) if info.out_rows else None
        ev[3].record(stream)
        torch.cuda.synchronize()
        phase_b["image_bytes"] = enc.info().image_size if enc else 0
        phase_b["n_blocks"] = enc.info().n_blocks if enc else 0
        phase_b["host_blocks"] = enc.info().n_host_blocks if enc else 0
        phase_b["checksums"] = [int(x) for x in enc.column_checksums()] if enc else [0] * 4
        if enc:
            enc.free()
        return res, info, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])

    for _ in range(args.warmup):
        r = step()[0]
        r.free()
    barrier()
    t0 = time.perf_counter()
    dec_ms, mrg_ms, enc_ms = [], [], []
    for _ in range(args.steps):
        res, info, a, b, c_ms = step()
        dec_ms.append(a)
        mrg_ms.append(b)
        enc_ms.append(c_ms)
        if _ + 1 < args.steps:
            res.free()
    barrier()
    step_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    out_rows = info.out_rows
    ksum = int(res.fetch(-1)[0].astype(np.uint64).sum()) if out_rows else 0
    vsum = [int(res.fetch(c)[0].astype(np.uint64).sum()) for c in range(3)]
    nsum = [int(res.fetch(c)[1].sum()) for c in range(3)]
    t = torch.tensor([step_ms, float(np.mean(dec_ms)), float(np.mean(mrg_ms)), float(np.mean(enc_ms))], device=dev, dtype=torch.float64)
    pb = torch.tensor([phase_b["image_bytes"], phase_b["n_blocks"], phase_b["host_blocks"]] + [c % (1 << 56) for c in phase_b["checksums"]],
                      device=dev, dtype=torch.int64)
    tot = torch.tensor([in_rows_local, out_rows, info.dropped_deletes, info.fused_rows, enc_bytes_local, ksum % (1 << 62)]
                       + [v % (1 << 62) for v in vsum] + nsum, device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(pb, op=dist.ReduceOp.SUM)
    step_ms, dec_mean, mrg_mean, enc_mean = t.tolist()
    pb = pb.tolist()
    tot = tot.tolist()
    if rank == 0:
        line = {"metric": "major-compaction merged input rows/sec", "value": tot[0] / (step_ms * 1e-3), "unit": "rows/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "scaling": "strong",
                "config": {"workload": f"cfg5 stand-in: {args.runs} sorted runs, window {args.window}, 50 % range overlap, "
                                       f"10 % duplicated rowkeys (NOP cells), 2 % deletes; INT64 rowkey + 3 INT64 payload",
                           "input_rows": tot[0], "output_rows": tot[1], "dropped_deletes": tot[2], "fused_rows": tot[3],
                           "encoded_bytes": tot[4], "gen_seconds": round(t_gen, 1)},
                "higher_is_better": True, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "phases_ms": {"decode_runs": dec_mean, "exchange_plus_merge": mrg_mean,
                              "exchange": "ncclAllGather + grouped ncclSend / ncclRecv inside libobgpu_scan.so" if comm is not None else
                                          ("torch.distributed from Python" if world > 1 else "none (one rank)"),
                              "timing": "wall clock per step incl. host orchestration; phases by CUDA events, max over ranks"}}
        peak, peak_src = bench.measured_peak_gbs()
        # phase B (device encoder): merged columns read once (8 B x 4 columns + 3 NULL bytes per row), image written once
        enc_alg = 35 * tot[1] + pb[0]
        line["phase_b"] = {"what": f"merged rows -> PAX micro-blocks of {ENC_ROWS_PER_BLOCK} rows (rowkey + 3 payload columns, every column RAW) "
                                   "+ column checksums, on the device (obgpu_merge_result_encode); inside ms_per_step",
                           "encode_ms": enc_mean, "image_bytes": pb[0], "n_blocks": pb[1], "host_encoded_blocks": pb[2],
                           "alg_gbps_per_gpu": enc_alg / (enc_mean * 1e-3) / 1e9 / world if enc_mean > 0 else None,
                           "frac": enc_alg / (enc_mean * 1e-3) / 1e9 / world / peak if enc_mean > 0 else None,
                           "d2h_bytes_as_blocks": pb[0], "d2h_bytes_as_rows": 35 * tot[1]}
        # bytes the merge has to move per step (SURVEY 8d style): encoded blocks read once, decoded cells (36 B / row: rowkey,
        # flag, 3 payload values + ext bytes) written by the decode and read by the merge, merged rows written (35 B / row)
        alg = tot[4] + 2 * 36 * tot[0] + 35 * tot[1]
        kern_ms = dec_mean + mrg_mean
        line["roofline"] = {"bound": "hbm", "achieved": alg / (kern_ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                            "frac": alg / (kern_ms * 1e-3) / 1e9 / world / peak, "alg_bytes_per_step": alg, "kernel_ms": kern_ms,
                            "peak_source": peak_src, "note": "per GPU: whole-job bytes / (N x max-over-ranks device time)"}
        if args.verify:
            import oracle_binding as ora
            want = ora.major_merge(runs, 3)
            M = 1 << 62
            u = lambda a: int(a.astype(np.uint64).sum()) % M
            sums_ok = tot[5] % M == u(want["key"]) and all(tot[6 + c] % M == u(want["vals"][c]) for c in range(3))
            o = ora.oracle()
            M56 = 1 << 56   # column checksums add up over the ranks' ranges (wrapping int64 sums): compared modulo 2^56
            want_ck = [o.ora_column_checksum(want["key"].ctypes.data, None, len(want["key"]), 8) % M56] + \
                      [o.ora_column_checksum(np.ascontiguousarray(want["vals"][c]).ctypes.data, np.ascontiguousarray(want["null"][c]).ctypes.data,
                                             len(want["key"]), 8) % M56 for c in range(3)]
            line["phase_b"]["column_checksums_match"] = [c % M56 for c in pb[3:7]] == want_ck
            line["parity"] = {"rows_match": tot[1] == len(want["key"]), "dropped_match": tot[2] == want["dropped"],
                              "fused_match": tot[3] == want["fused"], "null_counts_match": tot[9:12] == [int(x.sum()) for x in want["null"]],
                              "value_checksums_match": sums_ok}
        print(json.dumps(line))
    res.free()
    for b in batches.values():
        b.close()
    if comm is not None:
        comm.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_streamed(args, runs, t_gen, dev, local):
    """Host in, host out: the runs' SSTables sit in pinned host memory; every step merges them range by range (two ranges in flight)
    and fetches the merged rows into host buffers."""
    import torch
    import oceanbase_b200 as ob
    from oceanbase_b200 import capi
    from oceanbase_b200.compaction import streamed_major_merge
    rpb = 1400
    tables, end_keys = [], []
    for r in runs:
        tb = r["table"]
        pin = torch.empty(tb.image.size, dtype=torch.uint8).pin_memory()
        img = pin.numpy()
        img[:] = tb.image
        tables.append(ob.TableImage(img, tb.offsets, tb.sizes, tb.total_rows, tb.n_cols))
        k = r["key"]
        end_keys.append(k[np.minimum(np.arange(rpb - 1, len(k) + rpb - 1, rpb), len(k) - 1)])
        assert len(end_keys[-1]) == tb.n_blocks
    in_rows = sum(len(r["key"]) for r in runs)
    enc = sum(int(t.sizes.sum()) for t in tables)
    out_cap = in_rows
    host_out = [torch.empty(out_cap, dtype=torch.int64).pin_memory().numpy() for _ in range(4)]
    host_nl = [torch.empty(out_cap, dtype=torch.uint8).pin_memory().numpy() for _ in range(4)]
    state = {"rows": 0, "dropped": 0, "fused": 0}

    def sink(i, res):
        info = res.info()
        n = info.out_rows
        at = state["rows"]
        for k, c in enumerate((-1, 0, 1, 2)):   # straight into the pinned output buffers (no pageable staging)
            if n > 0:
                capi.check(capi.lib.obgpu_merge_result_fetch(res._h, c, 0, n, host_out[k][at:].ctypes.data, host_nl[k][at:].ctypes.data),
                           "obgpu_merge_result_fetch", res.ctx._h)
        state["rows"] += n
        state["dropped"] += info.dropped_deletes
        state["fused"] += info.fused_rows

    from oceanbase_b200.compaction import merge_runs_streamed

    def step():
        state.update(rows=0, dropped=0, fused=0)
        if getattr(args, "stream_python", False):   # the Python orchestration of the same loop
            streamed_major_merge(tables, end_keys, 0, 1, [2, 3, 4], args.stream_ranges, sink, device=dev)
        else:
            merge_runs_streamed(local, tables, end_keys, 0, 1, [2, 3, 4], args.stream_ranges, sink, n_streams=3)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    line = {"metric": "major-compaction merged input rows/sec", "value": in_rows / (ms * 1e-3), "unit": "rows/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "strong", "higher_is_better": True,
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": f"cfg5 stand-in, STREAMED: {args.runs} host-resident runs, window {args.window}, merged in "
                                   f"{args.stream_ranges} rowkey ranges (obgpu_merge_runs_streamed, 3 in flight), rows fetched back to pinned host buffers",
                       "input_rows": in_rows, "output_rows": state["rows"], "dropped_deletes": state["dropped"], "fused_rows": state["fused"],
                       "encoded_bytes": enc, "gen_seconds": round(t_gen, 1)},
            "e2e": {"value": in_rows / (ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": int(enc), "d2h_bytes_per_step": int(state["rows"] * 36),
                    "note": "boundary blocks of a range are copied by both neighbours: h2d is a lower bound"}}
    if args.verify:
        import oracle_binding as ora
        want = ora.major_merge(runs, 3)
        n = state["rows"]
        ok = n == len(want["key"]) and np.array_equal(host_out[0][:n], want["key"])
        for c in range(3):
            ok = ok and np.array_equal(host_nl[c + 1][:n] != 0, want["null"][c] != 0)
            ok = ok and np.array_equal(host_out[c + 1][:n][host_nl[c + 1][:n] == 0], want["vals"][c][want["null"][c] == 0])
        line["parity"] = {"rows_and_cells_match": bool(ok), "dropped_match": state["dropped"] == want["dropped"], "fused_match": state["fused"] == want["fused"]}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    main()
