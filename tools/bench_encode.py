"""Device encoder (phase B of the compaction) alone: config-5 shaped columns (INT64 rowkey + 3 INT64 payload columns, one of
them with NULLs) already in HBM -> PAX micro-blocks + column checksums. Prints one JSON line: rows/s, algorithmic GB/s
(input columns read once + image written once) against the measured HBM peak."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=64_000_000)
    ap.add_argument("--rows-per-block", type=int, default=500)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import torch
    import oceanbase_b200 as ob
    from oceanbase_b200 import capi, compaction
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ob.ScanContext(0, stream=stream.cuda_stream)   # CUDA events below time the stream the kernels run on
    g = torch.Generator(device="cuda").manual_seed(5)
    n = a.rows
    key = torch.arange(n, device="cuda", dtype=torch.int64) * 3 + 1_000_000_007
    c1 = torch.randint(0, 1 << 33, (n,), device="cuda", dtype=torch.int64, generator=g)
    c2 = torch.randint(-(1 << 62), 1 << 62, (n,), device="cuda", dtype=torch.int64, generator=g)
    c3 = torch.randint(0, 1 << 13, (n,), device="cuda", dtype=torch.int64, generator=g)
    n3 = (torch.rand((n,), device="cuda", generator=g) < 0.05).to(torch.uint8)
    cols = [(key.data_ptr(), None, capi.OBJ_INT, False), (c1.data_ptr(), None, capi.OBJ_INT, False),
            (c2.data_ptr(), None, capi.OBJ_INT, False), (c3.data_ptr(), n3.data_ptr(), capi.OBJ_INT, False)]
    in_bytes = n * (4 * 8 + 1)
    ms, img_bytes, nb = [], 0, 0
    for it in range(a.warmup + a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        enc = compaction.encode_columns(ctx, cols, n, a.rows_per_block, rowkey_cnt=1)
        e1.record()
        torch.cuda.synchronize()
        info = enc.info()
        img_bytes, nb = info.image_size, info.n_blocks
        assert info.n_host_blocks == 0
        enc.free()
        if it >= a.warmup:
            ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms))
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    peak = float(peaks.get("hbm_gbps_burst", peaks.get("hbm_gbps", 6570.3))) if isinstance(peaks, dict) else 6570.3
    alg = in_bytes + img_bytes
    print(json.dumps({"workload": "device encoder, cfg5 columns", "rows": n, "rows_per_block": a.rows_per_block, "n_blocks": nb,
                      "ms": round(t, 3), "rows_per_s": n / t * 1e3, "in_bytes": in_bytes, "image_bytes": img_bytes,
                      "alg_gbps": round(alg / t / 1e6, 1), "peak_gbps": peak, "frac": round(alg / t / 1e6 / peak, 3),
                      "note": "event-timed around obgpu_encode_columns (includes its allocation + the 64 KB table upload)"}))


if __name__ == "__main__":
    main()
