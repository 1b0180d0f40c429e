"""Host-buffer scan pipeline: H2D of page batch i+1, kernels of batch i and D2H of batch i-1 overlap.

A table that lives in host memory (the block cache) is cut into page batches of consecutive
micro-blocks; `n_workers` worker threads, each with its own obgpu_ctx (= its own CUDA stream), pull
batches from a queue and run open (H2D + index) -> scan -> fetch (D2H) through the public C-ABI.
ctypes releases the GIL inside every C call, so the copies of different batches run concurrently in
both PCIe directions while the kernels (sub-millisecond) slot in between. Results are delivered per
batch in block order (dense inside a batch), which is how a block-at-a-time consumer such as
ObSSTableRowScanner drains them.
"""
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .scan import ScanContext
from .sstable import TableImage


@dataclass
class BatchOutput:
    block_begin: int
    block_end: int
    total_rows: int
    selected_rows: int
    cols: List[np.ndarray]           # per projected column: payload (uint64 / uint32 / uint8)
    lens: List[Optional[np.ndarray]]  # string columns: int32 lens
    nulls: List[np.ndarray]
    has_null: List[int]
    image_lo: int = 0                # byte offset of the batch's first block inside the caller's table image


def batch_bounds(n_blocks: int, blocks_per_batch: int, ramp: int = 0) -> List[int]:
    """Block index where every page batch starts (+ n_blocks at the end). ramp > 0: the first `ramp`
    batches are 1/2^ramp, ..., 1/2 of a full batch, so that the first results start flowing back (D2H)
    while most of the input is still on its way in."""
    bounds, b0 = [0], 0
    for k in range(ramp, 0, -1):
        step = max(1, blocks_per_batch >> k)
        if b0 + step >= n_blocks:
            break
        b0 += step
        bounds.append(b0)
    while b0 < n_blocks:
        b0 = min(n_blocks, b0 + blocks_per_batch)
        bounds.append(b0)
    return bounds


def split_table(table: TableImage, blocks_per_batch: int, ramp: int = 0) -> List[TableImage]:
    parts = []
    n = table.n_blocks
    bounds = batch_bounds(n, blocks_per_batch, ramp)
    for b0, b1 in zip(bounds[:-1], bounds[1:]):
        lo = int(table.offsets[b0])
        hi = int(table.offsets[b1]) if b1 < n else int(table.image.size)
        part = TableImage(table.image[lo:hi], table.offsets[b0:b1] - lo, table.sizes[b0:b1], 0, table.n_cols)
        part.image_lo = lo   # byte offset of the part inside the caller's table image (string pointers are rebased by it)
        parts.append(part)
    return parts


class HostScanPipeline:
    """Reusable pipeline (contexts, streams and pinned output buffers are created once)."""

    def __init__(self, device: int, n_workers: int = 3):
        self.device = device
        self.ctxs = [ScanContext(device) for _ in range(n_workers)]

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def scan(self, table: TableImage, filter, proj: Sequence[int], blocks_per_batch: int, selectivity_hint: float,
             out_buffers: Optional[List[List[np.ndarray]]] = None, string_base: int = 0,
             null_buffers: Optional[List[List[np.ndarray]]] = None, ramp: int = 0,
             agg_rows: Optional[np.ndarray] = None, agg_off: Optional[np.ndarray] = None) -> List[BatchOutput]:
        """agg_rows / agg_off: the micro-blocks' serialized aggregate rows (block b: agg_rows[agg_off[b]:agg_off[b + 1]]);
        every page batch then carries its slice and the scan prunes with the skip index."""
        parts = split_table(table, blocks_per_batch, ramp)
        bounds = batch_bounds(table.n_blocks, blocks_per_batch, ramp)
        if agg_rows is not None:
            agg_rows = np.ascontiguousarray(agg_rows, dtype=np.uint8)
            agg_off = np.ascontiguousarray(agg_off, dtype=np.int64)
            assert len(agg_off) == table.n_blocks + 1
        outs: List[Optional[BatchOutput]] = [None] * len(parts)
        errors = []
        lock = threading.Lock()
        next_idx = [0]

        def worker(ctx: ScanContext):
            try:
                while True:
                    with lock:
                        i = next_idx[0]
                        next_idx[0] += 1
                    if i >= len(parts):
                        return
                    part = parts[i]
                    batch = ctx.open_batch(part)                      # H2D + index kernel
                    if agg_rows is not None:                           # offsets keep their table-wide base: no copy
                        batch.set_agg_rows(agg_rows, agg_off[bounds[i]:bounds[i + 1] + 1])
                    cap = int(batch.total_rows * selectivity_hint) + 1024
                    # VEC_DISCRETE pointers must address the CALLER's table buffer: the part's block offsets were rebased by
                    # image_lo, so the base moves up by the same amount
                    sbase = string_base + part.image_lo
                    res = batch.scan(filter, proj, string_base=sbase, max_selected_rows=min(cap, batch.total_rows))
                    try:
                        n = res.selected_rows                         # sync + status
                    except Exception as e:                             # capacity overflow: exact re-run
                        from .capi import ObGpuError, OB_BUF_NOT_ENOUGH
                        if isinstance(e, ObGpuError) and e.code == OB_BUF_NOT_ENOUGH:
                            need = res._info.selected_rows
                            res.free()
                            res = batch.scan(filter, proj, string_base=sbase, max_selected_rows=need)
                            n = res.selected_rows
                        else:
                            raise
                    cols, lens, nulls, hn = [], [], [], []
                    is_str = [res.col(c).is_string for c in range(len(proj))]
                    if not any(is_str):
                        # one synchronisation for the whole batch: every column's D2H is enqueued first
                        cols, nulls = res.fetch_cols(list(range(len(proj))), 0, n,
                                                     outs=out_buffers[i] if out_buffers is not None else None,
                                                     out_nulls=null_buffers[i] if null_buffers is not None else None)
                        lens = [None] * len(proj)
                        hn = [res.col(c).has_null for c in range(len(proj))]
                    else:
                        for c in range(len(proj)):
                            ob = out_buffers[i][c] if out_buffers is not None else None
                            d, l, nl = res.fetch_col(c, 0, n, out=ob)     # D2H
                            cols.append(d)
                            lens.append(l)
                            nulls.append(nl)
                            hn.append(res.col(c).has_null)
                    outs[i] = BatchOutput(bounds[i], bounds[i + 1], batch.total_rows,
                                          n, cols, lens, nulls, hn, part.image_lo)
                    res.free()
                    batch.close()
            except Exception as e:  # pragma: no cover
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(c,)) for c in self.ctxs]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return outs
