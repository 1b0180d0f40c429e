"""Regression anchors (tests/golden/writer_blocks.npz): one micro-block per codec family exactly as the writer emitted it when the
fixture was made -- each accepted by the real reference's ObMicroBlockHeader::deserialize_and_check_record at that time. The oracle
must decode the committed bytes to the recorded cells, the writer must still produce the same bytes from the same inputs, and on the
GPU box the device must read the same cells out of the committed bytes (no /root/reference, no oracle/_ref needed)."""
import os

import numpy as np
import pytest

import oracle_binding as ora

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "writer_blocks.npz"))
NAMES = [str(x) for x in G["names"]]


def recorded_cells(name, c):
    nulls = G[f"null_{name}_{c}"]
    if f"int_{name}_{c}" in G.files:
        return [None if nulls[i] else int(v) for i, v in enumerate(G[f"int_{name}_{c}"])], False
    heap, lens = G[f"str_{name}_{c}"].tobytes(), G[f"len_{name}_{c}"]
    out, at = [], 0
    for i, ln in enumerate(lens):
        out.append(None if nulls[i] else heap[at:at + int(ln)])
        at += int(ln)
    return out, True


@pytest.mark.parametrize("name", NAMES)
def test_oracle_decodes_the_committed_bytes(name):
    b = np.ascontiguousarray(G[f"block_{name}"])
    blk = ora.Block(b)
    assert blk.verify_checksums() == 0
    rows = G[f"rows_{name}"]
    for c in range(blk.column_count):
        want, is_str = recorded_cells(name, c)
        for i, r in enumerate(rows):
            cell = blk.cell(c, int(r))
            if want[i] is None:
                assert cell is None, (name, c, r)
            elif is_str:
                assert bytes(cell) == want[i], (name, c, r)
            else:
                assert (cell & 0xffffffffffffffff) == want[i], (name, c, r)


def test_writer_still_emits_the_committed_bytes():
    import test_micro_header_ref_kat as mk
    seen = 0
    for name, b in mk.blocks():
        assert np.array_equal(np.ascontiguousarray(b), G[f"block_{name}"]), name
        seen += 1
    assert seen == len(NAMES)


# the shapes whose device read-out path (scan -> fetch_col / fetch_strings) the other GPU suites exercise daily
DEVICE_NAMES = [n for n in NAMES if n in ("pax_dict_rle_const", "pax_base_diff_rowkey", "pax_auto", "one_row")]


@pytest.mark.gpu
@pytest.mark.parametrize("name", DEVICE_NAMES)
def test_device_reads_the_committed_bytes(name):
    import oceanbase_b200 as ob
    from oceanbase_b200.sstable import TableImage
    b = np.ascontiguousarray(G[f"block_{name}"])
    blk = ora.Block(b)
    image = np.concatenate([b, np.zeros((-b.size) % 128 + 128, dtype=np.uint8)])
    table = TableImage(image, np.array([0], dtype=np.int64), np.array([b.size], dtype=np.int64), blk.row_count, blk.column_count)
    ctx = ob.ScanContext(0)
    try:
        batch = ob.PageBatch(ctx, table)
    except ob.ObGpuError as e:   # codecs the device leaves to the host report OB_NOT_SUPPORTED at open
        assert e.code == ob.OB_NOT_SUPPORTED, name
        ctx.close()
        return
    rows = G[f"rows_{name}"]
    try:
        res = batch.scan(None, list(range(blk.column_count)))
    except ob.ObGpuError as e:
        assert e.code == ob.OB_NOT_SUPPORTED, name
        batch.close()
        ctx.close()
        return
    for c in range(blk.column_count):
        want, is_str = recorded_cells(name, c)
        if is_str:
            heap, off = res.fetch_strings(c)
            for i, r in enumerate(rows):
                got = bytes(heap[off[r]:off[r + 1]])
                assert got == (want[i] or b""), (name, c, r)
        else:
            data, _, nulls = res.fetch_col(c)
            el = data.dtype.itemsize
            for i, r in enumerate(rows):
                is_null = (int(nulls[r // 64]) >> (r % 64)) & 1
                if want[i] is None:
                    assert is_null, (name, c, r)
                else:
                    assert not is_null and int(data[r]) == (want[i] & ((1 << (8 * el)) - 1)), (name, c, r)
    batch.close()
    ctx.close()
