"""CS blocks whose integer streams carry the non-RAW codecs (K13): writer (stream encoders pinned byte-exact to the
reference's, tests/test_stream_codec_kat.py) -> oracle transform (decoders pinned to the reference's) -> the same
payload the RAW writer produces, cell for cell and byte for byte."""
import ctypes as C

import numpy as np
import pytest

import oceanbase_b200 as ob
import oracle_binding as ora
from oceanbase_b200.capi import lib

MODES = {0: "detect", 2: "DOUBLE_DELTA_ZIGZAG_RLE", 3: "DOUBLE_DELTA_ZIGZAG_PFOR", 4: "DELTA_ZIGZAG_RLE", 5: "DELTA_ZIGZAG_PFOR",
         6: "SIMD_FIXEDPFOR", 8: "XOR_FIXED_PFOR"}


def make_cols(n, seed):
    rng = np.random.default_rng(seed)
    strs = [bytes(rng.integers(97, 123, size=rng.integers(1, 18), dtype=np.uint8)) for _ in range(60)]
    nulls = (rng.random(n) < 0.1).astype(np.uint8)
    return [
        ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.arange(n, dtype=np.int64) * 7 + 1_000_000),                       # monotone
        ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, rng.integers(-500, 500, size=n)),                                      # base + small
        ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, np.repeat(rng.integers(0, 1 << 40, size=n // 50 + 1), 50)[:n], nulls=nulls),  # runs + NULLs
        ob.Column(ob.OBJ_DATE, ob.ENC_CS_INTEGER, rng.integers(8036, 10562, size=n)),
        ob.Column(ob.OBJ_INT, ob.ENC_CS_INT_DICT, rng.integers(0, 40, size=n) * 1_000_003, nulls=nulls),
        ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STRING, [strs[i] for i in rng.integers(0, 60, size=n)]),
        ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STR_DICT, [strs[i] for i in rng.integers(0, 60, size=n)], nulls=nulls),
        ob.Column(ob.OBJ_UINT64, ob.ENC_CS_INTEGER, (rng.integers(0, 1 << 62, size=n).astype(np.uint64) * np.uint64(3)).view(np.int64)),
    ]


def stream_types(block):
    """ObIntegerStream::EncodingType of every integer stream of a CS block, in stream order (string streams: None)."""
    hs = int(block[4:8].view(np.uint32)[0])
    ncol = int(block[10:12].view(np.uint16)[0])
    rows = int(block[16:20].view(np.uint32)[0])
    ah = block[hs:hs + 12]
    offsets_len = int(ah[6:10].view(np.uint32)[0])
    so = block[len(block) - offsets_len:]
    assert so[2] == 1                      # the stream-offsets stream itself stays RAW in this writer
    ow = 1 << int(so[3])
    n_streams = int(ah[10:12].view(np.uint16)[0])
    ends = so[5:5 + ow * n_streams].view({1: np.uint8, 2: np.uint16, 4: np.uint32}[ow]).astype(np.int64)
    pos, si, out = hs + 12 + 4 * ncol, 0, []
    bmb = (rows + 7) // 8
    for c in range(ncol):
        t, attrs = int(block[hs + 12 + 4 * c + 1]), int(block[hs + 12 + 4 * c + 2])
        if t in (0, 1):
            meta = (bmb if attrs & 2 else 0) + (bmb if attrs & 8 else 0)
            kinds = [False] if t == 0 else ([True] if attrs & 1 else [True, False])
        else:
            meta = 10 + (bmb if attrs & 8 else 0)
            distinct = int(block[pos + 2:pos + 6].view(np.uint32)[0])
            kinds = [] if distinct == 0 else ([False, False] if t == 2 else ([True, False] if attrs & 1 else [True, False, False]))
        at = pos + meta
        for is_str in kinds:
            out.append(None if is_str else int(block[at + 2]))
            at = int(ends[si])
            si += 1
        pos = pos + meta if not kinds else at
    return out


@pytest.mark.parametrize("mode", sorted(MODES))
def test_codec_blocks_transform_to_the_raw_blocks(mode):
    n, rpb = 3000, 700
    cols = make_cols(n, 100 + mode)
    assert lib.obgpu_writer_set_cs_stream_encoding(1) == 0
    raw = ob.encode_table(make_cols(n, 100 + mode), rpb)
    try:
        assert lib.obgpu_writer_set_cs_stream_encoding(mode) == 0
        enc = ob.encode_table(cols, rpb)
    finally:
        lib.obgpu_writer_set_cs_stream_encoding(1)
    assert enc.n_blocks == raw.n_blocks
    seen = set()
    for i in range(enc.n_blocks):
        blk, ref = enc.block(i), raw.block(i)
        types = [t for t in stream_types(blk) if t is not None]
        seen.update(types)
        if mode >= 2:
            assert mode in types, (MODES[mode], types)   # the forced codec is used wherever it does not lose to RAW
        t = ora.cs_transform(blk)
        assert all(x in (None, 1) for x in stream_types(t))
        hs = 64
        assert np.array_equal(t[hs:], ref[hs:]), f"block {i}: transformed payload differs from the RAW writer's"
        # and the oracle's block decoder reads the transformed block: every cell equals the RAW block's cell
        a, b = ora.Block(t), ora.Block(ref)
        for c in range(len(cols)):
            for r in (0, 1, a.row_count // 2, a.row_count - 1):
                assert a.cell(c, r) == b.cell(c, r)
    if mode == 0:
        assert len(seen - {1}) >= 2, f"detection never left RAW: {seen}"
    assert sum(enc.sizes) < sum(raw.sizes) or mode in (6, 8), (MODES[mode], sum(enc.sizes), sum(raw.sizes))
