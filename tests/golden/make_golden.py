"""Generates tests/golden/ref_vectors.npz: golden input / output vectors produced by the REAL reference code that compiles in
this container (oracle/_ref/*.so, built by oracle/Makefile from /root/reference, never copied):

  * integer stream codecs (deps/oblib/src/lib/codec): reference-ENCODED bytes of fixed datasets for every codec x width
  * ob_crc64_sse42 (deps/oblib/src/lib/checksum/ob_crc64.cpp): checksums of fixed buffers at three seeds
  * ObBitStream (storage/blocksstable/encoding/ob_bit_stream.h): buffers written by memory_safe_set at 64 widths
  * ObBitmap::get_row_ids (deps/oblib/src/lib/container/ob_bitmap.cpp): row-id windows with limit / id_offset / from
  * ObIntArrayFuncTable lower_bound / upper_bound (encoding/ob_integer_array.h)

The fixtures travel with the repo: tests/test_golden_vectors.py checks the oracle, the writer's encoders and (on the GPU box)
the device against them without /root/reference or oracle/_ref being present. Run from the repo root:
    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding as ora  # noqa: E402

CODEC_TYPES = (1, 2, 3, 4, 5, 6, 8)
WIDTHS = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


def codec_datasets(dt, rng):
    bits = np.dtype(dt).itemsize * 8
    top = (1 << bits) - 1
    out = []
    for n in (1, 31, 129, 300):
        out.append(("zeros", np.zeros(n, dtype=dt)))
        out.append(("minmax", np.array([0, top] * n, dtype=dt)[:n]))
        out.append(("mono", (np.arange(n, dtype=np.uint64) * 3 + 7).astype(dt)))
        out.append(("runs", np.repeat(rng.integers(0, min(top, 1000) + 1, size=n // 40 + 1), 40)[:n].astype(dt)))
        out.append(("random_full", rng.integers(0, top, size=n, dtype=np.uint64, endpoint=True).astype(dt)))
        spikes = rng.integers(0, 16, size=n).astype(np.uint64)
        spikes[rng.random(n) < 0.07] = top
        out.append(("spikes", spikes.astype(dt)))
        walk = np.cumsum(rng.integers(-40, 41, size=n)).astype(np.int64) + (1 << (bits - 2))
        out.append(("walk", (walk.astype(np.uint64) & np.uint64(top)).astype(dt)))
    return out


def main():
    ora.build_oracle()
    for lib in (ora.REF_CODEC_LIB, ora.REF_MISC_LIB, ora.REF_BITMAP_LIB, ora.REF_LIB):
        if not os.path.exists(lib):
            raise SystemExit(f"{lib} missing: needs /root/reference (oracle/Makefile)")
    out = {}
    # ---- stream codecs ------------------------------------------------------------------------------------------------
    L = C.CDLL(ora.REF_CODEC_LIB)
    L.ref_codec_encode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    index = []
    vals_parts, enc_parts = [], []
    vpos = epos = 0
    for t in CODEC_TYPES:
        for ub, dt in WIDTHS.items():
            rng = np.random.default_rng(7000 + 10 * t + ub)
            for name, a in codec_datasets(dt, rng):
                a = np.ascontiguousarray(a)
                buf = np.zeros(a.size * ub * 2 + 4096, dtype=np.uint8)
                ol = C.c_uint64(0)
                assert L.ref_codec_encode(t, ub, a.ctypes.data, a.size, buf.ctypes.data, buf.size, C.byref(ol)) == 0
                v64 = a.astype(np.uint64)
                index.append((t, ub, a.size, vpos, epos, ol.value))
                vals_parts.append(v64)
                enc_parts.append(buf[:ol.value].copy())
                vpos += a.size
                epos += ol.value
    out["codec_index"] = np.array(index, dtype=np.int64)          # type, width bytes, count, value offset, byte offset, byte count
    out["codec_values"] = np.concatenate(vals_parts)
    out["codec_bytes"] = np.concatenate(enc_parts)
    # ---- crc ---------------------------------------------------------------------------------------------------------------
    M = C.CDLL(ora.REF_MISC_LIB)
    M.ref_crc64_sse42.restype = C.c_uint64
    M.ref_crc64_sse42.argtypes = [C.c_uint64, C.c_void_p, C.c_int64]
    rng = np.random.default_rng(99)
    crc_buf = rng.integers(0, 256, size=20_000, dtype=np.uint8)
    rows = []
    for n in list(range(0, 40)) + [63, 64, 65, 255, 1000, 4096, 16340, 19_990]:
        for shift in (0, 1, 5):
            for seed in (0, 0x1234abcd, 0xffffffff):
                rows.append((n, shift, seed, M.ref_crc64_sse42(seed, crc_buf.ctypes.data + shift, n)))
    out["crc_buf"] = crc_buf
    out["crc_cases"] = np.array(rows, dtype=np.uint64)             # length, offset, seed, checksum
    # ---- integer arrays -------------------------------------------------------------------------------------------------
    for f in ("ref_int_array_lower_bound", "ref_int_array_upper_bound"):
        getattr(M, f).restype = C.c_int64
        getattr(M, f).argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
    ia_rows = []
    ia_arrays = {}
    for byte, dt in ((1, np.uint8), (2, np.uint16), (4, np.uint32), (8, np.uint64)):
        arr = np.sort(rng.integers(0, min((1 << (8 * byte)) - 1, 5000), size=400, dtype=np.uint64)).astype(dt)
        ia_arrays[byte] = arr
        out[f"int_array_{byte}"] = arr
        for key in [0, 1, int(arr[0]), int(arr[17]), int(arr[200]), int(arr[-1]), int(arr[-1]) + 1] + [int(x) for x in rng.integers(0, 5000, size=20)]:
            if key >= (1 << (8 * byte)):
                continue
            for (b, e) in ((0, 400), (10, 390), (200, 201), (5, 5)):
                ia_rows.append((byte, b, e, key, M.ref_int_array_lower_bound(arr.ctypes.data, byte, b, e, key),
                                M.ref_int_array_upper_bound(arr.ctypes.data, byte, b, e, key)))
    out["int_array_cases"] = np.array(ia_rows, dtype=np.int64)      # width, begin, end, key, lower_bound, upper_bound
    # ---- bit stream ----------------------------------------------------------------------------------------------------
    R = C.CDLL(ora.REF_LIB)
    R.ref_bs_memory_safe_set.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]
    R.ref_bs_memory_safe_set.restype = None
    bs_vals = rng.integers(0, 1 << 63, size=64 * 50, dtype=np.uint64) * 2 + rng.integers(0, 2, size=64 * 50, dtype=np.uint64)
    bs_bufs, bs_index, pos = [], [], 0
    for w in range(1, 65):
        v = bs_vals[(w - 1) * 50:w * 50] & np.uint64((1 << w) - 1 if w < 64 else 0xffffffffffffffff)
        nbytes = (50 * w + 7) // 8 + 8
        buf = np.zeros(nbytes, dtype=np.uint8)
        for i in range(50):
            R.ref_bs_memory_safe_set(buf.ctypes.data, i * w, w, int(v[i]))
        bs_bufs.append(buf)
        bs_index.append((w, pos, nbytes))
        pos += nbytes
        bs_vals[(w - 1) * 50:w * 50] = v
    out["bitstream_values"] = bs_vals                               # 50 values per width 1..64
    out["bitstream_bytes"] = np.concatenate(bs_bufs)
    out["bitstream_index"] = np.array(bs_index, dtype=np.int64)     # width, byte offset, byte count
    # ---- bitmap ---------------------------------------------------------------------------------------------------------
    B = C.CDLL(ora.REF_BITMAP_LIB)
    B.ref_bitmap_create.restype = C.c_void_p
    B.ref_bitmap_create.argtypes = [C.c_void_p, C.c_int64]
    B.ref_bitmap_destroy.argtypes = [C.c_void_p]
    B.ref_bitmap_get_row_ids.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64, C.c_int64, C.c_int64]
    bm_all, bm_cases, ids_all = [], [], []
    bpos = ipos = 0
    for n in (1, 7, 64, 65, 200, 1000):
        for dens in (0.0, 0.05, 0.5, 1.0):
            bm = (rng.random(n) < dens).astype(np.uint8)
            bm_all.append(bm)
            for (frm, to, limit, id_off) in ((0, n, n, 0), (0, n, 5, 0), (n // 3, n, 64, 1000), (n // 2, max(n // 2, n - 3), 16, 7)):
                h = B.ref_bitmap_create(bm.ctypes.data, n)
                ids = np.zeros(max(limit, 1) + 8, dtype=np.int32)
                cnt, f = C.c_int64(0), C.c_int64(frm)
                ret = B.ref_bitmap_get_row_ids(h, ids.ctypes.data, C.byref(cnt), C.byref(f), to, limit, id_off)
                B.ref_bitmap_destroy(h)
                bm_cases.append((bpos, n, frm, to, limit, id_off, ret, cnt.value, f.value, ipos))
                ids_all.append(ids[:cnt.value].copy())
                ipos += cnt.value
            bpos += n
    out["bitmap_bytes"] = np.concatenate(bm_all)
    out["bitmap_cases"] = np.array(bm_cases, dtype=np.int64)  # byte offset, n, from, to, limit, id_offset, ret, count, from after, id offset
    out["bitmap_ids"] = np.concatenate(ids_all).astype(np.int32)
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()


# ---- second fixture: micro-blocks of every codec as this repo's writer emits them today (regression anchors) -------------------------
# Not reference output (the reference ships no golden blocks and its encoder cannot be built here): these pin the writer's bytes against
# accidental format drift. Every block in the file passed the REAL reference's ObMicroBlockHeader::deserialize_and_check_record when it
# was generated (asserted below), and tests/test_golden_blocks.py requires the oracle to decode the bytes to the recorded cells.
def make_blocks():
    import test_micro_header_ref_kat as mk
    L = C.CDLL(ora.REF_MACRO_LIB)
    L.ref_micro_block_check.argtypes = [C.c_void_p, C.c_int64]
    out = {}
    names = []
    for name, b in mk.blocks():
        b = np.ascontiguousarray(b)
        assert L.ref_micro_block_check(b.ctypes.data, b.size) == 0, name
        blk = ora.Block(b)
        names.append(name)
        out[f"block_{name}"] = b
        rows = list(range(0, blk.row_count, max(1, blk.row_count // 40))) + [blk.row_count - 1]
        out[f"rows_{name}"] = np.array(rows, dtype=np.int32)
        for c in range(blk.column_count):
            cells = [blk.cell(c, r) for r in rows]
            is_null = np.array([x is None for x in cells], dtype=np.uint8)
            if any(isinstance(x, (bytes, bytearray)) for x in cells):
                heap = b"".join(b"" if x is None else bytes(x) for x in cells)
                lens = np.array([0 if x is None else len(x) for x in cells], dtype=np.int64)
                out[f"str_{name}_{c}"] = np.frombuffer(heap + b"\0", dtype=np.uint8).copy()
                out[f"len_{name}_{c}"] = lens
            else:
                out[f"int_{name}_{c}"] = np.array([0 if x is None else (x & 0xffffffffffffffff) for x in cells], dtype=np.uint64)
            out[f"null_{name}_{c}"] = is_null
    out["names"] = np.array(names)
    path = os.path.join(HERE, "writer_blocks.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(names), "blocks")


if __name__ == "__main__":
    make_blocks()
