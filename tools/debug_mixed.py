"""which combination breaks: CS coded alone / PAX rebuilt alone / mixed raw / mixed coded; per filter"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oceanbase_b200 as ob
rng = np.random.default_rng(21)
n = 2600
k = np.arange(n, dtype=np.int64) * 3
hexs = [bytes(np.frombuffer(b"0123456789abcdef", dtype=np.uint8)[rng.integers(0, 16, size=rng.integers(1, 14))]) for _ in range(n)]
plain = [b"p%d" % (i % 13) for i in range(n)]
nl = (rng.random(n) < 0.1).astype(np.uint8)
pax = ob.encode_table([ob.Column(ob.OBJ_INT, ob.ENC_INTEGER_BASE_DIFF, k), ob.Column(ob.OBJ_VARCHAR, ob.ENC_HEX_PACKING, hexs, nulls=nl),
                       ob.Column(ob.OBJ_VARCHAR, ob.ENC_DICT, plain)], 400)
cs_cols = [ob.Column(ob.OBJ_INT, ob.ENC_CS_INTEGER, k), ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STRING, hexs, nulls=nl),
           ob.Column(ob.OBJ_VARCHAR, ob.ENC_CS_STR_DICT, plain)]
cs_raw = ob.encode_table(cs_cols, 400)
ob.capi.lib.obgpu_writer_set_cs_stream_encoding(0)
cs = ob.encode_table(cs_cols, 400)
ob.capi.lib.obgpu_writer_set_cs_stream_encoding(1)

def interleave(a, b):
    image = np.concatenate([a.image, b.image])
    offs = np.concatenate([a.offsets, b.offsets + a.image.size])
    sizes = np.concatenate([a.sizes, b.sizes])
    order = np.argsort(np.concatenate([np.arange(a.n_blocks) * 2, np.arange(b.n_blocks) * 2 + 1]), kind="stable")
    return ob.TableImage(image, offs[order], sizes[order], a.total_rows + b.total_rows, 3)

ctx = ob.ScanContext(0)
filters = {"none": None, "and": ob.And([ob.White(0, ob.WHITE_OP_GE, (900,)), ob.White(1, ob.WHITE_OP_NE, (hexs[5],))]), "in": ob.White(2, ob.WHITE_OP_IN, (b"p3", b"p7"))}
for name, t in (("cs_coded", cs), ("cs_raw", cs_raw), ("pax", pax), ("mixed_raw", interleave(pax, cs_raw)), ("mixed_coded", interleave(pax, cs)),
                ("mixed_coded_cs_first", interleave(cs, pax))):
    for fn, flt in filters.items():
        for proj in ([0], [0, 1], [0, 2], [0, 1, 2]):
            try:
                b = ctx.open_batch(t)
                r = b.scan(flt, proj, string_base=t.image.ctypes.data)
                s = r.selected_rows
                r.free(); b.close()
                print(name, fn, proj, "ok", s)
            except Exception as e:
                print(name, fn, proj, "FAIL", str(e)[:90])
