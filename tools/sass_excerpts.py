"""profiles/r2_sass_excerpts.txt: per kernel, the SASS instruction count, a histogram of the memory / synchronisation mnemonics that
characterise the design and the first occurrence of the TMA / mbarrier / cp.async instructions with their neighbours.
usage: python tools/sass_excerpts.py [out]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "oceanbase_b200", "csrc", "libobgpu_scan.so")
lines = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
starts = [(i, l.split("Function : ")[1].strip()) for i, l in enumerate(lines) if "Function :" in l]
want = ["obgpu_count_pipe_kernel", "obgpu_project_pipe_kernel", "obgpu_count_kernel", "obgpu_project_kernelILb0", "obgpu_index_kernel",
        "pass_kernel", "kway_kernel", "fuse_kernel", "cs_decode_kernel", "obgpu_encode_blocks_kernel", "obgpu_macro_realign_kernel",
        "obgpu_macro_walk_kernel"]
out = ["cuobjdump -sass oceanbase_b200/csrc/libobgpu_scan.so (sm_100a).",
       "UBLKCP = cp.async.bulk (TMA bulk copy global->shared), SYNCS = mbarrier ops (arrive.expect_tx / try_wait), LDGSTS = cp.async (global->shared",
       "without registers), LDS/STS = shared loads/stores, SHF = funnel shift (the bit-granular load path), VOTE = ballot, LDL/STL = local memory.", ""]
keys = ["UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "DEPBAR", "LDG", "STG", "LDS", "STS", "SHF", "VOTE", "POPC", "SHFL", "REDUX", "ATOMS", "ATOMG", "RED",
        "BAR", "LDL", "STL", "BRA", "UBLKCP.S", "FENCE", "MEMBAR"]
for k, (i, name) in enumerate(starts):
    if not any(w in name for w in want):
        continue
    end = starts[k + 1][0] if k + 1 < len(starts) else len(lines)
    ins = []
    for l in lines[i:end]:
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(.*?);", l)
        if m:
            ins.append(m.group(1).strip())
    mn = collections.Counter()
    for x in ins:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", x)
        if m:
            mn[m.group(2)] += 1
    out.append("=" * 100)
    out.append(f"{name}: {len(ins)} SASS instructions")
    out.append("  " + "  ".join(f"{k}={mn.get(k, 0)}" for k in keys))
    for key in ["UBLKCP", "SYNCS", "LDGSTS"]:
        for j, x in enumerate(ins):
            if re.match(r"(@!?U?P\d+\s+)?" + key + r"\b", x):
                out.append(f"  first {key}:")
                out += ["      " + y for y in ins[max(0, j - 2): j + 3]]
                break
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_excerpts.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(out), "lines")
