"""Top source lines by executed warp instructions from an ncu report (--set full --import-source on).
usage: python tools/ncu_top_lines.py report.ncu-rep [N]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None
agg = collections.Counter(); stall = collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if "Instructions Executed" in r:
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    ie = r[hdr.index("Instructions Executed")]; st = r[hdr.index("Warp Stall Sampling (All Samples)")]
    try:
        ie = int(ie); st = int(st or 0)
    except ValueError:
        continue
    key = (r[0], r[1].strip()[:120])
    agg[key] += ie; stall[key] += st
tot = sum(agg.values()); tots = sum(stall.values())
print("total warp instructions", tot, "stall samples", tots)
for k, v in agg.most_common(n):
    print(f"{v:>12} {100*v/tot:5.1f}%  stall {100*stall[k]/max(tots,1):5.1f}%  {k[0]:>5} {k[1]}")
