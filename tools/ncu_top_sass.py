"""Reads an `ncu --page source --csv` export and prints the SASS instructions with the most warp-stall samples.
    python tools/ncu_top_sass.py source.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = []
hdr = None
kern = ""
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kern = r[1]; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    d = dict(zip(hdr, r))
    try:
        n = int(d["# Samples"])
    except (KeyError, ValueError):
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    out.append((n, kern[:40], d["Address"][-5:], d["Source"].strip()[:90], stalls, d.get("Instructions Executed", "")))
tot = sum(o[0] for o in out) or 1
print("total samples", tot, "instructions", len(out))
for idx, o in enumerate(out):
    pass
order = sorted(range(len(out)), key=lambda i: -out[i][0])[:top]
for i in sorted(order):
    n, kern, addr, src, stalls, ex = out[i]
    print("%5d %5.1f%%  #%-5d %s  | %s | exec %s | %s" % (n, 100.0 * n / tot, i, addr, src, ex, " ".join("%s=%d" % kv for kv in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])))
