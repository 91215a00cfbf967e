"""Summarise an `ncu --page source --csv` export: samples by opcode and the hottest SASS lines with their stall reasons
(dev tool).  usage: python tools/ncu_source_top.py <source.csv> [n_lines]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
h0 = next(i for i, r in enumerate(rows[:10]) if "Source" in r)
hdr, data = rows[h0], rows[h0 + 1:]
ix = {n: i for i, n in enumerate(hdr)}
def f(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("total samples", int(tot))
agg = defaultdict(lambda: [0, 0, 0])
for r in data:
    t = r[ix["Source"]].split()
    op = t[0] if t else ""
    if op.startswith("@") and len(t) > 1: op = t[1]
    agg[op][0] += f(r, "# Samples"); agg[op][1] += f(r, "Instructions Executed"); agg[op][2] += f(r, "L1 Wavefronts Shared")
for op, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{op:32s} samples {v[0]:9.0f} ({100 * v[0] / tot:5.1f}%) inst {v[1]:13.0f} smem wavefronts {v[2]:13.0f}")
stalls = [n for n in hdr if n.startswith("stall_") and "Not" not in n]
print({n[6:]: int(sum(f(r, n) for r in data)) for n in stalls})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
top = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:n]
for i in sorted(top):
    r = data[i]
    st = {s[6:]: int(f(r, s)) for s in stalls if f(r, s) > 0.1 * f(r, "# Samples")}
    print(i, r[ix["Source"]][:72].ljust(72), int(f(r, "# Samples")), int(f(r, "Instructions Executed")), st)
