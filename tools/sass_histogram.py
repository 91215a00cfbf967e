"""Per-kernel SASS opcode histogram of libvbmp_b200.so (evidence that the contraction kernels are tcgen05 / TMEM / TMA code):
    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
Counts the mnemonics that matter (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tiled load,
UBLKCP = bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2 / FMUL2 / FADD2 = packed fp32, MUFU) per kernel."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pyvbmp_b200", "libvbmp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA",
        "MUFU", "HMMA", "LDS", "STS", "LDG", "STG", "SHFL", "REDUX", "DFMA", "DADD")
cur, hist, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                hist[cur][op if k.startswith("UTC") or k in ("UTMALDG", "UBLKCP") else k] += 1
                break
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): opcode counts per kernel\n")
for fn, h in hist.items():
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
    name = re.sub(r"\(.*", "", name)
    items = ", ".join(f"{k} {v}" for k, v in sorted(h.items(), key=lambda kv: -kv[1]))
    print(f"{name}\n    {total[fn]} instructions: {items}\n")
