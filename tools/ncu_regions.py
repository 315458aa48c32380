"""Summarise an ncu source page (csv) of one kernel launch by code region (runs of equal execution
count), tagging each region with the distinctive SASS it contains. Usage:
   ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_regions.py src.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; hdr = rows[hi]
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
data = [r for r in rows[hi + 1:(his[1] - 1 if len(his) > 1 else None)] if len(r) > ie]
S = lambda lo, hi_: sum(int(data[k][si] or 0) for k in range(lo, hi_))
tot = S(0, len(data)); print("total samples", tot, "instructions", len(data))
cur = None; start = 0; segs = []
for k, r in enumerate(data):
    if r[ie] != cur:
        if cur is not None: segs.append((start, k, cur))
        cur = r[ie]; start = k
segs.append((start, len(data), cur))
KEYS = ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "MUFU.EX2", "SYNCS.PHASECHK", "UTCBAR", "STS", "BAR.SYNC", "STG", "LDG", "SHFL", "MEMBAR", "CCTL")
for lo, hi_, e in segs:
    s = S(lo, hi_)
    if 100.0 * s / tot < min_pct: continue
    tags = sorted({k for i in range(lo, hi_) for k in KEYS if k in data[i][src]})
    offs = sorted({data[i][src].split("+0x")[-1].split("]")[0] for i in range(lo, hi_) if "SYNCS.PHASECHK" in data[i][src] and "+0x" in data[i][src]})
    print(f"#{lo:5d}-{hi_:5d} exec={e:>10s} samples={s:7d} ({100*s/tot:5.1f}%) n_instr={hi_-lo:4d} {tags} {offs}")
