"""Dynamic instruction mix of one kernel from an ncu report captured with --import-source on:
  ncu -i report.ncu-rep --page source --csv --print-source sass > src.csv ; python scripts/sass_mix.py src.csv [steps]
`steps` = warp-level march steps of the launch (CTAs x warps x steps per CTA) to print instructions per step."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
c, smp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iE:
        continue
    t = re.sub(r"^\s*@!?U?P\d+\s+", "", r[iS].strip())
    op = t.split()[0] if t else "?"
    op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDS", "SHFL", "LDG", "STG", "ISETP")) else op.split(".")[0]
    n = int(r[iE] or 0)
    c[op] += n
    smp[op] += int(r[iSm] or 0)
    tot += n
print("warp instructions executed:", tot, " per step:" if steps else "", round(tot / steps, 1) if steps else "")
fp64 = sum(v for k, v in c.items() if k in ("DADD", "DFMA", "DMUL", "DSETP", "DMNMX"))
print(f"FP64 share {100 * fp64 / tot:.1f} %")
stot = sum(smp.values())
for k, v in sorted(c.items(), key=lambda kv: -kv[1]):
    if v * 1000 < tot:
        continue
    print(f"  {k:12s} {100 * v / tot:5.1f} %  " + (f"{v / steps:6.1f} / step  " if steps else "") + f"samples {100 * smp[k] / max(stot, 1):5.1f} %")
