"""Static opcode histogram of the loops of one kernel:  cuobjdump -sass -fun <mangled> obj.o | python scripts/sass_loops.py [min_len]"""
import collections
import re
import sys

minlen = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
loops = []
for a, t in ins:
    m = re.search(r"\bBRA\S*\s+.*?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
print(len(ins), "instructions")
for a, b in sorted(set(loops)):
    n = (b - a) // 16 + 1
    if n < minlen:
        continue
    c = collections.Counter()
    for ad, t in ins:
        if a <= ad <= b:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t.split()[0]
            op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDS", "SHFL")) else op.split(".")[0]
            c[op] += 1
    fp = sum(v for k, v in c.items() if k in ("DADD", "DFMA", "DMUL", "DSETP"))
    print(f"loop {a:#x}..{b:#x}: {n} instructions, FP64 {fp} ({100*fp/n:.0f} %)")
    print("   ", " ".join(f"{k}:{v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1]) if v >= 4))
