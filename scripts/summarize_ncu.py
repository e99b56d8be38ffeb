"""Summarise ncu outputs into profiles/ (run in the build container; no GPU needed).
  python scripts/summarize_ncu.py launches <launches.csv> <out.md>
  python scripts/summarize_ncu.py full <report.ncu-rep> <out.md>
"""
import collections, csv, io, re, subprocess, sys

def short(name):
    name = re.sub(r"kl::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]

def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                "(per-launch times are cold-cache and serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t/1e3:.1f} | {t/1e3/n:.1f} | {100*t/tot:.1f}% |\n")
    print(open(out).read())

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]

def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\nOne block per captured launch; traffic = dram read + write per launch.\n")
        for r in rows[2:]:
            f.write(f"\n## `{short(r[idx['Kernel Name']])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            rd = wr = None
            for w in WANT:
                if w in idx:
                    f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
                    if w == "dram__bytes_read.sum": rd = (float(r[idx[w]]), units[idx[w]])
                    if w == "dram__bytes_write.sum": wr = (float(r[idx[w]]), units[idx[w]])
            if rd and wr:
                sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
                t = rd[0] * sc.get(rd[1], 1) + wr[0] * sc.get(wr[1], 1)
                f.write(f"| **traffic (read+write)** | {t/1e9:.4f} | GB |\n")
    print(open(out).read()[:6000])

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
