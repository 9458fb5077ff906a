"""Summarise an ncu launch list (csv, --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum) of
bench.py: takes the LAST complete step (delimited by clip_minmax_kernel launches) and prints a per-kernel table."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ix = {k: i for i, k in enumerate(rows[hdr])}
launches = {}
for r in rows[hdr + 1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    d = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * \
        {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(r[ix["Metric Unit"]], 1.0)
ids = sorted(launches)
marks = [i for i in ids if "clip_minmax" in launches[i]["name"]]
lo, hi = marks[-2], marks[-1]
agg = {}
for i in ids:
    if lo <= i < hi:
        L = launches[i]
        name = re.sub(r"^void |\(.*$", "", L["name"])
        name = re.sub(r"^\(anonymous namespace\)::|^<unnamed>::", "", name)
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += L.get("gpu__time_duration.sum", 0); a[2] += L.get("dram__bytes_read.sum", 0); a[3] += L.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print(f"total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
print("| kernel | launches | total us | share | DRAM read MB | DRAM write MB | DRAM TB/s |\n|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:70]}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.2f}% | {a[2]:.1f} | {a[3]:.1f} | {(a[2] + a[3]) / max(a[1], 1e-9):.2f} |")
