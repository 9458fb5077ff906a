"""Join an `ncu --page source --csv` SASS listing of one launch with the line table of the current build
(`nvdisasm -g` on the kernel's cubin section) and print the hottest CUDA source lines with their stall reasons.

    ncu -i X.ncu-rep --page source --csv --launch-skip N --launch-count 1 > src.csv
    cuobjdump -xelf all libnm_b200.so ; nvdisasm -g conv_tc.sm_100a.cubin > all.sass   (cut out the kernel's .text section)
    python tools/ncu_source_lines.py src.csv kernel.sass path/to/source.cu [top]

The profile and the build must come from the same source (the instruction counts are checked)."""
import collections
import csv
import os
import re
import sys

src_csv, sass, cu = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line, ins = None, []
for l in open(sass):
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        line = int(m.group(2)) if m.group(1).endswith(cu.split("/")[-1]) else m.group(1).split("/")[-1] + ":" + m.group(2)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), line))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
data = rows[2:]
assert len(ins) == len(data), (len(ins), len(data))
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
by = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for (s, l), r in zip(ins, data):
    by[l][0] += int(r[ix["# Samples"]])
    by[l][1] += int(r[ix["Instructions Executed"]])
    for k in stalls:
        by[l][2][k[6:]] += int(r[ix[k]] or 0)
tot = sum(v[0] for v in by.values())
text = open(cu).read().split("\n")
print("| line | samples | warp instructions | top stall reasons | source |\n|---|---|---|---|---|")
for l, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    reasons = ", ".join(f"{k} {100 * c / max(v[0], 1):.0f}%" for k, c in v[2].most_common(3))
    if isinstance(l, int):
        code = text[l - 1].strip()[:100].replace("|", "\\|")
    elif l and os.path.exists(os.path.join(os.path.dirname(cu), l.split(":")[0])):
        code = open(os.path.join(os.path.dirname(cu), l.split(":")[0])).read().split("\n")[int(l.split(":")[1]) - 1].strip()[:100].replace("|", "\\|")
    else:
        code = "(inlined from another file / no line info)"
    print(f"| {l} | {100 * v[0] / tot:.1f}% | {v[1]} | {reasons} | `{code}` |")
