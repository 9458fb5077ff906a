"""Extract the judged metrics of an `ncu --set full` report into a markdown table per launch.
usage: ncu -i X.ncu-rep --page raw --csv > X.csv ; python tools/ncu_extract.py X.csv > profiles/...md"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size"]
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
ix = {k: i for i, k in enumerate(names)}
for n, r in enumerate(rows[hdr + 2:]):
    if len(r) < len(names):
        continue
    kn = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    print(f"\n### launch {n}: `{kn}` block {r[ix['Block Size']]} grid {r[ix['Grid Size']]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for m in WANT:
        if m in ix:
            print(f"| {m} | {r[ix[m]]} | {units[ix[m]]} |")
