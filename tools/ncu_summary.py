#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key metrics per captured launch + stall mix + hottest SASS ranges.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); print(f"   {w:68s} {r[i]:>18s} {units[i]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    h = None; kernels = []
    for r in src:
        if len(r) > 5 and r[0] == "Address":
            h = {n: i for i, n in enumerate(r)}; kernels.append([]); continue
        if h and len(r) >= len(h) and r[h["# Samples"]].isdigit():
            kernels[-1].append(r)
    if not kernels:
        return
    data = kernels[0]
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {s: sum(int(r[h[s]] or 0) for r in data) for s in stalls}
    tot = sum(agg.values()) or 1
    print("stall mix:", ", ".join(f"{k[6:]} {v / tot:.0%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v / tot >= 0.02))
    ex = sum(int(r[h["Instructions Executed"]]) for r in data); th = sum(int(r[h["Thread Instructions Executed"]]) for r in data)
    print(f"warp instructions {ex}, avg active threads {th / max(1, ex):.2f}")
    print(f"hottest {top} instructions (samples, avg threads, executions, SASS):")
    for r in sorted(data, key=lambda r: -int(r[h["# Samples"]]))[:top]:
        print(f"   {r[h['# Samples']]:>7s} {r[h['Avg. Threads Executed']]:>5s} {int(r[h['Instructions Executed']]):>10d}  {r[h['Source']][:100]}")


if __name__ == "__main__":
    main()
