#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output (one row per captured launch, thousands of
columns) into the metric x kernel table kept under profiles/, plus the per-launch DRAM traffic JSON
that bench.py reads for `roofline.traffic`.

usage: ncu_summary.py raw.csv profiles/rNN_ncu_full_summary.csv [profiles/rNN_traffic.json "label"]"""
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(raw)) if r and not r[0].startswith("==")]
    head, units, launches = rows[0], rows[1], rows[2:]
    col = {name: k for k, name in enumerate(head)}
    names = [r[col["Kernel Name"]] for r in launches]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [n[:28] for n in names])
        w.writerow(["Kernel Name", ""] + names)
        for m in KEEP:
            if m in col:
                w.writerow([m, units[col[m]]] + [r[col[m]] for r in launches])
    if len(sys.argv) > 3:
        traffic = {}
        for r, n in zip(launches, names):
            key = n.split("<")[0].replace("void ", "").strip()
            b = sum(float(r[col[m]].replace(",", "")) * TO_BYTES[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic.setdefault(key, b)
        json.dump(dict(workload=sys.argv[4] if len(sys.argv) > 4 else "", source=f"{out} (ncu --set full, one launch each)",
                       dram_bytes_per_launch=traffic), open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
