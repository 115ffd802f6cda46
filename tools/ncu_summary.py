"""Merge `ncu --page raw --csv` exports (gpurun_out/ncu_<name>.raw.csv, written by tools/gpu_check.sh) into
profiles/r02_ncu_summary.json: per capture the handful of metrics DESIGN.md quotes.
usage: python tools/ncu_summary.py name [name ...]     (names as in gpurun_out/ncu_<name>.raw.csv)"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main():
    path = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name in sys.argv[1:]:
        rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"ncu_{name}.raw.csv"))))
        head, units, vals = rows[0], rows[1], rows[2]
        entry = {"kernel": vals[head.index("Kernel Name")]}
        for k in KEEP:
            if k in head:
                i = head.index(k)
                entry[k] = [vals[i], units[i]]
        out[name] = entry
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, "with", len(out), "captures")


if __name__ == "__main__":
    main()
