"""Print the key metrics of every kernel in an `ncu --page raw --csv` dump.

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_keys.py raw.csv
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[idx["Kernel Name"]][:90])
        for k in KEYS:
            if k in idx:
                print(f"   {k:70s} {r[idx[k]]:>16s} {units[idx[k]]}")
        stalls = [(float(r[i].replace(",", "")), h) for h, i in idx.items()
                  if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i]]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"   stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {v:8.2f}")


if __name__ == "__main__":
    main()
