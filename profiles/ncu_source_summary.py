"""Summarise an `ncu --set full --import-source on` capture of one kernel: key metrics (raw page) and where the
instructions and the stall samples are (source page, SASS view), by code region and for the hottest instructions.

    python profiles/ncu_source_summary.py gpurun_out/X.ncu-rep [--regions 0x330,0xa80,...] > profiles/X.txt

Regions are SASS offsets that split the kernel (e.g. the warp roles of a warp-specialised kernel); without them
the kernel is cut into 0x400-byte blocks.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    regions = None
    if "--regions" in sys.argv:
        regions = [int(x, 16) for x in sys.argv[sys.argv.index("--regions") + 1].split(",")]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in raw[2:]:
        print("==", r[idx["Kernel Name"]][:100])
        for k in KEYS:
            if k in idx:
                print(f"   {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")
    src = page(rep, "source")
    hdr = src[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = src[2:]
    base = int(data[0][idx["Address"]], 16)
    inst = [int(r[idx["Instructions Executed"]] or 0) for r in data]
    samp = [int(r[idx["# Samples"]] or 0) for r in data]
    pcs = [int(r[idx["Address"]], 16) - base for r in data]
    ti, ts = sum(inst), sum(samp)
    print(f"\nsource page: {ti} warp instructions, {ts} stall samples")
    cuts = regions if regions else list(range(0, pcs[-1] + 0x400, 0x400))
    cuts = sorted(set([0] + cuts + [pcs[-1] + 16]))
    print("region (SASS offset)        instructions        samples")
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        i = sum(v for v, p in zip(inst, pcs) if lo <= p < hi)
        s = sum(v for v, p in zip(samp, pcs) if lo <= p < hi)
        if i or s:
            print(f"  [{lo:#07x}, {hi:#07x})   {100 * i / ti:6.2f} %  {i:>12d}   {100 * s / ts:6.2f} %")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("\nhottest instructions by stall samples")
    for j in sorted(range(len(data)), key=lambda j: -samp[j])[:25]:
        r = data[j]
        top = sorted(((int(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:2]
        print(f"  {pcs[j]:#07x} {r[idx['Source']][:58]:58s} inst {100 * inst[j] / ti:5.2f} %  samples {100 * samp[j] / ts:5.2f} %  "
              + ", ".join(f"{h[6:]} {v}" for v, h in top if v))


if __name__ == "__main__":
    main()
