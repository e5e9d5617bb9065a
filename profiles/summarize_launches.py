"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into one step's kernel table.

    python profiles/summarize_launches.py gpurun_out/launches.csv [launches_per_step]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        seq.append((name, v))
    own = [(n, v) for n, v in seq if not n.startswith("void at::") and "elementwise" not in n and "nccl" not in n]
    # one step = from a decode_mark_kernel up to (not including) the next one; take the last complete one
    starts = [i for i, (n, _) in enumerate(own) if "decode_mark" in n]
    crops = [i for i, (n, _) in enumerate(own) if "crop_resize_pad" in n]
    start = max(s for s in starts if any(c > s for c in crops))
    later = [s for s in starts if s > start]
    step = own[start:(later[0] if later else len(own))]
    tot = sum(v for _, v in step)
    print(f"# {path}: last complete step, {len(step)} launches, {tot:.1f} us (cold-cache, serialised; compare shares)")
    agg = {}
    for n, v in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f"{'kernel':48s} {'launches':>8s} {'us':>10s} {'share':>7s}")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:48s} {c:8d} {v:10.1f} {100 * v / tot:6.1f}%")


if __name__ == "__main__":
    main()
