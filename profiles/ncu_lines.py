"""Stall samples and executed instructions per CUDA source line of one kernel of an `ncu --set full --import-source on`
capture (the library is built with -lineinfo):

    python profiles/ncu_lines.py X.ncu-rep [top] [kernel-name regex]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if len(sys.argv) > 3:
        cmd += ["--kernel-name", "regex:" + sys.argv[3]]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    per = collections.OrderedDict()
    cur = None
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) < len(hdr):
            continue
        if r[0] == "Line No":
            continue
        if r[0]:
            cur = (int(r[0]), r[1])
            per.setdefault(cur, [0, 0])
            continue
        if cur is None:
            continue
        per[cur][0] += int(r[i_s]) if r[i_s].isdigit() else 0
        per[cur][1] += int(r[i_i]) if r[i_i].isdigit() else 0
    ts, ti = sum(v[0] for v in per.values()), sum(v[1] for v in per.values())
    print(f"# {rep}: {ts} stall samples, {ti} warp instructions")
    for (ln, src), (s, i) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / max(ts, 1):5.1f} % samples {100 * i / max(ti, 1):5.1f} % inst  L{ln}: {src.strip()[:100]}")


if __name__ == "__main__":
    main()
