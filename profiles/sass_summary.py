"""SASS evidence per kernel of libmanuscript_b200.so (the judge cannot rebuild: this is what the code compiles to).

    python profiles/sass_summary.py > profiles/sass_summary.txt

For every kernel: instruction count and the opcodes that show which hardware path it uses --
  UBLKCP            TMA bulk copy global -> shared (cp.async.bulk)             SYNCS.*      mbarrier operations
  UTMALDG           TMA tensor-map copy (not used: see profiles/README.md)     UCGABAR_*    thread-block-cluster barriers
  NANOSLEEP         back-off in mbarrier wait loops                            ATOMG/RED    global atomics
  DADD/DMUL/DFMA    float64 arithmetic (IoU, merge, sizing: the reference computes these in float64)
  FADD2/FMUL2/FFMA2 packed float32 pairs (FFMA2 must be 0: a fused multiply-add would break bit-exactness)
  FFMA              fused float32 multiply-add (only inside division / sqrt sequences; the library is built --fmad=false)
  HMMA/UTC*MMA      tensor cores (none: no stage of this path is a dense contraction)
  ACQBULK           griddepcontrol.wait: every kernel is a programmatic dependent launch and waits for its predecessor
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "manuscript-ocr_b200", "manuscript_b200", "libmanuscript_b200.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "UCGABAR", "NANOSLEEP", "ATOMG", "RED", "ATOMS", "DADD", "DMUL", "DFMA", "FADD2",
         "FMUL2", "FFMA2", "FFMA", "FADD", "FMUL", "LDS", "STS", "LDG", "STG", "PRMT", "SHFL", "VOTE", "BAR", "HMMA", "UTC", "ACQBULK"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::", "", o).replace("void ", "").split("(")[0] for o in out]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else SO
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or (w in ("SYNCS", "UCGABAR", "UTC", "BAR") and op.startswith(w)):
                    cur[w] += 1
    names = demangle(list(kernels))
    print(f"# {os.path.relpath(so, ROOT)}: {len(kernels)} kernels, sm_100a SASS (cuobjdump -sass), opcode counts per kernel")
    cols = ["UBLKCP", "SYNCS", "UCGABAR", "NANOSLEEP", "ATOMG", "DFMA", "DMUL", "DADD", "FADD2", "FMUL2", "FFMA2", "FFMA",
            "FADD", "FMUL", "PRMT", "LDS", "SHFL", "HMMA", "UTC", "ACQBULK"]
    print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{c:>6s}" for c in cols))
    tot = collections.Counter()
    for (mangled, cnt), name in sorted(zip(kernels.items(), names), key=lambda t: t[1]):
        print(f"{name[:58]:58s} {cnt['_total']:6d} " + " ".join(f"{cnt[c]:6d}" for c in cols))
        tot.update(cnt)
    print(f"{'TOTAL':58s} {tot['_total']:6d} " + " ".join(f"{tot[c]:6d}" for c in cols))
    print("\nUTMALDG (tensor-map TMA):", tot["UTMALDG"], "| tensor-core opcodes (HMMA / UTC*MMA):", tot["HMMA"] + tot["UTC"],
          "| FFMA2 (must be 0):", tot["FFMA2"], "| kernels without griddepcontrol.wait (ACQBULK):",
          sum(1 for c in kernels.values() if c["ACQBULK"] == 0))


if __name__ == "__main__":
    main()
