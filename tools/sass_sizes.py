"""Per-kernel SASS size table of a built library: instructions, IMAD.WIDE (one per 32x32->64 multiply-add of the field
arithmetic), CALL sites, registers are in `nvcc -Xptxas -v`.  Runs on the CPU (cuobjdump only).
Usage: python tools/sass_sizes.py halo2-verifier_b200/libh2v_b200.so [other.so]   (two libraries: side by side)"""
import re
import subprocess
import sys
from collections import OrderedDict


def table(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    rows, name = OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            name = name.replace("void ", "").replace("h2v::", "")
            rows[name] = [0, 0, 0]
            continue
        if name is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            continue
        rows[name][0] += 1
        if "IMAD.WIDE" in line:
            rows[name][1] += 1
        if " CALL" in line:
            rows[name][2] += 1
    return rows


def main():
    tabs = [table(p) for p in sys.argv[1:3]]
    names = sorted(set().union(*[t.keys() for t in tabs]), key=lambda n: -max(t.get(n, [0])[0] for t in tabs))
    hdr = "%-34s" % "kernel" + "".join("%12s%12s%8s" % ("instr", "IMAD.WIDE", "CALL") for _ in tabs)
    print(hdr)
    for n in names:
        print("%-34s" % n[:33] + "".join(("%12d%12d%8d" % tuple(t[n])) if n in t else "%12s%12s%8s" % ("-", "-", "-") for t in tabs))
    print("%-34s" % "total" + "".join("%12d%12d%8d" % tuple(sum(v[i] for v in t.values()) for i in range(3)) for t in tabs))


if __name__ == "__main__":
    main()
