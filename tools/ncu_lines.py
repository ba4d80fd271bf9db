#!/usr/bin/env python
"""Per source line: warp instructions executed and stall samples of one kernel in an .ncu-rep (--import-source on).
usage: python tools/ncu_lines.py rep.ncu-rep [kernel-substring] [top-n]"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, sub="", top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] +
                         (["-k", f"regex:{sub}"] if sub else []), capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    inst = collections.Counter()
    samp = collections.Counter()
    src = {}
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
            continue
        ln = int(r[0])
        src.setdefault(ln, r[1].strip())
        try:
            inst[ln] += int(float(r[iI] or 0))
            samp[ln] += int(float(r[iS] or 0))
        except ValueError:
            pass
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    print(f"# {rep} {sub}: {ti} warp instructions, {ts} samples")
    for ln, n in inst.most_common(top):
        print(f"{ln:5d} {100.0 * n / ti:5.1f}% inst {100.0 * samp[ln] / ts:5.1f}% samp  {src[ln][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 40)
