#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep (captured with
--import-source on, built with -lineinfo).  Usage: python tools/ncu_lines.py x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, lines = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 20 and r[0] == "Line No":
            hdr = r
        elif len(r) > 20 and hdr and r[2] == "-":       # a source line (SASS rows carry an address)
            try:
                lines.append((fname, int(r[0]), r[1].strip(), int(r[hdr.index("Instructions Executed")] or 0),
                              int(r[hdr.index("# Samples")] or 0)))
            except ValueError:
                pass
    tot_i = sum(x[3] for x in lines) or 1
    tot_s = sum(x[4] for x in lines) or 1
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for f, ln, src, ni, ns in sorted(lines, key=lambda x: -x[4])[:top]:
        print(f"{f}:{ln:5d} {100 * ni / tot_i:5.1f}% inst {100 * ns / tot_s:5.1f}% samples  {src[:110]}")


if __name__ == "__main__":
    main()
