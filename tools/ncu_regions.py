#!/usr/bin/env python
"""Instruction / stall-sample shares of the regions of tps_window_bp_kernel in an .ncu-rep (needs -lineinfo and
--import-source on).  Regions are found by marker strings in tps_kernels.cuh.  Usage: ncu_regions.py x.ncu-rep"""
import csv, io, os, subprocess, sys
from collections import defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKS = [("stage_fn", "tps_stage_linear(const TpsPacked"), ("oriented_fn", "tps_oriented_word(const uint32_t"),
         ("match_fn", "per-position matching"), ("rowlogic", "step-1 row logic"), ("cp", "#define TPS_K4_THREADS 128"),
         ("k4_kernel", "Stand-alone K4"), ("scan", "tps_block_excl_scan(uint32_t v"), ("dilate", "tps_dilate(const uint2"),
         ("planes", "tps_presence_planes(const uint2"), ("prefetch", "tps_cp_async16(void"),
         ("kernel_setup", "tps_window_bp_kernel(const TpsScanArgs a"),
         ("stage_raw", "tps_stage_entry_raw(const uint32_t"), ("bp_window_fn", "tps_bp_window(uint32_t ls"),
         ("prologue", "auto tiles_of = "), ("read_loop", "for (uint32_t i = 0;; ++i) {"),
         ("tile_top", "for (uint32_t t = 0; t < ntiles; ++t, buf ^= 1u) {"), ("stagecall", "const uint32_t phase = (uint32_t)(g0 & 15u);"),
         ("pass1", "(1) match words of word q"), ("pass2", "(2) presence rows R_p of word q"),
         ("windows", "/* (3) windows, one group of five"), ("tile_end", "tps_cp_async_wait_all();\n      __syncthreads();\n    }"),
         ("cp_call", "/* the read's change point, straight"), ("end", "K5")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, L = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 20 and r[0] == "Line No":
            hdr = r
        elif len(r) > 20 and hdr and r[2] == "-":
            try:
                L.append((fname, int(r[0]), int(r[hdr.index("Instructions Executed")] or 0), int(r[hdr.index("# Samples")] or 0)))
            except ValueError:
                pass
    text = open(os.path.join(REPO, "topsicle_b200", "csrc", "tps_kernels.cuh")).read()
    marks = []
    for name, m in MARKS:
        i = text.find(m)
        if i >= 0:
            marks.append((name, text.count("\n", 0, i) + 1))
    marks.sort(key=lambda x: x[1])

    def region(f, ln):
        if f == "tps_bitops.h":
            return "bitops:stage" if ln < 130 else ("bitops:greedy" if ln < 182 else "bitops:cp")
        if f != "tps_kernels.cuh":
            return "intr:" + f
        name = "pre"
        for n, l in marks:
            if ln >= l:
                name = n
        return name

    ti = sum(x[2] for x in L) or 1
    ts = sum(x[3] for x in L) or 1
    d = defaultdict(lambda: [0, 0])
    for f, ln, ni, ns in L:
        k = region(f, ln)
        d[k][0] += ni
        d[k][1] += ns
    print(f"total warp instructions {ti / 1e6:.1f}M, samples {ts}")
    for k, (ni, ns) in sorted(d.items(), key=lambda x: -x[1][0]):
        print(f"{k:24s} {100 * ni / ti:5.1f}% inst ({ni / 1e6:5.1f}M) {100 * ns / ts:5.1f}% samples")


if __name__ == "__main__":
    main()
