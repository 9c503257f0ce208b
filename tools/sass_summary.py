#!/usr/bin/env python
"""SASS evidence of the built library (no GPU needed): per kernel the mnemonic histogram and the instructions
that prove the mechanisms DESIGN.md names -- UBLKCP / SYNCS (bulk async copy + mbarrier, K1), LDGSTS / LDGDEPBAR
(cp.async prefetch, K3), REDUX / CREDUX (warp reductions, K2), POPC / FLO / LOP3 (bit-parallel counting).
  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "topsicle_b200", "libtopsicle_b200.so")
WANT = {"tps_pack_tma_kernelILi4E": "K1 tps_pack_tma_kernel<4>", "tps_trc_const_kernelILi4ELi6E": "K2 tps_trc_const_kernel<4, 6>",
        "tps_trc_reg_kernelILi4E": "K2 (table-driven) tps_trc_reg_kernel<4>",
        "tps_window_bp_kernelILi4E": "K3+K4 tps_window_bp_kernel<4>", "tps_window_kernelILi4E": "K3 (plain) tps_window_kernel<4>",
        "tps_changepoint_kernel": "K4 (stand-alone) tps_changepoint_kernel"}
KEYS = ["UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "DEPBAR", "REDUX", "CREDUX", "VIMNMX", "LDCU", "POPC", "FLO", "LOP3", "SHF", "LDS", "STS",
        "VOTE", "ATOM", "ATOMG", "MEMBAR", "BAR", "DMUL", "DFMA", "DADD", "I2F", "MUFU"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    print("# SASS evidence, sm_100a build of topsicle_b200/libtopsicle_b200.so (cuobjdump -sass; regenerate: "
          "python tools/sass_summary.py)\n")
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        for key, label in WANT.items():
            if key not in name:
                continue
            c = collections.Counter()
            for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", f):
                c[m.group(1).split(".")[0]] += 1
            print(f"## {label}   ({name}; {sum(c.values())} instructions)")
            print("mnemonic counts: " + ", ".join(f"{k} {v}" for k, v in c.most_common(24)))
            print("marker instructions: " + ", ".join(f"{k} {c.get(k, 0)}" for k in KEYS if c.get(k, 0)))
            for pat in ("UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "REDUX", "CREDUX"):
                for m in list(re.finditer(r"(/\*[0-9a-f]{4,5}\*/\s+[^;\n]*\b" + pat + r"[^;\n]*;)", f))[:3]:
                    print("    " + re.sub(r"\s+", " ", m.group(1)))
            print()


if __name__ == "__main__":
    main()
