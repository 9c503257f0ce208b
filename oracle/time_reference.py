#!/usr/bin/env python
"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- how generous is the CPU port to the reference?

Times the UNMODIFIED reference (`/root/reference/Topsicle/main.py:main`, imported through the four shims of
oracle/shims: Bio, ruptures, seaborn, matplotlib) on N shard files of synthetic config-2 reads with
`--threads N` -- the reference parallelises over input files only (main.py:232-235) -- and, on the same reads and
the same cores, the port that bench.py uses as `cpu_baseline` / `--impl reference` (oracle/cpu_baseline.py:
in-memory reads, no parsing, no O(p^2) temp-file rescans).  Runs only where /root/reference exists (the build
container); the result is recorded in BASELINE.md.

  python oracle/time_reference.py [--reads 16000] [--shards 8]
"""
import argparse
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(1, REF)
sys.path.insert(2, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=16000)
    ap.add_argument("--shards", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    from oracle import cpu_baseline
    from topsicle_b200 import synth
    spec = synth.CONFIGS[2]
    bases, off, _ = synth.generate(spec, 0, a.reads)
    buf = bases.tobytes().decode("ascii")
    seqs = [buf[int(off[i]):int(off[i + 1])] for i in range(a.reads)]
    n_bases = int(off[-1])
    work = tempfile.mkdtemp(prefix="tps_ref_")
    try:
        ind, out = os.path.join(work, "in"), os.path.join(work, "out")
        os.makedirs(ind)
        shards = cpu_baseline.deal_shards(seqs, a.shards)          # equal-base shards: the best case for a Pool over files
        for k, sh in enumerate(shards):
            with open(os.path.join(ind, f"shard_{k}.fastq"), "w") as fh:
                for gi, s in sh:
                    fh.write(f"@syn2_{gi}\n{s}\n+\n{'I' * len(s)}\n")
        import Topsicle.main as tmain
        old = sys.argv
        sys.argv = ["topsicle", "--inputDir", ind, "--outputDir", out, "--pattern", "CCCTAA", "--minSeqLength", "9000",
                    "--threads", str(a.shards)]
        t0 = time.perf_counter()
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                tmain.main()
        finally:
            sys.argv = old
        t_ref = time.perf_counter() - t0
        rows_ref = sum(1 for _ in open(os.path.join(out, "telolengths_all.csv"))) - 1
        kw = dict(pattern="CCCTAA", phrase=4, cutoff=0.7, min_len=9000, W=100, slide=6, trim=100, maxlen=20000)
        times, n_pass, _ = cpu_baseline.time_sample(seqs, kw, a.shards, steps=1)
        t_port = times[0]
        print(json.dumps({
            "workload": spec["name"], "reads": a.reads, "gbases": round(n_bases / 1e9, 4), "cores": a.shards,
            "reference_unmodified": {"seconds": round(t_ref, 2), "gbases_per_s": round(n_bases / t_ref / 1e9, 5),
                                     "csv_rows": rows_ref,
                                     "what": "Topsicle.main on shard FASTQ files (shim parser instead of Biopython, "
                                             "restated ruptures 1.1.9), parse + step 1 + subset file + O(p^2) "
                                             "per-read rescans + step 2"},
            "port": {"seconds": round(t_port, 2), "gbases_per_s": round(n_bases / t_port / 1e9, 5), "trc_pass": n_pass,
                     "what": "oracle/cpu_baseline.py on the same reads in memory (no parsing, no rescans)"},
            "port_over_reference": round(t_ref / t_port, 2)}))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
