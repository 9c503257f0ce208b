#!/usr/bin/env python
"""Wall time of the `topsicle` CLI of this repo on a synthetic FASTQ file (config 2 reads), including
process start-up work (context creation, pinned allocations), parsing, the scan and every output file.
Usage: python tools/cli_e2e.py [n_reads] [extra CLI flags ...]"""
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from topsicle_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 63488
extra = sys.argv[2:]
shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
work = tempfile.mkdtemp(prefix="tps_cli_", dir=shm)
bases, off, _ = synth.generate(synth.CONFIGS[2], 0, n)
path = os.path.join(work, "reads.fastq")
synth.write_fastq(path, bases, off, prefix="syn2")
size = os.path.getsize(path)
for rep in range(2):
    out = os.path.join(work, f"out{rep}")
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "topsicle_b200.main", "-i", path, "-o", out, "--pattern", "CCCTAA",
                        "--minSeqLength", "9000"] + extra, cwd=REPO, capture_output=True, text=True,
                       env=dict(os.environ, TOPSICLE_TIMING="1"))
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    rows = sum(1 for _ in open(os.path.join(out, "telolengths_all.csv"))) - 1
    line = [ln for ln in r.stdout.splitlines() if "scanned" in ln][-1]
    print("\n".join("    " + ln for ln in r.stderr.strip().splitlines() if "[timing]" in ln))
    print(f"run {rep}: {n} reads, {int(off[-1]) / 1e9:.3f} Gbases, {size / 1e9:.2f} GB FASTQ -> {rows} CSV rows; process wall "
          f"{dt:.2f} s = {int(off[-1]) / dt / 1e9:.2f} Gbases/s; {line.split('] ', 1)[1]}")
subprocess.run(["rm", "-rf", work])
