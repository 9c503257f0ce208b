#!/usr/bin/env python
"""Wall time of the `topsicle` CLI of this repo (a fresh process: start-up, context creation, pinned allocations,
parsing, the scan and every output file) on synthetic FASTQ input in page cache.
  python tools/cli_e2e.py [n_reads_per_file] [--files N] [--config C] [--reps R] [--both-modes] [extra CLI flags ...]
With --files N the input is a directory of N files (the reference's unit of parallelism, main.py:232-235)."""
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from topsicle_b200 import synth  # noqa: E402

argv = sys.argv[1:]


def take(flag, default, cast=int):
    if flag in argv:
        i = argv.index(flag)
        v = cast(argv[i + 1])
        del argv[i:i + 2]
        return v
    return default


n_files = take("--files", 1)
both = "--both-modes" in argv          # every run once without and once with --ends-first (input generated once)
if both:
    argv.remove("--both-modes")
config = take("--config", 2)
reps = take("--reps", 2)
n = int(argv.pop(0)) if argv and argv[0].isdigit() else 63488
extra = argv
spec = synth.CONFIGS[config]
cli = spec["cli"]
shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
work = tempfile.mkdtemp(prefix="tps_cli_", dir=shm)
indir = os.path.join(work, "in")
os.makedirs(indir)
total_bases = size = 0
for f in range(n_files):
    bases, off, _ = synth.generate(spec, f * n, n)
    path = os.path.join(indir, f"reads_{f}.fastq")
    synth.write_fastq(path, bases, off, prefix=f"syn{config}", first_read=f * n)
    total_bases += int(off[-1])
    size += os.path.getsize(path)
flags = ["--pattern", cli["pattern"], "--minSeqLength", str(cli.get("minSeqLength", 9000)),
         "--windowSize", str(cli.get("windowSize", 100)), "--maxlengthtelo", str(cli.get("maxlengthtelo", 20000))]
if cli.get("slide"):
    flags += ["--slide", str(cli["slide"])]
src = indir if n_files > 1 else os.path.join(indir, "reads_0.fastq")
runs = [(rep, e) for rep in range(reps) for e in ([[], ["--ends-first"]] if both else [[]])]
for rep, mode in runs:
    extra = [a for a in argv if a != "--ends-first"] + mode if both else argv
    out = os.path.join(work, f"out{rep}{'e' if mode else ''}")
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "topsicle_b200.main", "-i", src, "-o", out] + flags + extra, cwd=REPO,
                       capture_output=True, text=True, env=dict(os.environ, TOPSICLE_TIMING="1"))
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    rows = sum(1 for _ in open(os.path.join(out, "telolengths_all.csv"))) - 1
    line = [ln for ln in r.stdout.splitlines() if "scanned" in ln][-1]
    print("\n".join("    " + ln for ln in r.stderr.strip().splitlines() if "[timing]" in ln and "device" not in ln))
    print(f"run {rep}: {n_files} file(s) x {n} reads of {spec['name'].split(':')[0]}, {total_bases / 1e9:.3f} Gbases, "
          f"{size / 1e9:.2f} GB FASTQ, flags {' '.join(extra) or '-'} -> {rows} CSV rows; process wall {dt:.2f} s = "
          f"{total_bases / dt / 1e9:.2f} Gbases/s; {line.split('] ', 1)[1]}", flush=True)
subprocess.run(["rm", "-rf", work])
