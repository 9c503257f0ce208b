#!/usr/bin/env python
"""The product path on the GPUs of one box: ONE process, `pipeline.Scanner(devices=[0..N-1])`, synthetic FASTQ
files of a BASELINE configuration in page cache (the reference's unit of parallelism is the file,
main.py:232-235; here every GPU takes batches of any file).  Config 5 = three pattern sets (TTAGGG, TTTAGGG,
AAACCCT), every file scanned under its own (one Scanner, a context per pattern set and device); its rows are
compared with three single-pattern, single-device runs.

  python tools/scanner_multi_gpu.py --config 4 --gpus 1 2 4 8 [--files-per-gpu 2] [--reads-per-file N]   -> JSON lines
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--gpus", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--files-per-gpu", type=int, default=2)
    ap.add_argument("--gbases-per-file", type=float, default=1.0)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--no-ends-first", action="store_true")
    a = ap.parse_args()
    import numpy as np
    import torch
    from topsicle_b200 import pipeline, synth
    from topsicle_b200.patterns import patterns_to_search
    have = torch.cuda.device_count()
    spec = synth.CONFIGS[a.config]
    cli = spec["cli"]
    motifs = list(spec.get("sub_batches") or [cli["pattern"]])
    nm = len(motifs)
    gmax = max(g for g in a.gpus if g <= have)
    n_files = max(nm, a.files_per_gpu * gmax) // nm * nm
    mean_len = float(synth.read_lengths(spec, 0, 4096)[-1]) / 4096
    reads_per_file = max(256, int(a.gbases_per_file * 1e9 / mean_len))

    def cfg_for(motif):
        cut = cli.get("cutoff", 0.7)
        k = (cli.get("telophrase") or [len(motif) - 2])[0]
        return pipeline.ScanConfig(patterns=patterns_to_search(motif, k), len_telopattern=len(motif), phrase=k,
                                   cutoff=min(cut) if isinstance(cut, list) else cut,
                                   min_seq_length=cli.get("minSeqLength", 9000), window_size=cli.get("windowSize", 100),
                                   slide=cli.get("slide") or len(motif), trimfirst=cli.get("trimfirst", 100),
                                   maxlengthtelo=cli.get("maxlengthtelo", 20000))

    cfgs = [cfg_for(m) for m in motifs]
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    work = tempfile.mkdtemp(prefix="tps_multi_", dir=shm)
    try:
        t0 = time.perf_counter()
        files, total_bases = [], 0
        for f in range(n_files):
            mi = f % nm
            mot = motifs[mi] if nm > 1 else None
            first = (f // nm) * reads_per_file
            off = synth.read_lengths(spec, first, reads_per_file, mot)
            bases = np.empty(int(off[-1]), dtype=np.uint8)
            synth.fill_reads(spec, first, off, bases, motif=mot)
            path = os.path.join(work, f"reads_{f}_{motifs[mi]}.fastq")
            synth.write_fastq(path, bases, off, prefix=f"syn{a.config}_{motifs[mi]}", first_read=first)
            files.append((path, mi))
            total_bases += int(off[-1])
        print(json.dumps({"generated": n_files, "reads_per_file": reads_per_file, "gbases": round(total_bases / 1e9, 2),
                          "fastq_gb": round(sum(os.path.getsize(p) for p, _ in files) / 1e9, 1),
                          "seconds": round(time.perf_counter() - t0, 1)}), flush=True)
        key = lambda pbs: [(p.index, p.read_id, p.tail, p.count, p.telo_length) for pb in pbs for p in pb]  # noqa: E731

        def run(devices, ends_first, use_files):
            got = {p: [] for p, _ in use_files}
            with pipeline.Scanner(cfgs, devices=devices, leaders=tuple(range(nm)), ends_first=ends_first,
                                  threads=len(os.sched_getaffinity(0))) as sc:
                def jobs():
                    return [pipeline.FileJob(p, (lambda res, p=p: got[p].append(res.passes[0])), cfg_ids=[mi])
                            for p, mi in use_files]
                sc.scan_files(jobs())                    # warm-up: page cache, first launches
                t = time.perf_counter()
                for _ in range(a.passes):
                    for g in got.values():
                        g.clear()
                    stats = sc.scan_files(jobs())
                dt = (time.perf_counter() - t) / a.passes
            return got, stats, dt

        ref_rows = None
        for n in [g for g in a.gpus if g <= have]:
            use = files[:max(nm, a.files_per_gpu * n) // nm * nm]
            nbases = None
            for ends_first in ([False] if a.no_ends_first else [False, True]):
                got, stats, dt = run(list(range(n)), ends_first, use)
                nbases = sum(st.n_bases for st in stats)
                rows = {p: key(g) for p, g in got.items()}
                if ref_rows is None:                     # first run (fewest devices, whole reads) is the yardstick
                    ref_rows = {}
                same = all(ref_rows.setdefault(p, r) == r for p, r in rows.items())
                print(json.dumps({"config": a.config, "workload": spec["name"], "gpus": n, "files": len(use),
                                  "patterns": motifs, "ends_first": ends_first, "gbases_per_pass": round(nbases / 1e9, 2),
                                  "seconds_per_pass": round(dt, 4), "gbases_per_s": round(nbases / dt / 1e9, 2),
                                  "trc_pass_reads": sum(len(pb) for g in got.values() for pb in g),
                                  "uploaded_fraction": round(sum(st.n_uploaded for st in stats) / max(1, nbases), 4),
                                  "rows_identical_to_first_run": same,
                                  "host_threads": len(os.sched_getaffinity(0))}), flush=True)
        if nm > 1:       # every pattern set on its own, one device: the rows one Scanner gave must be these
            ok = True
            for mi, mot in enumerate(motifs):
                mine = [(p, 0) for p, m in files if m == mi and p in ref_rows]
                with pipeline.Scanner([cfgs[mi]], devices=[0]) as sc:
                    for p, _ in mine:
                        rows = []
                        sc.scan_file(p, lambda res: rows.append(res.passes[0]))
                        ok = ok and key(rows) == ref_rows[p]
            print(json.dumps({"config": a.config, "per_pattern_rows_identical_to_single_pattern_runs": ok}), flush=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
