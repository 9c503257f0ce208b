#!/usr/bin/env python
"""Throughput of plain-gzip FASTQ input (one deflate stream per file, what `gzip reads.fastq` writes and what the
reference opens with gzip.open, allsteps.py:142-146): the reader with its parallel inflater (csrc/tps_pgz.c), the
same reader on zlib (TPS_FX_NO_PGZ=1), and the uncompressed file, on synthetic config-2 reads with random
(poorly compressible) quality lines; and, when a GPU is present, the whole scan from the `.gz` file through
`pipeline.Scanner` against the scan of the plain file (rows must be identical).

  python tools/gz_throughput.py [--gbases 1.1] [--level 6] [--threads N]          -> JSON lines
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

import numpy as np  # noqa: E402


def read_file(path, threads, env=None):
    from topsicle_b200 import fastx
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        h, hl = hashlib.md5(), hashlib.md5()
        t0 = time.perf_counter()
        n_reads = n_bases = 0
        with fastx.FastxFile(path, threads=threads) as fx:
            bases = np.empty(1 << 28, dtype=np.uint8)
            starts = np.empty((1 << 17) + 1, dtype=np.uint64)
            lens = np.empty(1 << 17, dtype=np.uint32)
            while True:
                b = fx.next_spans(bases, starts, lens)
                if b is None:
                    break
                n_reads += b.n_reads
                n_bases += b.n_bases
                b.release()
            st = fx.inflate_stats()
        dt = time.perf_counter() - t0
        # second pass for the checksum (kept out of the timing): sequence text of every read, in order
        with fastx.FastxFile(path, threads=threads) as fx:
            bases = np.empty(1 << 28, dtype=np.uint8)
            offsets = np.empty((1 << 17) + 1, dtype=np.uint64)
            while True:
                b = fx.next_batch(bases, offsets)
                if b is None:
                    break
                # independent of where the reader cut its batches: all bases back to back, all read lengths
                h.update(bases[:int(offsets[b.n_reads])].tobytes())
                hl.update(np.diff(offsets[:b.n_reads + 1]).astype(np.uint64).tobytes())
                b.release()
        return dict(seconds=round(dt, 3), reads=n_reads, gbases=round(n_bases / 1e9, 3),
                    gbases_per_s=round(n_bases / dt / 1e9, 3), md5=h.hexdigest() + hl.hexdigest(), inflate=st)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gbases", type=float, default=1.1)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--no-zlib", action="store_true")
    a = ap.parse_args()
    from topsicle_b200 import synth
    spec = synth.CONFIGS[2]
    mean_len = float(synth.read_lengths(spec, 0, 4096)[-1]) / 4096
    n = int(a.gbases * 1e9 / mean_len)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    work = tempfile.mkdtemp(prefix="tps_gz_", dir=shm)
    try:
        bases, off, _ = synth.generate(spec, 0, n)
        rng = np.random.default_rng(3)
        plain = os.path.join(work, "reads.fastq")
        with open(plain, "wb") as fh:
            for i in range(n):
                s = bases[int(off[i]):int(off[i + 1])].tobytes()
                fh.write(b"@syn2_%d\n%s\n+\n%s\n" % (i, s, rng.integers(35, 74, len(s), dtype=np.uint8).tobytes()))
        t0 = time.perf_counter()
        subprocess.check_call(["gzip", f"-{a.level}", "-k", plain])
        gz = plain + ".gz"
        print(json.dumps({"reads": n, "gbases": round(int(off[-1]) / 1e9, 3), "fastq_gb": round(os.path.getsize(plain) / 1e9, 2),
                          "gz_gb": round(os.path.getsize(gz) / 1e9, 2), "gzip_level": a.level,
                          "gzip_seconds": round(time.perf_counter() - t0, 1), "threads": a.threads}), flush=True)
        res = {"plain": read_file(plain, a.threads), "gz_parallel": read_file(gz, a.threads)}
        if not a.no_zlib:
            res["gz_zlib"] = read_file(gz, a.threads, {"TPS_FX_NO_PGZ": "1"})
        res["identical"] = len({r["md5"] for r in res.values() if isinstance(r, dict)}) == 1
        if "gz_zlib" in res:
            res["speedup_over_zlib"] = round(res["gz_parallel"]["gbases_per_s"] / res["gz_zlib"]["gbases_per_s"], 1)
        print(json.dumps(res), flush=True)
        try:
            import torch
            have_gpu = torch.cuda.is_available()
        except Exception:  # noqa: BLE001
            have_gpu = False
        if have_gpu:
            from topsicle_b200 import pipeline
            from topsicle_b200.patterns import patterns_to_search
            cfg = pipeline.ScanConfig(patterns=patterns_to_search("CCCTAA", 4), len_telopattern=6, phrase=4)
            out = {}
            with pipeline.Scanner([cfg], devices=[0], threads=a.threads) as sc:
                for name, path in (("plain", plain), ("gz", gz)):
                    rows = []
                    sc.scan_file(path, lambda res: None)
                    t0 = time.perf_counter()
                    st = sc.scan_file(path, lambda res: rows.append(res.passes[0]))
                    dt = time.perf_counter() - t0
                    out[name] = dict(seconds=round(dt, 3), gbases_per_s=round(st.n_bases / dt / 1e9, 2),
                                     rows=[(p.index, p.read_id, p.tail, p.count, p.telo_length) for pb in rows for p in pb])
            print(json.dumps({"scan_from_plain_gbases_per_s": out["plain"]["gbases_per_s"],
                              "scan_from_gz_gbases_per_s": out["gz"]["gbases_per_s"],
                              "trc_pass_reads": len(out["gz"]["rows"]),
                              "rows_identical": out["gz"]["rows"] == out["plain"]["rows"]}), flush=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
