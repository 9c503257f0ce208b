#!/usr/bin/env python
"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the reference's CPU path, timed.

A port of what `Topsicle/main.py:process_file` makes the CPU do for one file, operation
for operation, on reads already in memory (so the reference's Biopython parsing and its
O(p^2) temp-file re-scans, main.py:83-86 / allsteps.py:252-259, are NOT charged -- this
favours the reference):

  step 1  allsteps.py:175-198  `re.finditer` of every literal over seq[:1000] and the
          reversed seq[-1000:]
  step 2  allsteps.py:263-297  windows of BOTH orientations, `len(finditer) or 1` per
          literal, mean (one orientation is then discarded, exactly as upstream)
          allsteps.py:310-311  ruptures Binseg(l2).fit(y).predict(pen=4, n_bkps=1)
          (restated 1.1.9, oracle/shims/ruptures)

Parallelism: the reference runs one process per input FILE (main.py:232-235); here the
sample is dealt into `cores` equal-base shards, one worker process each, i.e. the best
case for the reference.  Used by `bench.py` (cpu_baseline leg and `--impl reference`).
Run as a script it prints one JSON object.
"""
import argparse
import json
import os
import re
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, os.path.join(HERE, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def _windows(s, window_size, step):
    """allsteps.py:207-225."""
    out = []
    for i in range(0, len(s) - window_size + 1, step):
        end = i + window_size - 1
        if end > len(s):
            end = len(s)
        out.append((i, s[i:end]))
    return out


def scan_shard(args):
    """Reference-style scan of a list of (index, read string) -> (n_scanned, n_pass, [(index, tail, count, telo,
    md5 of the chosen tail's uint8 count table [window][literal] or None)])."""
    import hashlib
    import ruptures as rpt  # restated 1.1.9 (oracle/shims)
    from oracle.topsicle_oracle import patterns_to_search
    seqs, pattern, phrase, cutoff, min_len, W, slide, trim, maxlen, want_raw = args
    literals = patterns_to_search(pattern, phrase)
    compiled = [re.compile(p) for p in literals]
    ratio = 1000 / len(pattern)
    passing = []
    n_scanned = 0
    for gi, seq in seqs:                               # ---- step 1, allsteps.py:174-198
        if len(seq) > min_len:
            n_scanned += 1
            head = seq[:1000].upper()
            tail = seq[-1000:][::-1].upper()
            rows = []
            for pat in compiled:
                ms = len([m.start() for m in pat.finditer(head)])
                me = len([m.start() for m in pat.finditer(tail)])
                rows.append((ms / ratio, me / ratio))
            best_s = max(r[0] for r in rows)
            best_e = max(r[1] for r in rows)
            if best_s > best_e:
                if best_s > cutoff:
                    passing.append((gi, seq, "forward", best_s))
            elif best_e > cutoff:
                passing.append((gi, seq, "reverse", best_e))
    telo = []
    for gi, seq, tail, trc in passing:                 # ---- step 2, allsteps.py:263-315
        m = min(maxlen, len(seq))
        s_fwd = seq[trim:m].upper()
        s_rev = seq[::-1].upper()[trim:m]
        mean_s, mean_e = [], []
        tab_s, tab_e = [], []
        for start, cut in _windows(s_fwd, W, slide):
            c = [len([mm.start() for mm in pat.finditer(cut)]) or 1 for pat in compiled]
            mean_s.append((start, sum(c) / len(c)))
            tab_s.append(c)
        for start, cut in _windows(s_rev, W, slide):
            c = [len([mm.start() for mm in pat.finditer(cut)]) or 1 for pat in compiled]
            mean_e.append((start, sum(c) / len(c)))
            tab_e.append(c)
        mean = mean_s if tail == "forward" else mean_e
        digest = None
        if want_raw:     # rawCountPattern's table of the chosen tail (allsteps.py:401-416), as the kernels return it
            tab = tab_s if tail == "forward" else tab_e
            digest = hashlib.md5(np.asarray(tab, dtype=np.uint8).tobytes()).hexdigest()
        x = [a + trim for a, _ in mean]
        y = [b for _, b in mean]
        if len(y) < 7:
            telo.append((gi, tail, trc, -1, digest))
            continue
        res = rpt.Binseg(model="l2").fit(np.array(y)).predict(pen=4, n_bkps=1)
        telo.append((gi, tail, trc, int(x[res[0]]), digest))
    return n_scanned, len(passing), telo


def deal_shards(seqs, n):
    """n shards of (nearly) equal bases, reads kept in order inside a shard."""
    shards = [[] for _ in range(n)]
    load = [0] * n
    for gi, s in enumerate(seqs):
        i = load.index(min(load))
        shards[i].append((gi, s))
        load[i] += len(s)
    return shards


def time_sample(seqs, scan_kw, cores, steps=1, warmup=0):
    """Wall time per pass over `seqs` on `cores` worker processes."""
    import multiprocessing as mp
    shards = deal_shards(seqs, cores)
    argl = [(sh, scan_kw["pattern"], scan_kw["phrase"], scan_kw["cutoff"], scan_kw["min_len"], scan_kw["W"],
             scan_kw["slide"], scan_kw["trim"], scan_kw["maxlen"], scan_kw.get("want_raw", False)) for sh in shards]
    times, last = [], None
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=cores) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            last = pool.map(scan_shard, argl)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    n_pass = sum(r[1] for r in last)
    rows = sorted(t for r in last for t in r[2])
    return times, n_pass, rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--reads", type=int, default=4000)
    ap.add_argument("--first-read", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--cores", type=int, default=0)
    ap.add_argument("--motif", default="", help="pattern of this sub-batch (config 5: TTAGGG / TTTAGGG / AAACCCT)")
    ap.add_argument("--phrase", type=int, default=0, help="telophrase (default: the configuration's first)")
    ap.add_argument("--raw", action="store_true", help="also return the md5 of every TRC-pass read's count table")
    a = ap.parse_args()
    from topsicle_b200 import synth
    spec = synth.CONFIGS[a.config]
    cli = dict(spec["cli"])
    if a.motif:
        cli["pattern"] = a.motif
    bases, off, _ = synth.generate(spec, a.first_read, a.reads, motif=a.motif or None)
    buf = bases.tobytes().decode("ascii")
    seqs = [buf[int(off[i]):int(off[i + 1])] for i in range(a.reads)]
    cores = a.cores or len(os.sched_getaffinity(0))
    pattern = cli["pattern"]
    phrases = cli.get("telophrase") or [len(pattern) - 2]
    cut = cli.get("cutoff", 0.7)
    kw = dict(pattern=pattern, phrase=a.phrase or phrases[0], want_raw=a.raw,
              cutoff=min(cut) if isinstance(cut, list) else cut,
              min_len=cli.get("minSeqLength", 9000), W=cli.get("windowSize", 100),
              slide=cli.get("slide") or len(pattern), trim=cli.get("trimfirst", 100),
              maxlen=cli.get("maxlengthtelo", 20000))
    times, n_pass, rows = time_sample(seqs, kw, cores, a.steps, a.warmup)
    n_bases = int(off[-1])
    print(json.dumps(dict(config=a.config, reads=a.reads, bases=n_bases, cores=cores, n_pass=n_pass,
                          seconds=times, gbases_per_s=[n_bases / t / 1e9 for t in times],
                          reads_per_s=[a.reads / t for t in times],
                          pattern=pattern, phrase=kw["phrase"],
                          pass_rows=[[gi + a.first_read, tail, trc, telo, dig] for gi, tail, trc, telo, dig in rows])))


if __name__ == "__main__":
    main()
