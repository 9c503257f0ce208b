#!/usr/bin/env python
"""bench.py -- Gbases/s scanned by the B200 telomere scan (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N ...          (the reference's CPU path, rank 0 only)

A *step* is one pass of the hot path (K1 pack -> K2 TRC -> K3 windows -> K4 change point)
over one batch of synthetic reads of the named configuration (default: BASELINE.json
configs[1] = config 2, synth-v1 seed 2002).  Each rank owns its own reads (weak scaling);
there is no data-path collective.  `value` = bases scanned by all ranks / max-over-ranks
time with the batch resident in HBM; `e2e` = same metric through tps_submit/tps_wait with
pinned HOST buffers (H2D of bases+offsets and D2H of the result rows inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
# before anything loads libgomp (torch does): the FASTQ parser's OpenMP team shares the host cores with the
# Python threads and the other ranks, so idle team members sleep instead of spinning
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

import numpy as np  # noqa: E402

METRIC = "telomere_scan_throughput"
UNIT = "Gbases/s"
ALG_BYTES_PER_BASE = 1.25  # K1: 1 B ASCII read + 0.25 B 2-bit code written
# the dominant kernel: bulk-copy (TMA) staged pack kernel unless the register-staged one is forced
K1_KERNEL = "tps_pack_kernel" if os.environ.get("TPS_K1_TMA") == "0" else "tps_pack_tma_kernel"

# stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner under torchrun)
# are sent to stderr for the whole run, the line itself goes to the saved descriptor.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config number (2..5)")
    ap.add_argument("--reads-per-step", type=int, default=0, help="reads per batch (default: ~1.5 Gbases)")
    ap.add_argument("--distinct-batches", type=int, default=2)
    ap.add_argument("--cpu-sample-reads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=3,
                    help="kernel-only leg: issue consecutive (independent) batches on this many batch slots/streams")
    ap.add_argument("--no-parse", action="store_true", help="skip the FASTQ-file end-to-end leg")
    ap.add_argument("--parse-passes", type=int, default=2)
    return ap.parse_args()


def scan_kwargs(spec, motif=None):
    cli = spec["cli"]
    pattern = motif or cli["pattern"]
    phrases = cli.get("telophrase") or [len(pattern) - 2]
    cut = cli.get("cutoff", 0.7)
    return dict(pattern=pattern, phrase=phrases[0], cutoff=min(cut) if isinstance(cut, list) else cut,
                min_len=cli.get("minSeqLength", 9000), W=cli.get("windowSize", 100),
                slide=cli.get("slide") or len(pattern), trim=cli.get("trimfirst", 100),
                maxlen=cli.get("maxlengthtelo", 20000))


def default_reads_per_step(spec):
    from topsicle_b200 import synth
    off = synth.read_lengths(spec, 0, 4096)
    mean_len = float(off[-1]) / 4096
    return max(1024, int(1.5e9 / mean_len) // 1024 * 1024)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def hbm_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_base():
    """(dram bytes per base of the pack kernel K1, where that figure comes from): the committed
    ncu --set full capture of the same command, not a live measurement."""
    try:
        with open(os.path.join(REPO, "profiles", "k1_traffic.json")) as fh:
            d = json.load(fh)
            return float(d["dram_bytes_per_base"]), ("profiles/k1_traffic.json (ncu --set full, dram__bytes_read.sum + "
                                                     f"dram__bytes_write.sum per base; captured {d.get('captured', 'round 1')})")
    except Exception:
        return None, None


def run_cpu_baseline(config, reads, steps=1, warmup=0, first_read=0, motif=None, phrase=0, raw=False):
    cmd = [sys.executable, os.path.join(REPO, "oracle", "cpu_baseline.py"), "--config", str(config),
           "--reads", str(reads), "--steps", str(steps), "--warmup", str(warmup), "--first-read", str(first_read)]
    if motif:
        cmd += ["--motif", motif]
    if phrase:
        cmd += ["--phrase", str(phrase)]
    if raw:
        cmd.append("--raw")
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    return json.loads(out)


def reference_arm(a):
    """`--impl reference`: the reference's CPU path (oracle port: the reference is pure Python and
    cannot travel to the GPU box) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from topsicle_b200 import synth
    spec = synth.CONFIGS[a.config]
    cores = len(os.sched_getaffinity(0))
    # the repo arm's batch (same `config`) unless K + W passes over it would not end within a few minutes on
    # these cores (the port does ~900 reads/s/core on config 2); then a bounded sample, and the line says so
    reads = a.cpu_sample_reads or a.reads_per_step or default_reads_per_step(spec)
    budget = int(240.0 / max(1, a.steps + a.warmup) * 900 * cores)
    if not a.cpu_sample_reads and reads > budget:
        reads = max(1024, budget // 1024 * 1024)
    r = run_cpu_baseline(a.config, reads, steps=a.steps, warmup=a.warmup)
    t = sum(r["seconds"])
    value = r["bases"] * a.steps / t / 1e9
    sample = (f"{reads} reads ({r['bases'] / 1e9:.3f} Gbases, {r['n_pass']} TRC-pass) of {spec['name']} per step; "
              "in-memory reads, parsing and the reference's O(p^2) temp-file rescans not charged")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": t / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic (synth-v1)",
            "config": {"workload": spec["name"], "reads_per_step": reads, "bases_per_step": r["bases"]},
            "reads_per_s": reads * a.steps / t,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    a = parse_args()
    if a.impl == "reference":
        return reference_arm(a)

    import torch
    import torch.distributed as dist
    from topsicle_b200 import engine, synth
    from topsicle_b200.patterns import patterns_to_search

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    spec = synth.CONFIGS[a.config]
    # pattern sets of the workload: one, or (config 5) one per species sub-batch -- the reference takes one
    # --pattern per invocation (main.py:321), so every sub-batch is scanned under its own pattern set
    motifs = list(spec.get("sub_batches") or [spec["cli"]["pattern"]])
    nm = len(motifs)
    kws = [scan_kwargs(spec, m) for m in motifs]
    kw = kws[0]
    phrases_all = spec["cli"].get("telophrase") or [kw["phrase"]]
    want_raw_cfg = bool(spec["cli"].get("rawcountpattern"))
    reads_per_step = a.reads_per_step or default_reads_per_step(spec)
    nb = (max(1, a.distinct_batches, nm) + nm - 1) // nm * nm     # batch b holds reads of sub-batch b % nm

    def ctx_kwargs(k):
        return dict(len_telopattern=len(k["pattern"]), cutoff=k["cutoff"], min_seq_length=k["min_len"], window_size=k["W"],
                    slide=k["slide"], trimfirst=k["trim"], maxlengthtelo=k["maxlen"], device=local_rank)

    # CPU baseline first (separate process, no CUDA in it), rank 0 at N=1 only: the timed sample is the first
    # pattern set / telophrase; the other pattern sets and telophrases get a smaller untimed sample, for parity
    cpu = None
    cpu_runs = {}      # (motif index, phrase) -> (sample reads, port output)
    if world == 1 and not a.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        reg = max(0, kw["maxlen"] - kw["trim"])
        nw = (reg - kw["W"]) // kw["slide"] + 1 if reg >= kw["W"] else 0
        per_read = 1.0 / 900 + spec["f_telo"] * 0.2 * nw / 3301          # seconds per read and core of the port
        sample_reads = a.cpu_sample_reads or int(min(reads_per_step, 200000, max(2048, 25.0 * cores / per_read)))
        r = run_cpu_baseline(a.config, sample_reads, motif=motifs[0] if nm > 1 else None, raw=want_raw_cfg)
        cpu_runs[(0, kw["phrase"])] = (sample_reads, r)
        cpu = {"value": r["gbases_per_s"][0], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "reads_per_s": r["reads_per_s"][0],
               "sample": f"first {sample_reads} reads ({r['bases'] / 1e9:.3f} Gbases, {r['n_pass']} TRC-pass) of the "
                         f"workload, {r['seconds'][0]:.1f} s; in-memory reads (parsing not charged)"}
        extra = max(2048, sample_reads // 4)
        for mi in range(nm):
            for ph in (phrases_all if nm == 1 else [kws[mi]["phrase"]]):
                if (mi, ph) not in cpu_runs:
                    cpu_runs[(mi, ph)] = (extra, run_cpu_baseline(a.config, extra, motif=motifs[mi] if nm > 1 else None,
                                                                  phrase=ph, raw=want_raw_cfg))

    # ---- synthetic batches: rank r owns reads [r*nb/nm*R + ...) of every sub-batch -> pinned host + device copies
    # host placement: this rank's pinned buffers and parser threads on the CPUs local to its GPU
    from topsicle_b200 import numa
    cores_before = len(os.sched_getaffinity(0))
    near = None if os.environ.get("TOPSICLE_NO_NUMA") else numa.cpus_near_device(local_rank)
    if near and world > 1:
        os.sched_setaffinity(0, near)
    t_gen = time.perf_counter()
    host_bases, host_off, dev_bases, dev_off, nbases, first_reads = [], [], [], [], [], []
    for b in range(nb):
        mi = b % nm
        first = (rank * (nb // nm) + b // nm) * reads_per_step
        mot = motifs[mi] if nm > 1 else None
        off = synth.read_lengths(spec, first, reads_per_step, mot)
        n = int(off[-1])
        hb = engine.PinnedBuffer(n)
        synth.fill_reads(spec, first, off, hb.array, motif=mot)
        ho = engine.PinnedBuffer(off.nbytes)
        ho.array.view(np.uint64)[:] = off
        pad = (n + 2047) // 2048 * 2048
        db = torch.empty(pad, dtype=torch.uint8, device=dev)
        db[:n].copy_(torch.from_numpy(hb.array[:n]))
        do = torch.from_numpy(off.view(np.int64)).to(dev)
        host_bases.append(hb); host_off.append(ho); dev_bases.append(db); dev_off.append(do); nbases.append(n)
        first_reads.append(first)
    t_gen = time.perf_counter() - t_gen
    max_bases = max(nbases)
    n_streams = max(1, min(a.streams, 3))
    d_rows_s = [torch.empty(reads_per_step * 40, dtype=torch.uint8, device=dev) for _ in range(n_streams * nm)]

    ctxs = [engine.ScanContext(patterns_to_search(k["pattern"], k["phrase"]), n_slots=3, max_batch_reads=reads_per_step,
                               max_batch_bases=max_bases, **ctx_kwargs(k)) for k in kws]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- kernel-only: batch resident in HBM; every step's input (1.5 GB) is larger than the 126 MB L2
    def step_device(i):
        b = i % nb
        mi = b % nm
        sl = (i // nm) % n_streams
        ctxs[mi].scan_device(dev_bases[b].data_ptr(), dev_off[b].data_ptr(), reads_per_step, nbases[b],
                             d_rows_s[mi * n_streams + sl].data_ptr(), slot=sl)
        return mi * n_streams + sl

    for i in range(a.warmup):
        step_device(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = sum(c.kernel_launches() for c in ctxs)
    t0 = time.perf_counter()
    last_buf = 0
    for i in range(a.steps):
        last_buf = step_device(i)
    barrier()
    t1 = time.perf_counter()
    launches = sum(c.kernel_launches() for c in ctxs) - launches0
    # The K steps as the DEVICE saw them: CUDA events of the library's own streams (the scans run on the slot /
    # pack / tail streams, which torch events do not see), from the first kernel of the first timed step to the end
    # of whichever of the last steps finished last.  The host clock around the two barriers also holds the latency
    # of the closing barrier (a NCCL all-reduce at N > 1), which is 5-25 % of a 20-step region of 6 ms.
    step_ctx = [(i % nb) % nm for i in range(a.steps)]
    per_ctx = [step_ctx.count(m) for m in range(nm)]
    dev_span = None
    if max(per_ctx) <= 250:

        def back_of(i):     # scan_device calls of step i's context issued after step i
            return sum(1 for j in range(i + 1, a.steps) if step_ctx[j] == step_ctx[i])
        c0 = ctxs[step_ctx[0]]
        dev_span = max(c0.elapsed_to(back_of(0), 0, ctxs[step_ctx[i]], back_of(i), 3)
                       for i in range(max(0, a.steps - n_streams * nm), a.steps)) * 1e-3
    rows_dev = np.frombuffer(d_rows_s[last_buf].cpu().numpy().tobytes(), dtype=engine.ROW_DTYPE).copy()
    # device-side (CUDA event) times per kernel, from the library's event ring.  With several streams
    # the kernels of consecutive batches overlap and per-kernel event times are not attributable, so the
    # per-kernel figures (and the K1 roofline) come from an extra single-stream pass over the same batches.
    n_attr = max(nm, min(a.steps, 16) // nm * nm)
    barrier()
    for i in range(n_attr):
        b = i % nb
        ctxs[b % nm].scan_device(dev_bases[b].data_ptr(), dev_off[b].data_ptr(), reads_per_step, nbases[b],
                                 d_rows_s[(b % nm) * n_streams].data_ptr(), slot=0)
        ctxs[b % nm].sync()
    barrier()
    ring = [c.timings(back) for c in ctxs for back in range(n_attr // nm)]
    k1_ms = statistics.mean(t["k1_pack"] for t in ring)
    k2_ms = statistics.mean(t["k2_trc"] for t in ring)
    k3_ms = statistics.mean(t["k3_windows_cp"] for t in ring)
    dev_ms = statistics.mean(t["total"] for t in ring)
    bases_timed = sum(nbases[i % nb] for i in range(a.steps))
    bases_attr = sum(nbases[i % nb] for i in range(n_attr)) / n_attr
    host_wall = max_over_ranks(t1 - t0)
    wall = max_over_ranks(dev_span) if dev_span is not None else host_wall
    total_bases = sum_over_ranks(float(bases_timed))
    value = total_bases / wall / 1e9

    # ---- parity at bench size: the CPU port's rows (float64 ruptures, as the reference computes them) for the
    # sampled reads against the GPU rows of the same reads: every pattern set, every telophrase, and (config 3) the
    # md5 of every TRC-pass read's raw-count table
    parity_all = []
    if cpu_runs and rank == 0:
        import hashlib
        for (mi, ph), (n_cmp, r) in sorted(cpu_runs.items()):
            k = dict(kws[mi], phrase=ph)
            n_cmp = min(n_cmp, reads_per_step)
            off = host_off[mi].array.view(np.uint64)[:n_cmp + 1]
            hb = host_bases[mi].array[:int(off[-1])]
            with engine.ScanContext(patterns_to_search(k["pattern"], ph), n_slots=1, max_batch_reads=n_cmp,
                                    max_batch_bases=int(off[-1]) + 4096, want_rawcount=want_raw_cfg,
                                    max_pass_reads=min(n_cmp, 16384), **ctx_kwargs(k)) as pc:
                rows0, raw0 = pc.scan(hb, np.ascontiguousarray(off))
                want = {gi: (tail, trc, telo, dig) for gi, tail, trc, telo, dig in r["pass_rows"] if gi < n_cmp}
                got = {}
                for i in np.nonzero(rows0["status"] >= engine.ST_PASS)[0]:
                    dig = None
                    if want_raw_cfg:
                        tab = pc.rawcount_table(rows0, raw0, int(i))
                        dig = hashlib.md5(tab.tobytes()).hexdigest() if tab is not None else None
                    got[int(i)] = (engine.TAIL_NAMES[int(rows0["tail"][i])],
                                   engine.trc_value(int(rows0["match_count"][i]), len(k["pattern"])),
                                   int(rows0["telo_length"][i]), dig)
            both = sorted(set(want) & set(got))
            diffs = [abs(want[i][2] - got[i][2]) for i in both]
            parity_all.append({
                "pattern": k["pattern"], "telophrase": ph, "reads_compared": int(n_cmp), "trc_pass_cpu": len(want),
                "trc_pass_gpu": len(got), "pass_sets_identical": set(want) == set(got),
                "tail_and_trc_identical": all(want[i][:2] == got[i][:2] for i in both),
                "telo_length_exact": sum(1 for d in diffs if d == 0), "telo_length_max_abs_diff": max(diffs, default=0),
                "rawcount_tables_identical": (sum(1 for i in both if want[i][3] == got[i][3]) if want_raw_cfg else None),
                "note": "CPU side = float64 numpy.var argmax (ruptures restatement); GPU side = exact rational argmax"})
    parity = parity_all[0] if parity_all else None

    # ---- end to end: pinned host buffers -> tps_submit / tps_wait (3 slots in flight per pattern set)
    e2e = None
    if not a.no_e2e:
        depth = 3

        def run_e2e(nsteps):
            pending, last = [], None
            for i in range(nsteps):
                b = i % nb
                c = ctxs[b % nm]
                pending.append((c, c.submit(host_bases[b].array[:nbases[b]], host_off[b].array.view(np.uint64))))
                if len(pending) >= depth:
                    c0, bid = pending.pop(0)
                    last = c0.wait(bid)
            while pending:
                c0, bid = pending.pop(0)
                last = c0.wait(bid)
            return last

        run_e2e(max(a.warmup, 1))
        barrier()
        t2 = time.perf_counter()
        last = run_e2e(a.steps)
        barrier()
        t3 = time.perf_counter()
        e_wall = max_over_ranks(t3 - t2)
        rows_e2e = last[0]
        # both loops end on batch (steps-1) % nb: the two paths must produce identical rows
        assert rows_e2e.tobytes() == rows_dev.tobytes(), "host-buffer path and device path disagree"
        e2e = {"value": total_bases / e_wall / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(max_bases + (reads_per_step + 1) * 8),
               "d2h_bytes_per_step": int(reads_per_step * 40 + 32),
               "ms_per_step": e_wall / a.steps * 1e3, "slots_in_flight": depth}
    clocks = sampler.stop(t0, time.perf_counter())

    # ---- end to end from FASTQ files: C reader (host threads) -> pinned batches -> H2D -> kernels -> D2H
    # -> harvest of the TRC-pass reads, through the same Scanner the `topsicle` CLI uses.  One file per pattern
    # set, each scanned under its own set (config 5); every telophrase + raw-count tables from one pass (config 3)
    e2e_file = None
    if not a.no_parse:
        from topsicle_b200 import pipeline
        for c in ctxs:
            c.close()
        shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
        paths = [os.path.join(shm, f"tps_bench_{os.getpid()}_r{rank}_{mi}.fastq") for mi in range(nm)]
        try:
            for mi, path in enumerate(paths):
                synth.write_fastq(path, host_bases[mi].array[:nbases[mi]], host_off[mi].array.view(np.uint64),
                                  prefix=f"syn{a.config}", first_read=first_reads[mi])
            fsize = sum(os.path.getsize(p) for p in paths)
            if nm == 1:
                cfgs = [pipeline.ScanConfig(patterns=patterns_to_search(kw["pattern"], k), len_telopattern=len(kw["pattern"]),
                                            phrase=k, cutoff=kw["cutoff"], min_seq_length=kw["min_len"],
                                            window_size=kw["W"], slide=kw["slide"], trimfirst=kw["trim"],
                                            maxlengthtelo=kw["maxlen"], want_rawcount=want_raw_cfg)
                        for k in phrases_all]
                groups, leaders = [None], (0,)
            else:
                cfgs = [pipeline.ScanConfig(patterns=patterns_to_search(k["pattern"], k["phrase"]),
                                            len_telopattern=len(k["pattern"]), phrase=k["phrase"], cutoff=k["cutoff"],
                                            min_seq_length=k["min_len"], window_size=k["W"], slide=k["slide"],
                                            trimfirst=k["trim"], maxlengthtelo=k["maxlen"]) for k in kws]
                groups, leaders = [[mi] for mi in range(nm)], tuple(range(nm))
            # ranks share the host cores (a rank bound to its GPU's CPUs already has its share)
            host_threads = max(1, cores_before // world)

            def timed_passes(ends_first):
                got = [[] for _ in paths]
                with pipeline.Scanner(cfgs, devices=[local_rank], max_batch_bases=1 << 28, max_batch_reads=1 << 17,
                                      depth=3, threads=host_threads, ends_first=ends_first, leaders=leaders) as sc:
                    def jobs():
                        # the sink keeps the batch's columnar result (PassBatch); rows are compared after the clock stops
                        return [pipeline.FileJob(p, (lambda res, j=j: got[j].append(res.passes[0])), cfg_ids=groups[j])
                                for j, p in enumerate(paths)]
                    sc.scan_files(jobs(), readers=1)           # warm-up pass (page cache, first launches)
                    barrier()
                    ta = time.perf_counter()
                    for _ in range(a.parse_passes):
                        for g in got:
                            g.clear()
                        stats = sc.scan_files(jobs(), readers=1)
                    barrier()
                    tb = time.perf_counter()
                return got, stats, max_over_ranks(tb - ta)

            got, stats, p_wall = timed_passes(False)
            n_bases_pass = sum(st.n_bases for st in stats)
            p_bases = sum_over_ranks(float(n_bases_pass * a.parse_passes))
            assert sum(st.n_reads for st in stats) == reads_per_step * nm and n_bases_pass == sum(nbases[:nm])
            # the same files through the ends-first mode (reported separately: the interior of the reads is
            # neither copied nor uploaded nor packed; B_alg' = the bytes actually touched)
            got2, stats2, q_wall = timed_passes(True)
            key = lambda gs: [[(p.index, p.tail, p.count, p.telo_length) for pb in g for p in pb] for g in gs]  # noqa: E731
            timing = lambda sts: {k: round(sum(st.timing[k] for st in sts), 4) for k in sts[0].timing}  # noqa: E731
            e2e_ends = {"value": p_bases / q_wall / 1e9, "unit": UNIT, "ms_per_pass": q_wall / a.parse_passes * 1e3,
                        "uploaded_bases_per_pass": int(sum(st.n_uploaded for st in stats2)),
                        "bases_per_pass": int(n_bases_pass),
                        "uploaded_fraction": sum(st.n_uploaded for st in stats2) / max(1, n_bases_pass),
                        "rows_identical_to_whole_read_scan": key(got2) == key(got),
                        "host_seconds_last_pass": timing(stats2),
                        "what": "same files, Scanner(ends_first=True): head + tail of every read uploaded and "
                                "scanned (K1 + K2), then the regions of the TRC-pass reads (K1..K4); reported "
                                "separately from e2e_from_fastq because the interior of the reads is never touched"}
            e2e_file = {"value": p_bases / p_wall / 1e9, "unit": UNIT, "file_bytes": fsize, "files": len(paths),
                        "passes": a.parse_passes, "ms_per_pass": p_wall / a.parse_passes * 1e3,
                        "host_threads_per_rank": host_threads, "trc_pass_reads": sum(len(pb) for g in got for pb in g),
                        "patterns": motifs, "telophrases": phrases_all if nm == 1 else [k["phrase"] for k in kws],
                        "rawcount_tables": want_raw_cfg, "host_seconds_last_pass": timing(stats),
                        "what": "uncompressed FASTQ in page cache -> telomere rows (parse + PCIe + kernels + harvest)",
                        "ends_first": e2e_ends}
        finally:
            for path in paths:
                if os.path.exists(path):
                    os.remove(path)

    n_pass = int((rows_dev["status"] >= engine.ST_PASS).sum())
    peak, peak_src = hbm_peak()
    alg_bytes = ALG_BYTES_PER_BASE * bases_attr
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    tpb, tpb_src = ncu_traffic_per_base()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": wall / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic (synth-v1, per-read xoshiro streams; see topsicle_b200/synth.py)",
                "config": {"workload": spec["name"], "reads_per_step": reads_per_step,
                           "bases_per_step": int(bases_timed / a.steps), "distinct_batches": nb, "streams": n_streams,
                           "pattern": kw["pattern"] if nm == 1 else motifs,
                           "telophrase": kw["phrase"] if nm == 1 else [k["phrase"] for k in kws], "cutoff": kw["cutoff"],
                           "minSeqLength": kw["min_len"], "windowSize": kw["W"],
                           "slide": kw["slide"] if nm == 1 else [k["slide"] for k in kws],
                           "trimfirst": kw["trim"], "maxlengthtelo": kw["maxlen"],
                           "l2_policy": "each step reads a batch (>1 GB) far larger than the 126 MB L2",
                           "parallelism": f"reads sharded over {world} GPU(s), no collective"},
                "reads_per_s": total_bases / wall / (bases_timed / a.steps / reads_per_step),
                "trc_pass_reads_per_step": n_pass,
                "device_ms_per_step": {"k1_pack": k1_ms, "k2_trc": k2_ms, "k3k4_windows_changepoint": k3_ms,
                                       "scan_total": dev_ms},
                "roofline": {"bound": "hbm", "kernel": K1_KERNEL, "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                             "algorithmic_bytes_per_base": ALG_BYTES_PER_BASE,
                             "traffic": (tpb * bases_timed / a.steps) if tpb else None, "traffic_source": tpb_src,
                             "whole_scan_frac": alg_bytes / (dev_ms * 1e-3) / 1e9 / peak,
                             "pipelined_scan_frac": ALG_BYTES_PER_BASE * value / peak / world},
                "cpu_baseline": cpu, "parity_sample": parity, "parity_all": parity_all or None, "e2e": e2e,
                "e2e_from_fastq": e2e_file, "gpu_launches": int(launches),
                "timing": {"method": ("CUDA events on the library's streams: first kernel of the first timed step -> end "
                                      "of the last steps, between the two barriers; max over ranks") if dev_span is not None
                           else "host clock between the two barriers (more timed scans than the event ring holds)",
                           "device_span_ms": wall * 1e3, "host_wall_between_barriers_ms": host_wall * 1e3},
                "clocks": clocks,
                "generate_s": t_gen,
                "host_placement": {"cpus_local_to_gpu0": near, "bound": bool(near and world > 1)}}
        emit(line)
    for c in ctxs:
        c.close()
    for hb in host_bases + host_off:
        hb.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
