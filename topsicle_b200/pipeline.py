"""Batched per-file scan driver: FASTQ/FASTA reader -> pinned batches -> GPU(s) -> ordered results.

This is what `process_file` (Topsicle/main.py:52-154) becomes on a B200 box.  The reference
parses the file (at least) twice plus once per TRC-pass read and runs step 1 / step 2 read
by read in one Python process per *file*; here one pass over the file feeds fixed-size
batches of reads to every visible GPU:

    reader (C, threads) --> pinned batch --tps_submit--> [H2D | K1 K2 K3 K4 | D2H] --tps_wait-->
    harvest TRC-pass reads --> ordered sink (file order restored across devices)

One host thread per device owns a `ScanContext` per telophrase and a ring of pinned batch
buffers; batches are dealt to whichever device asks next (reads are independent units, no
collective).  Several phrases (`--telophrase 4 5 6`) are scanned from ONE parse of the file.
Results are handed to the caller's sink strictly in file order, whatever order the devices
finish in, so CSV rows, `rawcount_{phrase}_{i}.csv` numbering and the subset file come out
exactly as the reference writes them.

Ends-first mode (`Scanner(ends_first=True)`, CLI `--ends-first`): step 1 only looks at the first and last
`no_bp` bases of a read and steps 2/3 only at one end of the TRC-pass reads, so the reader copies just
head + tail of every read (`tps_fastx_next_ends`), the device runs K1 + K2 on those (`tps_submit_ends`),
and the regions of the few TRC-pass reads are taken from the file text afterwards and scanned as a second,
small batch (`tps_submit_regions`).  Same rows, same files; about a tenth of the bytes cross PCIe.
"""
from __future__ import annotations

import os
import queue
import sys
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np

from . import engine, fastx, numa


@dataclass
class ScanConfig:
    """The arguments main.py:process_file passes down for one telophrase (main.py:57,129-130,147-148)."""
    patterns: Sequence[str]
    len_telopattern: int
    phrase: int = 0
    cutoff: float = 0.7
    min_seq_length: int = 9000
    no_bp: int = 1000
    window_size: int = 100
    slide: int = 6
    trimfirst: int = 100
    maxlengthtelo: int = 20000
    want_rawcount: bool = False
    step1_only: bool = False
    force_tail: str | None = None
    count_threshold_override: int | None = None


@dataclass
class PassRead:
    """One TRC-pass read (a row of telolengths_all.csv)."""
    index: int            # position of the read in its file
    read_id: str
    literal: str          # first-max literal (allsteps.py:190-191)
    tail: str             # 'forward' | 'reverse'
    count: int
    trc: float            # count / (no_bp / len(pattern))
    status: int           # engine.ST_PASS or engine.ST_BADSEG
    n_windows: int
    telo_length: int      # trimfirst + slide * bkp, -1 if BADSEG
    length: int
    counts: np.ndarray | None = None   # uint8 [n_windows][n_patterns] when want_rawcount
    record: bytes | memoryview | None = None   # SeqIO.write text when want_records


class PassBatch:
    """The TRC-pass reads of one batch under one config, as columns: what a device hands back is arrays, and a
    batch of a telomere-enriched library holds thousands of passing reads, so the rows stay arrays until somebody
    asks for one read.  Behaves as a sequence of PassRead (len, iteration, indexing) built on demand.  The raw-count
    tables are views of one array per batch: `table(j)` hands out the view (valid until the sink returns, when the
    pipeline gives the landing buffer back), a PassRead owns a copy.

    Columns (numpy, length n): index (position of the read in its file), count, literal_idx, tail_code, status,
    n_windows, telo_length, length; `ids` (list of str), `trc` (float64 = count / (no_bp / len(pattern)),
    allsteps.py:178-186), `records` (SeqIO.write text per read when asked for, else None)."""

    __slots__ = ("cfg", "patterns", "index", "ids", "count", "literal_idx", "tail_code", "status", "n_windows",
                 "telo_length", "length", "trc", "records", "_raw", "_raw_off", "_tables", "_leases")

    def __init__(self, cfg, patterns, index, ids, sel, raw=None, raw_off=None, tables=None, records=None, leases=()):
        self.cfg, self.patterns = cfg, patterns
        self.index, self.ids = index, ids
        self.count = sel["match_count"]
        self.literal_idx = sel["best_pattern"]
        self.tail_code = sel["tail"]
        self.status = sel["status"]
        self.n_windows = sel["n_windows"]
        self.telo_length = sel["telo_length"]
        self.length = sel["length"]
        self.trc = self.count / (cfg.no_bp / cfg.len_telopattern)
        self.records = records
        self._raw, self._raw_off, self._tables = raw, raw_off, tables
        self._leases = list(leases)

    def __len__(self):
        return len(self.index)

    def release(self):
        """Give the page-locked landing buffers behind the raw-count tables back to their contexts.  The pipeline
        calls this when the sink has returned: a sink that keeps tables beyond its call copies them."""
        for ls in self._leases:
            ls.release()
        self._leases = []
        if self._raw is not None or self._tables is not None:
            self._raw = self._tables = None

    def table(self, j):
        """uint8 counts[n_windows][n_patterns] of the j-th passing read, or None."""
        if self._tables is not None:
            return self._tables[j]
        if self._raw is None or int(self._raw_off[j]) == engine.NO_RAWCOUNT:
            return None
        nw, npat, off = int(self.n_windows[j]), len(self.patterns), int(self._raw_off[j])
        return self._raw[off:off + nw * npat].reshape(nw, npat)

    def _own_table(self, j):
        """table(j) for a PassRead, which may outlive the sink call: a copy when the batch's tables are views of
        leased landing buffers."""
        t = self.table(j)
        return t.copy() if (t is not None and self._leases) else t

    def __getitem__(self, j):
        if isinstance(j, slice):
            return [self[i] for i in range(*j.indices(len(self)))]
        if j < 0:
            j += len(self)
        return PassRead(index=int(self.index[j]), read_id=self.ids[j], literal=self.patterns[int(self.literal_idx[j])],
                        tail=engine.TAIL_NAMES[int(self.tail_code[j])], count=int(self.count[j]), trc=float(self.trc[j]),
                        status=int(self.status[j]), n_windows=int(self.n_windows[j]),
                        telo_length=int(self.telo_length[j]), length=int(self.length[j]),
                        counts=self._own_table(j) if self.cfg.want_rawcount else None,
                        record=self.records[j] if self.records is not None else None)

    def __iter__(self):
        return (self[j] for j in range(len(self)))


@dataclass
class BatchResult:
    seq: int
    first_read: int
    n_reads: int
    n_bases: int
    n_scanned: int                     # reads with L > minSeqLength
    passes: list = field(default_factory=list)   # per config of the job: PassBatch (a sequence of PassRead)
    n_uploaded: int = 0                # bases that crossed PCIe for this batch (ends-first: ends + regions)

    def release(self):
        for p in self.passes:
            if hasattr(p, "release"):
                p.release()


@dataclass
class FileStats:
    n_reads: int = 0
    n_bases: int = 0
    n_scanned: int = 0
    n_batches: int = 0
    n_uploaded: int = 0                # bases that crossed PCIe (== n_bases unless ends-first)
    format_name: str = ""
    timing: dict = field(default_factory=dict)   # seconds: open, parse (reader thread), submit, wait, harvest, close


class _OrderedSink:
    """Delivers BatchResults to `fn` in batch order."""

    def __init__(self, fn):
        self.fn = fn
        self.lock = threading.Lock()
        self.next_seq = 0
        self.held = {}

    def put(self, res: BatchResult):
        with self.lock:
            self.held[res.seq] = res
            while self.next_seq in self.held:
                res = self.held.pop(self.next_seq)
                try:
                    self.fn(res)
                finally:
                    res.release()      # the raw-count tables were views of leased landing buffers
                self.next_seq += 1


class _Slot:
    def __init__(self, max_bases, max_reads):
        self.bases = engine.PinnedBuffer(max_bases)
        self.offsets = engine.PinnedBuffer((max_reads + 1) * 8)     # read starts (uint64)
        self.lens = engine.PinnedBuffer(max_reads * 4)              # read lengths (uint32)
        self.true_lens = engine.PinnedBuffer(max_reads * 4)         # ends batches: real read lengths (uint32)
        self.recs = np.empty(max_reads, dtype=fastx.REC_DTYPE)

    def free(self):
        self.bases.free()
        self.offsets.free()
        self.lens.free()
        self.true_lens.free()


def make_context(cfg: ScanConfig, device: int, max_batch_reads: int, max_batch_bases: int, n_slots: int,
                 max_pass_reads: int = 0, rawcount_capacity: int = 0) -> engine.ScanContext:
    return engine.ScanContext(
        cfg.patterns, len_telopattern=cfg.len_telopattern, cutoff=cfg.cutoff, min_seq_length=cfg.min_seq_length,
        no_bp=cfg.no_bp, window_size=cfg.window_size, slide=cfg.slide, trimfirst=cfg.trimfirst,
        maxlengthtelo=cfg.maxlengthtelo, want_rawcount=cfg.want_rawcount, device=device, n_slots=n_slots,
        max_batch_reads=max_batch_reads, max_batch_bases=max_batch_bases, max_pass_reads=max_pass_reads,
        rawcount_capacity=rawcount_capacity, count_threshold_override=cfg.count_threshold_override,
        step1_only=cfg.step1_only, force_tail=cfg.force_tail)


def windows_per_read(cfg: ScanConfig) -> int:
    reg = max(0, cfg.maxlengthtelo - cfg.trimfirst)
    return (reg - cfg.window_size) // cfg.slide + 1 if reg >= cfg.window_size else 0


def harvest(cfg: ScanConfig, ctx, batch, rows, raw, want_records: bool, keep=None, tables=None,
            leases=()) -> PassBatch:
    """TRC-pass reads of one finished batch, in read order, as one PassBatch (no per-read Python work: a handful of
    array operations and two C calls per batch).  `tables` (ends-first mode) maps a read index to its raw-count table
    instead of `raw` + the rows' offsets."""
    idx = np.nonzero(rows["status"] >= engine.ST_PASS)[0]
    ids = batch.read_ids(idx) if len(idx) else []   # one C call for the ids of all TRC-pass reads of the batch
    if keep is not None and len(idx):
        sel_keep = np.fromiter((rid in keep for rid in ids), dtype=bool, count=len(ids))
        idx = idx[sel_keep]
        ids = [rid for rid, k in zip(ids, sel_keep) if k]
    sel = rows[idx]
    raw_copy = raw_off = tabs = None
    if cfg.want_rawcount and tables is not None:
        tabs = [tables.get(int(i)) for i in idx]
    elif cfg.want_rawcount and raw is not None:
        # the tables are views of `raw`: a leased landing buffer (no copy), else one copy out of the context's own
        raw_copy = raw if leases else np.array(raw, copy=True)
        raw_off = sel["rawcount_offset"]
    records = None
    if want_records and len(idx):          # SeqIO.write text of every kept read: one C call, views of one buffer
        records = batch.records_text(idx)
    return PassBatch(cfg, ctx.patterns, batch.first_read + idx, ids, sel, raw_copy, raw_off, tabs, records, leases)


class _DeviceWorker:
    """One device: a context per config, a ring of pinned batch slots.  The first context uploads and
    packs each batch; the others scan the same device-resident batch (tps_submit_shared)."""

    def __init__(self, device, cfgs, max_batch_reads, max_batch_bases, depth, max_pass_reads, rawcount_capacity,
                 context_factory, ends_first=False, ends_raw_bytes=1 << 28, leaders=(0,)):
        self.device = device
        self.cfgs = cfgs
        self.leaders = set(leaders)     # configs whose context can upload and pack a batch (full-size buffers)
        self.max_batch_reads = max_batch_reads
        self.max_batch_bases = max_batch_bases
        self.max_pass_reads = max_pass_reads
        self.ctxs = []
        # ends-first: every context also scans its own region batches (the TRC-pass reads differ per phrase)
        self.region_cap = 0
        if ends_first:
            longest = max(max(1, c.maxlengthtelo) for c in cfgs)
            self.region_cap = int(min(max_batch_bases, max(1 << 20, max_pass_reads * longest)))
        # page-locking the batch slots is slow (about 1.4 GB/s on a cloud VM: 0.5 s for 3 x 256 MiB, more than
        # scanning a 10 GB file takes), so a helper thread does it while the contexts are created and the
        # first batches are parsed; the scan takes the slots as they appear
        self.depth = depth
        self.slot_bases = max_batch_bases
        if ends_first:      # head + tail of a batch's reads never exceed half of its file text
            self.slot_bases = int(min(max_batch_bases, max(1 << 20, ends_raw_bytes // 2 + (1 << 20))))
            self.region_cap = int(min(self.region_cap, 64 << 20))
        self.slots = []
        self._slot_cv = threading.Condition()
        self._slot_error = None
        self._slot_thread = threading.Thread(target=self._alloc_slots, name=f"tps-pin{device}", daemon=True)
        self._slot_thread.start()
        t0 = time.perf_counter()
        try:
            for k, c in enumerate(cfgs):
                # followers never upload whole batches: they need no base / code buffers beyond the region batches
                self.ctxs.append(context_factory(c, device, max_batch_reads,
                                                 max_batch_bases if k in self.leaders else max(1, self.region_cap), depth,
                                                 max_pass_reads, rawcount_capacity if c.want_rawcount else 0))
        except BaseException:
            self._slot_thread.join()
            raise
        self.regs = {}                       # ends-first: region staging per config, page-locked at its first use
        if os.environ.get("TOPSICLE_TIMING"):
            print(f"[timing] device {device}: {len(cfgs)} context(s) {time.perf_counter() - t0:.3f} s",
                  file=sys.stderr)

    def _alloc_slots(self):
        try:
            t0 = time.perf_counter()
            with numa.near_device(self.device):      # pinned staging on the GPU's own NUMA node
                for _ in range(self.depth):
                    s = _Slot(self.slot_bases, self.max_batch_reads)
                    with self._slot_cv:
                        self.slots.append(s)
                        self._slot_cv.notify_all()
            if os.environ.get("TOPSICLE_TIMING"):
                print(f"[timing] device {self.device}: {self.depth} pinned slots of {self.slot_bases >> 20} MiB in "
                      f"{time.perf_counter() - t0:.3f} s (background)", file=sys.stderr)
        except BaseException as e:  # noqa: BLE001 - handed to whoever waits for a slot
            with self._slot_cv:
                self._slot_error = e
                self._slot_cv.notify_all()

    def slot(self, i):
        """Slot i, waiting for the helper thread to page-lock it if need be."""
        with self._slot_cv:
            while len(self.slots) <= i and self._slot_error is None:
                self._slot_cv.wait()
            if len(self.slots) <= i:
                raise self._slot_error
            return self.slots[i]

    def _region_staging(self, ci):
        if ci not in self.regs:
            with numa.near_device(self.device):
                self.regs[ci] = (_Slot(self.region_cap, self.max_pass_reads), engine.PinnedBuffer(self.max_pass_reads))
        return self.regs[ci]

    def close(self):
        self._slot_thread.join()
        for c in self.ctxs:
            c.close()
        for s in self.slots:
            s.free()
        self.slots = []
        for reg, tails in self.regs.values():
            reg.free()
            tails.free()
        self.regs = {}

    def submit(self, bases, starts, lens, n_reads, true_lens=None, group=None):
        """Upload and scan one batch under the configs `group` (default: all): the first uploads and packs,
        the others scan the same device-resident batch."""
        group = group if group is not None else range(len(self.ctxs))
        lead = self.ctxs[group[0]]
        if true_lens is not None:
            bid = lead.submit_ends(bases, starts, lens, true_lens, n_reads)
        else:
            bid = lead.submit_spans(bases, starts, lens, n_reads)
        for k in group[1:]:
            self.ctxs[k].submit_shared(lead, bid)
        return bid

    # -- ends-first mode: steps 2/3 of the TRC-pass reads of an ends batch
    def _regions_submit(self, ci, batch, rows, group):
        """Gather the regions of the longest prefix of reads `group` that fits one region batch into config ci's
        staging (one call of the reader library, threads) and submit it; returns (reads used, batch id, bases,
        reads left over)."""
        cfg, ctx = self.cfgs[ci], self.ctxs[ci]
        reg, reg_tails = self._region_staging(ci)
        rb = reg.bases.array
        rstarts = reg.offsets.array.view(np.uint64)
        rlens = reg.lens.array.view(np.uint32)
        rtails = reg_tails.array
        cut = group[:self.max_pass_reads]
        gt = rows["tail"][cut].astype(np.uint8)
        n = batch.gather_regions(cut, gt, cfg.maxlengthtelo, rb[:self.region_cap], rstarts, rlens)
        if n == 0:
            raise engine.TpsError(-4, f"the region of read {cut[0]} does not fit the batch capacity "
                                      f"of {self.region_cap}")
        rtails[:n] = gt[:n]
        at = int(rstarts[n - 1]) + int(rlens[n - 1])
        bid = ctx.submit_regions(rb[:at], rstarts[:n], rlens[:n], rtails[:n], n)
        return cut[:n], bid, at, group[n:]

    def _scan_regions(self, ci, batch, rows, first, tables, leases):
        """Steps 2/3 of the TRC-pass reads of an ends batch under config ci: `first` is the region batch already in
        flight (_regions_submit); status / n_windows / bkp / telo_length are merged into `rows`, raw-count tables
        go to `tables[i]` (views of leased landing buffers, collected in `leases`).  Reads that did not fit the
        first batch follow one batch at a time; a batch that overflows the raw-count buffer is halved."""
        cfg, ctx = self.cfgs[ci], self.ctxs[ci]
        pending, inflight = [], first
        while pending or inflight is not None:
            if inflight is None:
                inflight = self._regions_submit(ci, batch, rows, pending.pop(0))
            used, bid, at, rest = inflight
            inflight = None
            if rest:
                pending.insert(0, rest)
            n = len(used)
            try:
                rows2, raw2, lease = ctx.wait_leased(bid)
                self._region_bases += at
            except engine.TpsError as e:
                if e.code != -4 or n <= 1:
                    raise
                pending.insert(0, used[n // 2:])
                pending.insert(0, used[:n // 2])
                continue
            for f in ("status", "n_windows", "bkp", "telo_length"):
                rows[f][used] = rows2[f][:n]
            if cfg.want_rawcount and raw2 is not None:
                if lease is not None:
                    leases.append(lease)
                else:
                    raw2 = np.array(raw2, copy=True)
                for j, i in enumerate(used):
                    tables[i] = ctx.rawcount_table(rows2, raw2, j)
            elif lease is not None:
                lease.release()

    def finish_ends(self, item, records_cfg, keep, group):
        slot, batch, bid, seq = item
        res = BatchResult(seq=seq, first_read=batch.first_read, n_reads=batch.n_reads, n_bases=batch.n_bases,
                          n_scanned=0, n_uploaded=int(batch.span))
        self._region_bases = 0
        waited = [self.ctxs[k].wait(bid)[0] for k in group]      # step-1 rows under every config of the job
        firsts = {}
        for gi, ci in enumerate(group):    # the region batches of all configs go out before any is waited for
            rows = waited[gi]
            idx = np.nonzero(rows["status"] == engine.ST_PASS)[0]
            if len(idx) and keep is not None:
                idx = np.array([i for i in idx if batch.read_id(int(i)) in keep], dtype=np.int64)
            if len(idx) and not self.cfgs[ci].step1_only:
                firsts[ci] = self._regions_submit(ci, batch, rows, [int(i) for i in idx])
        for gi, ci in enumerate(group):
            cfg, ctx = self.cfgs[ci], self.ctxs[ci]
            rows = waited[gi]
            tables, leases = {}, []
            if ci in firsts:
                self._scan_regions(ci, batch, rows, firsts[ci], tables, leases)
            res.passes.append(harvest(cfg, ctx, batch, rows, None, records_cfg == ci, keep,
                                      tables=tables if cfg.want_rawcount else None, leases=leases))
            if gi == 0:
                res.n_scanned = int((rows["status"] != engine.ST_FILTERED).sum())
        res.n_uploaded += self._region_bases
        batch.release()
        return res

    def _scan_sub(self, ci, bases, starts, lens, lo, hi, lead=0):
        """Synchronous scan of reads [lo, hi) of a batch under config ci (uploaded by the context of config
        `lead`), splitting again on capacity overflow (more TRC-pass reads or raw counts than the context's
        per-batch capacity)."""
        base0 = int(starts[lo])
        sub_starts = (starts[lo:hi] - starts[lo]).astype(np.uint64)
        sub_lens = np.ascontiguousarray(lens[lo:hi])
        sub_bases = np.ascontiguousarray(bases[base0:int(starts[hi - 1]) + int(lens[hi - 1])])
        try:
            bid = self.ctxs[lead].submit_spans(sub_bases, sub_starts, sub_lens, hi - lo)
            if ci == lead:
                rows, raw, lease = self.ctxs[lead].wait_leased(bid)
            else:
                self.ctxs[ci].submit_shared(self.ctxs[lead], bid)
                try:
                    self.ctxs[lead].wait(bid)
                except engine.TpsError as e:       # the leader's own overflow is irrelevant here
                    if e.code != -4:
                        raise
                rows, raw, lease = self.ctxs[ci].wait_leased(bid)
            return [(lo, hi, rows, raw, lease)]
        except engine.TpsError as e:
            if e.code != -4 or hi - lo <= 1:
                raise
            mid = (lo + hi) // 2
            return (self._scan_sub(ci, bases, starts, lens, lo, mid, lead)
                    + self._scan_sub(ci, bases, starts, lens, mid, hi, lead))

    def finish(self, item, records_cfg, keep, group=None):
        """Results of one submitted batch: BatchResult.passes[j] = TRC-pass reads under config group[j]."""
        group = list(group) if group is not None else list(range(len(self.ctxs)))
        if isinstance(item[1], fastx.EndsBatch):
            return self.finish_ends(item, records_cfg, keep, group)
        slot, batch, bid, seq = item
        res = BatchResult(seq=seq, first_read=batch.first_read, n_reads=batch.n_reads, n_bases=batch.n_bases,
                          n_scanned=0, n_uploaded=int(batch.span))
        n = batch.n_reads
        waited = []
        for k in group:                            # release every context's slot before any re-scan
            try:
                waited.append(self.ctxs[k].wait_leased(bid))  # tables: views of a leased landing buffer
            except engine.TpsError as e:
                if e.code != -4 or n <= 1:
                    raise
                waited.append(None)
        for gi, ci in enumerate(group):
            cfg, ctx = self.cfgs[ci], self.ctxs[ci]
            if waited[gi] is not None:
                parts = [(0, n, waited[gi][0], waited[gi][1], waited[gi][2])]
            else:
                mid = n // 2
                parts = (self._scan_sub(ci, slot.bases.array, batch.offsets, batch.lens, 0, mid, group[0])
                         + self._scan_sub(ci, slot.bases.array, batch.offsets, batch.lens, mid, n, group[0]))
            passes = []
            scanned = 0
            for lo, hi, rows, raw, lease in parts:
                scanned += int((rows["status"] != engine.ST_FILTERED).sum())
                view = _BatchView(batch, lo)
                passes.append(harvest(cfg, ctx, view, rows, raw, records_cfg == ci, keep,
                                      leases=[lease] if lease is not None else ()))
            # one part unless the batch overflowed a capacity and was re-scanned in halves (then: plain list)
            if len(passes) == 1:
                res.passes.append(passes[0])
            else:
                flat = []
                for part in passes:
                    for p in part:
                        if p.counts is not None:
                            p.counts = p.counts.copy()     # the part's landing buffer goes back right away
                        flat.append(p)
                    part.release()
                res.passes.append(flat)
            if gi == 0:
                res.n_scanned = scanned
        batch.release()
        return res


class _BatchView:
    """Reads [lo, ...) of a batch addressed from 0 (for split re-scans)."""

    def __init__(self, batch, lo):
        self.b = batch
        self.lo = lo
        self.first_read = batch.first_read + lo

    def read_id(self, i):
        return self.b.read_id(self.lo + i)

    def read_ids(self, indices):
        return self.b.read_ids(np.asarray(indices, dtype=np.int64) + self.lo)

    def record_text(self, i):
        return self.b.record_text(self.lo + i)

    def records_text(self, indices):
        return self.b.records_text(np.asarray(indices, dtype=np.int64) + self.lo)


class Scanner:
    """Contexts + pinned batch rings on a set of devices, reusable across files (creating them costs
    far more than scanning a small file).  `scan_file` may be called any number of times."""

    def __init__(self, cfgs: Sequence[ScanConfig], *, devices: Sequence[int] = (0,), threads: int = 0,
                 max_batch_bases: int = 1 << 28, max_batch_reads: int = 1 << 17, depth: int = 3,
                 max_pass_reads: int = 0, rawcount_capacity: int = 0, context_factory=None,
                 ends_first: bool = False, ends_raw_bytes: int = 1 << 28, leaders: Sequence[int] = (0,)):
        """`leaders`: the configs that may come first in a FileJob's `cfg_ids` (their contexts get full-size
        upload buffers).  The CLI scans every file under all telophrases of one pattern, so config 0 leads; a mixed
        batch of files with different `--pattern`s (one config per pattern) makes every config a leader."""
        context_factory = context_factory or make_context
        self.ends_first = bool(ends_first)
        self.ends_raw_bytes = int(ends_raw_bytes)        # file text per ends batch
        self.end_len = max(int(c.no_bp) for c in cfgs)   # head / tail bases every config needs
        # whole-read mode, a record longer than a batch: the ends the reader keeps of it hold every base any config
        # looks at (step 1: no_bp; step 2: maxlengthtelo from either end) and stay longer than min_seq_length
        self.clip_bases = max(max(int(c.no_bp), int(c.maxlengthtelo), int(c.min_seq_length) // 2 + 1) for c in cfgs)
        if not max_pass_reads:
            max_pass_reads = max(1024, max_batch_reads // 8)
        if not rawcount_capacity and any(c.want_rawcount for c in cfgs):
            per_read = max(windows_per_read(c) * len(c.patterns) for c in cfgs if c.want_rawcount)
            rawcount_capacity = max(1 << 20, min(1 << 30, per_read * max_pass_reads))
        self.cfgs = list(cfgs)
        self.threads = threads
        self.workers = []
        devices = list(devices)
        made = [None] * len(devices)
        failed = []

        def make(i, d):      # CUDA context creation + device allocations take ~0.4 s per GPU: all devices at once
            try:
                made[i] = _DeviceWorker(d, self.cfgs, max_batch_reads, max_batch_bases, depth, max_pass_reads,
                                        rawcount_capacity, context_factory, ends_first=self.ends_first,
                                        ends_raw_bytes=self.ends_raw_bytes, leaders=leaders)
            except BaseException as e:  # noqa: BLE001 - re-raised below
                failed.append(e)

        if len(devices) > 1:
            ts = [threading.Thread(target=make, args=(i, d), name=f"tps-init{d}") for i, d in enumerate(devices)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        else:
            for i, d in enumerate(devices):
                make(i, d)
        self.workers = [w for w in made if w is not None]
        if failed:
            self.close()
            raise failed[0]

    def close(self):
        for w in self.workers:
            w.close()
        self.workers = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def scan_file(self, path: str, sink: Callable[[BatchResult], None], *, records_cfg: int | None = None,
                  keep_ids=None) -> FileStats:
        """Scan every read of `path` under each config (one parse, one pass).

        `sink(BatchResult)` is called in file order; BatchResult.passes[k] holds the TRC-pass reads
        under cfgs[k].  `records_cfg=k` attaches the SeqIO.write text of the reads passing cfgs[k].
        `keep_ids` restricts the harvest to those read ids.  A parse error is raised (scan_files, which walks
        many files, hands it to the job's on_error instead and goes on)."""
        job = FileJob(path, sink, records_cfg=records_cfg, keep_ids=keep_ids)
        stats = self.scan_files([job], readers=1)[0]
        if job.error is not None:
            raise job.error
        return stats

    def scan_files(self, jobs: Sequence["FileJob"], *, readers: int = 0) -> list:
        """Scan several files concurrently (the reference's unit of parallelism is the input file,
        main.py:232-235): `readers` reader threads each parse one file at a time into the shared pool of
        pinned slots, one worker per device scans whatever is ready.  Every job's sink still sees its
        batches in file order; `job.on_done(stats)` fires after its last batch.  Returns the FileStats
        in job order.

        Threads: the C parser is itself multi-threaded on plain files; `.gz` files are inflated by one
        thread each, which is why several readers matter for directories of compressed files."""
        jobs = list(jobs)
        for j in jobs:
            if j.cfg_ids is not None and (not j.cfg_ids or j.cfg_ids[0] not in self.workers[0].leaders):
                raise ValueError(f"{j.path}: cfg_ids {j.cfg_ids} must start with one of the Scanner's leaders "
                                 f"{sorted(self.workers[0].leaders)}")
        n_workers = len(self.workers)
        n_threads = self.threads or len(os.sched_getaffinity(0))
        if readers <= 0:
            readers = min(len(jobs), max(1, n_threads // 2), 8)
        readers = max(1, min(readers, len(jobs)))
        fx_threads = max(1, n_threads // readers)
        errors = []
        ready = queue.Queue()
        free = queue.Queue()
        def feed_slots():        # slots reach the readers as the helper threads finish page-locking them
            try:
                for i in range(max(w.depth for w in self.workers)):
                    for w in self.workers:
                        if i < w.depth:
                            free.put(w.slot(i))
            except BaseException as e:  # noqa: BLE001
                errors.append(e)
                free.put(None)

        feeder = threading.Thread(target=feed_slots, name="tps-slots")
        w0 = self.workers[0]
        todo = queue.Queue()
        for j in jobs:
            j._reset()
            todo.put(j)
        lock = threading.Lock()
        tm_parse = [0.0]

        def finalize(job):
            t = time.perf_counter()
            if job.fx is not None:
                job.fx.close()
                job.fx = None
            job.stats.timing["close"] = time.perf_counter() - t
            if job.on_done is not None:
                job.on_done(job.stats)

        def delivered(job, res):
            """Called by a worker after it finished one batch of `job`."""
            job.ordered.put(res)
            with lock:
                job.stats.n_uploaded += res.n_uploaded
                job.n_delivered += 1
                done = job.eof and job.n_delivered == job.stats.n_batches
            if done:
                finalize(job)

        def parse_failed(job, e):
            """An unreadable file or record ends THAT file, not the run: the reference walks every file of the
            directory, logs the parse error and keeps the reads parsed before it (allsteps.py:137-149,
            main.py:224-235)."""
            job.error = e
            if job.on_error is not None:
                job.on_error(e)

        def read_loop():
            try:
                while not errors:
                    try:
                        job = todo.get_nowait()
                    except queue.Empty:
                        return
                    t = time.perf_counter()
                    try:
                        job.fx = fastx.FastxFile(job.path, threads=fx_threads)
                    except (fastx.FastxError, OSError) as e:
                        parse_failed(job, e)
                    if job.fx is not None and not self.ends_first:
                        # a record longer than a batch (a chromosome): the scan reads a bounded stretch of either end
                        job.fx.set_clip(self.clip_bases)
                    job.stats.timing["open"] = time.perf_counter() - t
                    seq = 0
                    if job.fx is not None:
                        job.stats.format_name = job.fx.format_name
                    while job.fx is not None and not errors:
                        slot = free.get()
                        if slot is None:
                            free.put(None)      # let the other readers see the stop signal too
                            return
                        t = time.perf_counter()
                        try:
                            if self.ends_first:
                                batch = job.fx.next_ends(slot.bases.array, slot.offsets.array.view(np.uint64),
                                                         slot.lens.array.view(np.uint32),
                                                         slot.true_lens.array.view(np.uint32), self.end_len,
                                                         raw_cap=self.ends_raw_bytes, max_reads=w0.max_batch_reads,
                                                         max_bases=w0.max_batch_bases, recs=slot.recs)
                            else:
                                batch = job.fx.next_spans(slot.bases.array, slot.offsets.array.view(np.uint64),
                                                          slot.lens.array.view(np.uint32),
                                                          max_reads=w0.max_batch_reads, max_span=w0.max_batch_bases,
                                                          recs=slot.recs)
                        except fastx.FastxError as e:
                            free.put(slot)
                            parse_failed(job, e)
                            break
                        job.stats.timing["parse"] += time.perf_counter() - t
                        if batch is None:
                            free.put(slot)
                            break
                        job.stats.n_reads += batch.n_reads
                        job.stats.n_bases += batch.n_bases
                        job.stats.n_batches += 1
                        ready.put((job, slot, batch, seq))
                        seq += 1
                    with lock:
                        job.eof = True
                        done = job.n_delivered == job.stats.n_batches
                    if done:
                        finalize(job)
            except BaseException as e:  # noqa: BLE001 - reported to the caller below
                errors.append(e)
                free.put(None)   # wake the readers that wait for a slot: the workers stop returning them

        def work_loop(w: _DeviceWorker):
            inflight = []
            depth = w.depth

            def finish_oldest():
                job, item = inflight.pop(0)
                t = time.perf_counter()
                res = w.finish(item, job.records_cfg, job.keep_ids, job.cfg_ids)
                job.stats.timing["finish"] += time.perf_counter() - t
                free.put(item[0])
                delivered(job, res)

            try:
                while not errors:
                    if inflight and (len(inflight) >= depth or ready.empty()):
                        finish_oldest()
                        continue
                    got = ready.get()
                    if got is None:
                        break
                    job, slot, batch, seq = got
                    t = time.perf_counter()
                    bid = w.submit(slot.bases.array[:batch.span], batch.offsets[:batch.n_reads],
                                   batch.lens[:batch.n_reads], batch.n_reads,
                                   batch.true_lens[:batch.n_reads] if isinstance(batch, fastx.EndsBatch) else None,
                                   job.cfg_ids)
                    job.stats.timing["submit"] += time.perf_counter() - t
                    inflight.append((job, (slot, batch, bid, seq)))
                while inflight and not errors:
                    finish_oldest()
            except BaseException as e:  # noqa: BLE001
                errors.append(e)
                free.put(None)   # unblock the readers

        # the reader threads re-acquire the GIL between two C parser calls while the workers harvest in
        # Python: with the default 5 ms switch interval that wait would dominate the parse of a batch
        old_switch = sys.getswitchinterval()
        sys.setswitchinterval(1e-4)
        rthreads = [threading.Thread(target=read_loop, name=f"tps-reader{i}") for i in range(readers)]
        wthreads = [threading.Thread(target=work_loop, args=(w,), name=f"tps-dev{w.device}") for w in self.workers]
        try:
            for t in [feeder] + rthreads + wthreads:
                t.start()
            for t in rthreads:
                t.join()
            for _ in range(n_workers):
                ready.put(None)
            for t in wthreads:
                t.join()
            feeder.join()
            if errors:
                raise errors[0]
        finally:
            sys.setswitchinterval(old_switch)
            for j in jobs:
                if j.fx is not None:
                    j.fx.close()
                    j.fx = None
        return [j.stats for j in jobs]


class FileJob:
    """One input file of a `Scanner.scan_files` call."""

    def __init__(self, path: str, sink: Callable[[BatchResult], None], *, records_cfg: int | None = None,
                 keep_ids=None, on_done: Callable[[FileStats], None] | None = None,
                 on_error: Callable[[Exception], None] | None = None, cfg_ids: Sequence[int] | None = None):
        self.path = path
        # the Scanner's configs this file is scanned under (default: all); BatchResult.passes follows this order
        self.cfg_ids = list(cfg_ids) if cfg_ids is not None else None
        self.sink = sink
        self.records_cfg = records_cfg
        self.keep_ids = keep_ids
        self.on_done = on_done
        self.on_error = on_error     # parse error of this file (the scan goes on with the other files)
        self._reset()

    def _reset(self):
        self.stats = FileStats()
        for k in ("open", "parse", "submit", "finish", "close"):
            self.stats.timing[k] = 0.0
        self.ordered = _OrderedSink(self.sink)
        self.fx = None
        self.eof = False
        self.n_delivered = 0
        self.error = None            # the FastxError / OSError that ended this file early, if any


def scan_file(path: str, cfgs: Sequence[ScanConfig], sink: Callable[[BatchResult], None], *,
              records_cfg: int | None = None, keep_ids=None, **scanner_kw) -> FileStats:
    """One-shot convenience: Scanner(cfgs, **scanner_kw).scan_file(path, sink, ...)."""
    with Scanner(cfgs, **scanner_kw) as sc:
        return sc.scan_file(path, sink, records_cfg=records_cfg, keep_ids=keep_ids)


def collect_file(path: str, cfgs: Sequence[ScanConfig], **kw):
    """scan_file with an in-memory sink -> (stats, per-config list[PassRead], scanned counts)."""
    per_cfg = [[] for _ in cfgs]
    scanned = [0]

    def sink(res: BatchResult):
        for k, ps in enumerate(res.passes):
            per_cfg[k].extend(ps)
        scanned[0] += res.n_scanned

    stats = scan_file(path, cfgs, sink, **kw)
    stats.n_scanned = scanned[0]
    return stats, per_cfg


def scan_named_read(path: str, cfgs: Sequence[ScanConfig], read_id: str, *, device: int = 0, threads: int = 0,
                    context_factory=None):
    """Per-read entry (bound_detect / rawCountPattern, allsteps.py:252-259, 375-382): find the record(s)
    whose id is `read_id`, scan only those under every config.  Returns per-config list[PassRead]."""
    context_factory = context_factory or make_context
    picked, lengths, indices = [], [], []
    fx = fastx.FastxFile(path, threads=threads)
    try:
        bases = np.empty(1 << 27, dtype=np.uint8)
        offsets = np.empty((1 << 17) + 1, dtype=np.uint64)
        while True:
            try:
                b = fx.next_batch(bases, offsets)
            except fastx.FastxError as e:
                if e.code != -4:
                    raise
                bases = np.empty(bases.size * 4, dtype=np.uint8)
                continue
            if b is None:
                break
            for i in b.find_id(read_id):
                picked.append(b.sequence(i))
                lengths.append(len(picked[-1]))
                indices.append(b.first_read + i)
            b.release()
    finally:
        fx.close()
    per_cfg = [[] for _ in cfgs]
    if not picked:
        return per_cfg
    sel_bases, sel_off = engine.pack_reads(picked)
    for k, cfg in enumerate(cfgs):
        ctx = context_factory(cfg, device, len(picked), max(1, int(sel_off[-1])), 1, len(picked),
                              max(1, windows_per_read(cfg) * len(cfg.patterns) * len(picked))
                              if cfg.want_rawcount else 0)
        try:
            rows, raw = ctx.scan(sel_bases, sel_off)
            for i, r in enumerate(rows):
                if r["status"] < engine.ST_PASS:
                    continue
                cnt = int(r["match_count"])
                pr = PassRead(index=indices[i], read_id=read_id, literal=ctx.patterns[int(r["best_pattern"])],
                              tail=engine.TAIL_NAMES[int(r["tail"])], count=cnt,
                              trc=engine.trc_value(cnt, cfg.len_telopattern, cfg.no_bp), status=int(r["status"]),
                              n_windows=int(r["n_windows"]), telo_length=int(r["telo_length"]),
                              length=int(r["length"]))
                if cfg.want_rawcount and raw is not None:
                    tab = ctx.rawcount_table(rows, raw, i)
                    pr.counts = None if tab is None else tab.copy()
                per_cfg[k].append(pr)
        finally:
            ctx.close()
    return per_cfg
