"""TEST INFRASTRUCTURE ONLY -- an oracle-backed stand-in for `engine.ScanContext`.

Lets the CPU test tier (`-m "not gpu"`) drive the host-side logic (reader -> pipeline ->
CLI writers, sharding, gathering) without a GPU: batches are answered by the CPU oracle in
the `tps_row` layout the CUDA library produces.  It is injected by monkeypatching
`pipeline.make_context` / `engine.PinnedBuffer`; the product never imports it.
"""
import numpy as np

from oracle import topsicle_oracle as orc
from topsicle_b200 import engine


class NumpyPinned:
    """PinnedBuffer stand-in (plain host memory)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.array = np.zeros(max(self.nbytes, 1), dtype=np.uint8)
        self.ptr = self.array.ctypes.data

    def free(self):
        self.array = None


class OracleContext:
    def __init__(self, cfg, device, max_batch_reads, max_batch_bases, n_slots, max_pass_reads=0,
                 rawcount_capacity=0):
        self.cfg = cfg
        self.patterns = [p.upper() for p in cfg.patterns]
        self.max_pass = max_pass_reads or max_batch_reads
        self.max_batch_reads, self.max_batch_bases = max_batch_reads, max_batch_bases
        self.rawcount_capacity = rawcount_capacity
        self.want_rawcount = cfg.want_rawcount
        self.thr = (cfg.count_threshold_override if cfg.count_threshold_override is not None
                    else engine.count_threshold(cfg.cutoff, cfg.len_telopattern, cfg.no_bp))
        self._pending = {}
        self.closed = False

    def submit(self, bases, offsets):
        bid = next(engine._batch_ids)
        assert len(offsets) - 1 <= self.max_batch_reads and int(offsets[-1]) <= self.max_batch_bases
        self._pending[bid] = (bases.tobytes(), np.array(offsets, dtype=np.uint64))
        return bid

    def submit_spans(self, bases, starts, lens, n_reads):
        """Span batch -> the same back-to-back form the oracle loop below reads."""
        bid = next(engine._batch_ids)
        assert n_reads <= self.max_batch_reads and bases.size <= self.max_batch_bases
        parts = [bases[int(starts[i]):int(starts[i]) + int(lens[i])].tobytes() for i in range(n_reads)]
        assert all(int(starts[i]) + int(lens[i]) <= int(starts[i + 1]) for i in range(n_reads - 1)), "reads overlap"
        off = np.zeros(n_reads + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(p) for p in parts], dtype=np.uint64)
        self._pending[bid] = (b"".join(parts), off)
        return bid

    def _submit_parts(self, bases, starts, lens, n_reads, kind, extra):
        bid = next(engine._batch_ids)
        assert n_reads <= self.max_batch_reads and bases.size <= self.max_batch_bases
        parts = [bases[int(starts[i]):int(starts[i]) + int(lens[i])].tobytes() for i in range(n_reads)]
        off = np.zeros(n_reads + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(p) for p in parts], dtype=np.uint64)
        self._pending[bid] = (b"".join(parts), off, kind, np.array(extra[:n_reads]))
        return bid

    def submit_ends(self, bases, starts, lens, true_lens, n_reads):
        """tps_submit_ends: head + tail of every read, real lengths aside; step 1 only."""
        return self._submit_parts(bases, starts, lens, n_reads, "ends", true_lens)

    def submit_regions(self, bases, starts, lens, tails, n_reads):
        """tps_submit_regions: regions of reads that passed step 1, tail forced per read."""
        assert not self.cfg.step1_only
        return self._submit_parts(bases, starts, lens, n_reads, "regions", tails)

    def submit_shared(self, owner, bid):
        self._pending[bid] = owner._pending[bid]
        return bid

    def wait_leased(self, bid):
        rows, raw = self.wait(bid)
        return rows, raw, None

    def wait(self, bid, raw_view=False):
        buf, off, *mode = self._pending.pop(bid)
        kind, extra = mode if mode else (None, None)
        assert kind != "regions" or len(mode) == 2
        cfg = self.cfg
        n = len(off) - 1
        rows = np.zeros(n, dtype=engine.ROW_DTYPE)
        rows["bkp"] = -1
        rows["telo_length"] = -1
        rows["rawcount_offset"] = engine.NO_RAWCOUNT
        raw_parts, raw_at, n_pass = [], 0, 0
        for i in range(n):
            seq = buf[int(off[i]):int(off[i + 1])].decode("latin-1")
            true_len = int(extra[i]) if kind == "ends" else len(seq)
            rows["length"][i] = true_len
            if kind != "regions" and not true_len > cfg.min_seq_length:
                continue
            tail, bi, cnt, ms, me, hc, tc = orc.trc_read(seq, self.patterns, cfg.len_telopattern, cfg.no_bp)
            forced = cfg.force_tail
            if kind == "regions":
                forced = "forward" if int(extra[i]) == 0 else "reverse"
            if forced:
                tail = forced
                cs = hc if tail == "forward" else tc
                cnt = max(cs)
                bi = cs.index(cnt)
            rows["tail"][i] = 0 if tail == "forward" else 1
            rows["best_pattern"][i] = bi
            rows["match_count"][i], rows["head_max"][i], rows["tail_max"][i] = cnt, ms, me
            if cnt < self.thr and kind != "regions":
                rows["status"][i] = engine.ST_BELOW
                continue
            rows["status"][i] = engine.ST_PASS
            if cfg.step1_only or kind == "ends":
                continue
            n_pass += 1
            if n_pass > self.max_pass:
                raise engine.TpsError(-4, "max_pass_reads exceeded")
            region = orc.oriented_region(seq, tail, cfg.trimfirst, cfg.maxlengthtelo)
            counts = orc.window_counts(region, self.patterns, cfg.window_size, cfg.slide)
            rows["n_windows"][i] = counts.shape[0]
            if cfg.want_rawcount and counts.shape[0]:
                rows["rawcount_offset"][i] = raw_at
                raw_parts.append(counts.astype(np.uint8).reshape(-1))
                raw_at += counts.size
                if raw_at > self.rawcount_capacity:
                    raise engine.TpsError(-4, "rawcount_capacity exceeded")
            if counts.shape[0] < 7:
                rows["status"][i] = engine.ST_BADSEG
                continue
            b = orc.change_point_exact(counts.sum(axis=1))
            rows["bkp"][i] = b
            rows["telo_length"][i] = cfg.trimfirst + cfg.slide * b
        raw = np.concatenate(raw_parts) if raw_parts else (np.zeros(0, np.uint8) if cfg.want_rawcount else None)
        return rows, raw

    def scan(self, bases, offsets):
        return self.wait(self.submit(bases, offsets))

    def rawcount_table(self, rows, raw, i):
        off = int(rows["rawcount_offset"][i])
        if raw is None or off == engine.NO_RAWCOUNT:
            return None
        nw, npat = int(rows["n_windows"][i]), len(self.patterns)
        return raw[off:off + nw * npat].reshape(nw, npat)

    def close(self):
        self.closed = True


def install(monkeypatch):
    """Route pipeline contexts / pinned buffers to the CPU stand-ins."""
    from topsicle_b200 import pipeline
    monkeypatch.setattr(pipeline, "make_context", OracleContext)
    monkeypatch.setattr(engine, "PinnedBuffer", NumpyPinned)
