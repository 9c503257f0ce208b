"""ctypes binding of the C-ABI in include/topsicle_b200.h plus a small batch engine.

The CUDA library is the only implementation of the scan: if `libtopsicle_b200.so`
is missing or no CUDA device is visible this module raises -- there is no CPU
fallback (nothing under `oracle/` is ever imported from here).
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from typing import Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOPSICLE_B200_LIB") or os.path.join(_HERE, "libtopsicle_b200.so")  # override: tuning builds

TPS_MAX_PATTERNS = 64
TPS_MAX_PATTERN_LEN = 32

ST_FILTERED, ST_BELOW, ST_PASS, ST_BADSEG = 0, 1, 2, 3
FLAG_STEP1_ONLY, FLAG_FORCE_FORWARD, FLAG_FORCE_REVERSE = 1, 2, 4
TAIL_NAMES = ("forward", "reverse")
NO_RAWCOUNT = 0xFFFFFFFFFFFFFFFF


class TpsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"topsicle_b200 error {code}: {msg}")
        self.code = code


class TpsParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("n_patterns", C.c_uint32),
        ("pattern_len", C.c_uint8 * TPS_MAX_PATTERNS),
        ("patterns", (C.c_char * TPS_MAX_PATTERN_LEN) * TPS_MAX_PATTERNS),
        ("min_seq_length", C.c_uint32),
        ("no_bp", C.c_uint32),
        ("count_threshold", C.c_uint32),
        ("window_size", C.c_uint32),
        ("slide", C.c_uint32),
        ("trimfirst", C.c_uint32),
        ("maxlengthtelo", C.c_uint32),
        ("want_rawcount", C.c_uint32),
        ("n_slots", C.c_uint32),
        ("max_batch_reads", C.c_uint32),
        ("max_pass_reads", C.c_uint32),
        ("flags", C.c_uint32),
        ("max_batch_bases", C.c_uint64),
        ("rawcount_capacity", C.c_uint64),
    ]


ROW_DTYPE = np.dtype([
    ("length", "<u4"), ("status", "u1"), ("tail", "u1"), ("best_pattern", "u1"), ("reserved0", "u1"),
    ("match_count", "<u2"), ("head_max", "<u2"), ("tail_max", "<u2"), ("reserved1", "<u2"),
    ("n_windows", "<u4"), ("bkp", "<i4"), ("telo_length", "<i4"), ("reserved2", "<u4"),
    ("rawcount_offset", "<u8"),
])
assert ROW_DTYPE.itemsize == 40

_lib = None
_batch_ids = itertools.count(1)   # batch ids are unique across contexts (shared batches keep their owner's id)


def load_library() -> C.CDLL:
    """Load libtopsicle_b200.so (built in-tree by `__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TpsError(-100, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                             "g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, u64p, u32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    lib.tps_abi_version.restype = C.c_int
    lib.tps_build_info.restype = C.c_char_p
    lib.tps_device_count.restype = C.c_int
    lib.tps_last_error.restype = C.c_char_p
    lib.tps_last_error.argtypes = [vp]
    lib.tps_create.restype = C.c_int
    lib.tps_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(TpsParams)]
    lib.tps_destroy.restype = None
    lib.tps_destroy.argtypes = [vp]
    lib.tps_alloc_pinned.restype = vp
    lib.tps_alloc_pinned.argtypes = [C.c_size_t]
    lib.tps_free_pinned.restype = None
    lib.tps_free_pinned.argtypes = [vp]
    lib.tps_submit.restype = C.c_int
    lib.tps_submit.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64]
    lib.tps_submit_spans.restype = C.c_int
    lib.tps_submit_spans.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint32, C.c_uint64]
    lib.tps_submit_ends.restype = C.c_int
    lib.tps_submit_ends.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint32, C.c_uint64]
    lib.tps_submit_regions.restype = C.c_int
    lib.tps_submit_regions.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint32, C.c_uint64]
    lib.tps_follow_scan.restype = C.c_int
    lib.tps_follow_scan.argtypes = [C.c_int, vp, vp, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_uint64]
    lib.tps_submit_shared.restype = C.c_int
    lib.tps_submit_shared.argtypes = [vp, vp, C.c_uint64]
    lib.tps_wait.restype = C.c_int
    lib.tps_wait.argtypes = [vp, C.c_uint64, vp, u32p, vp, C.c_uint64, u64p]
    lib.tps_batch_info.restype = C.c_int
    lib.tps_batch_info.argtypes = [vp, C.c_uint64, u32p, u64p]
    lib.tps_scan_device.restype = C.c_int
    lib.tps_scan_device.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, vp]
    lib.tps_scan_device_slot.restype = C.c_int
    lib.tps_scan_device_slot.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint32, C.c_uint64, vp]
    lib.tps_sync.restype = C.c_int
    lib.tps_sync.argtypes = [vp]
    lib.tps_get_timings.restype = C.c_int
    lib.tps_get_timings.argtypes = [vp, C.c_uint32, C.POINTER(C.c_float * 4)]
    lib.tps_get_timeline.restype = C.c_int
    lib.tps_get_timeline.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_float * 4)]
    lib.tps_elapsed_between.restype = C.c_int
    lib.tps_elapsed_between.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    lib.tps_kernel_launches.restype = C.c_uint64
    lib.tps_kernel_launches.argtypes = [vp]
    lib.tps_debug_copy.restype = C.c_int
    lib.tps_debug_copy.argtypes = [vp, C.c_int, vp, C.c_size_t]
    if lib.tps_abi_version() != 1:
        raise TpsError(-101, "ABI version mismatch")
    _lib = lib
    return lib


def device_count() -> int:
    """CUDA devices visible to the library."""
    return int(load_library().tps_device_count())


def count_threshold(cutoff: float, len_telopattern: int, no_bp: int = 1000) -> int:
    """Smallest integer count c with `c / (no_bp / len(telopattern)) > cutoff` evaluated in
    float64 exactly as the reference does (allsteps.py:178,185-186,194,197).  The quotient
    is monotone in c, so the reference's float test equals `count >= threshold`."""
    ratio = no_bp / len_telopattern
    for c in range(0, no_bp + 2):
        if c / ratio > cutoff:
            return c
    return 0xFFFFFFFF


def trc_value(count: int, len_telopattern: int, no_bp: int = 1000) -> float:
    """The reference's TRC float: matches / (no_bp / len(telopattern)) (allsteps.py:178,185)."""
    return int(count) / (no_bp / len_telopattern)


def pack_reads(seqs: Iterable) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate reads (str / bytes) back to back -> (uint8 bases, uint64 offsets[n+1])."""
    bufs = [s.encode("ascii", "replace") if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bufs) + 1, dtype=np.uint64)
    if bufs:
        offsets[1:] = np.cumsum([len(b) for b in bufs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bufs), dtype=np.uint8)
    return bases, offsets


def follow_scan(seqs: Iterable, kmers: Sequence[str], match_len: int, min_seq_length: int, skip: int = 100,
                upto: int = 2000, device: int = 0) -> np.ndarray:
    """Match starts of `kmer(.{match_len - k})`, leftmost non-overlapping, in `seq[skip:upto]` (strand 0) and in the
    complement of `reversed(seq)[skip:upto]` (strand 1) of every read longer than min_seq_length
    (descriptive_plot.py:233-313).  Returns bool[n_reads][2][n_kmers][upto - skip]."""
    lib = load_library()
    bases, offsets = pack_reads(seqs)
    n = len(offsets) - 1
    k = len(kmers[0])
    if any(len(m) != k for m in kmers):
        raise ValueError("the k-mers of a follower scan share one length")
    wpr = (upto - skip + 31) // 32
    sel = np.zeros((n, 2, len(kmers), wpr), dtype=np.uint32)
    if n:
        bases = np.ascontiguousarray(bases)
        rc = lib.tps_follow_scan(device, bases.ctypes.data, offsets.ctypes.data, n, "".join(kmers).upper().encode(),
                                 len(kmers), k, match_len, min_seq_length, skip, upto, sel.ctypes.data, sel.size)
        if rc != 0:
            raise TpsError(rc, lib.tps_last_error(None).decode())
    bits = np.unpackbits(sel.view(np.uint8), axis=-1, bitorder="little")
    return bits[..., :upto - skip].astype(bool)


def group_sums(c_w) -> np.ndarray:
    """Sums of c_w over the groups of five consecutive windows [5j, 5j+5) (the last group may be shorter)."""
    c = np.asarray(c_w, dtype=np.int64)
    pad = (-len(c)) % 5
    return np.concatenate([c, np.zeros(pad, np.int64)]).reshape(-1, 5).sum(axis=1)


class PinnedBuffer:
    """Page-locked host memory from the library, exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        self._lib = load_library()
        self.nbytes = int(nbytes)
        self.ptr = self._lib.tps_alloc_pinned(max(self.nbytes, 1))
        if not self.ptr:
            raise TpsError(-3, f"cannot pin {nbytes} bytes of host memory")
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self._lib.tps_free_pinned(self.ptr)
            self.ptr = None
            self.array = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class RawLease:
    """A page-locked landing buffer that holds the raw-count tables of one finished batch, on loan from its
    context's pool: the tables are handed out as views of it (no second copy of tens of megabytes per batch) and
    stay valid until `release()`."""

    __slots__ = ("_ctx", "_buf")

    def __init__(self, ctx, buf):
        self._ctx, self._buf = ctx, buf

    def release(self):
        if self._buf is not None:
            self._ctx._raw_pool.append(self._buf)
            self._buf = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class ScanContext:
    """One GPU scan context == the arguments of one process_file call (main.py:52-150)."""

    def __init__(self, patterns: Sequence[str], *, len_telopattern: int | None = None, cutoff: float = 0.7,
                 min_seq_length: int = 9000, no_bp: int = 1000, window_size: int = 100, slide: int = 6,
                 trimfirst: int = 100, maxlengthtelo: int = 20000, want_rawcount: bool = False,
                 device: int = 0, n_slots: int = 2, max_batch_reads: int = 1 << 16,
                 max_batch_bases: int = 1 << 28, max_pass_reads: int = 0, rawcount_capacity: int = 0,
                 count_threshold_override: int | None = None, step1_only: bool = False,
                 force_tail: str | None = None):
        self.lib = load_library()
        self.patterns = [p.upper() for p in patterns]
        self.len_telopattern = len_telopattern if len_telopattern is not None else len(self.patterns[0])
        self.no_bp = no_bp
        p = TpsParams()
        p.struct_size = C.sizeof(TpsParams)
        if len(self.patterns) > TPS_MAX_PATTERNS:
            raise TpsError(-1, f"at most {TPS_MAX_PATTERNS} literals are supported")
        p.n_patterns = len(self.patterns)
        for i, lit in enumerate(self.patterns):
            b = lit.encode("ascii")
            if not 1 <= len(b) <= TPS_MAX_PATTERN_LEN:
                raise TpsError(-1, f"literal {lit!r}: length must be in 1..{TPS_MAX_PATTERN_LEN}")
            p.pattern_len[i] = len(b)
            C.memmove(C.addressof(p.patterns[i]), b, len(b))
        p.min_seq_length = min_seq_length
        p.no_bp = no_bp
        p.count_threshold = (count_threshold_override if count_threshold_override is not None
                             else count_threshold(cutoff, self.len_telopattern, no_bp))
        p.window_size, p.slide, p.trimfirst, p.maxlengthtelo = window_size, slide, trimfirst, maxlengthtelo
        p.want_rawcount = 1 if want_rawcount else 0
        p.flags = (FLAG_STEP1_ONLY if step1_only else 0) | {None: 0, "forward": FLAG_FORCE_FORWARD,
                                                             "reverse": FLAG_FORCE_REVERSE}[force_tail]
        p.n_slots = n_slots
        p.max_batch_reads = max_batch_reads
        p.max_batch_bases = max_batch_bases
        p.max_pass_reads = max_pass_reads
        if want_rawcount and not rawcount_capacity:
            nreg = max(0, maxlengthtelo - trimfirst)
            nw = (nreg - window_size) // slide + 1 if nreg >= window_size else 0
            rawcount_capacity = max(1, min(max_batch_reads, 4096) * nw * len(self.patterns))
        p.rawcount_capacity = rawcount_capacity
        self.params = p
        self.max_batch_reads = max_batch_reads
        self.max_batch_bases = max_batch_bases
        self.want_rawcount = bool(want_rawcount)
        self._h = C.c_void_p()
        rc = self.lib.tps_create(C.byref(self._h), device, C.byref(p))
        if rc != 0:
            raise TpsError(rc, self.lib.tps_last_error(None).decode())
        self._inflight = {}
        self._raw_pin = None
        self._raw_pool = []          # free landing buffers of wait_leased()

    # -- lifetime
    def close(self):
        if getattr(self, "_raw_pin", None) is not None:
            self._raw_pin.free()
            self._raw_pin = None
        for b in getattr(self, "_raw_pool", []):
            b.free()
        self._raw_pool = []
        if getattr(self, "_h", None) and self._h.value:
            self.lib.tps_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise TpsError(rc, self.lib.tps_last_error(self._h).decode())

    # -- host-buffer path (reference-facing: H2D and D2H inside)
    def submit(self, bases: np.ndarray, offsets: np.ndarray) -> int:
        assert bases.dtype == np.uint8 and offsets.dtype == np.uint64
        assert bases.flags.c_contiguous and offsets.flags.c_contiguous
        n_reads = len(offsets) - 1
        bid = next(_batch_ids)
        self._check(self.lib.tps_submit(self._h, bases.ctypes.data, offsets.ctypes.data, n_reads, bid))
        self._inflight[bid] = (bases, offsets, n_reads)  # keep buffers alive
        return bid

    def submit_spans(self, bases: np.ndarray, starts: np.ndarray, lens: np.ndarray, n_reads: int) -> int:
        """Submit a span batch: read i = bases[starts[i] : starts[i] + lens[i]] (gaps in between are ignored)."""
        assert bases.dtype == np.uint8 and starts.dtype == np.uint64 and lens.dtype == np.uint32
        assert bases.flags.c_contiguous and starts.flags.c_contiguous and lens.flags.c_contiguous
        bid = next(_batch_ids)
        self._check(self.lib.tps_submit_spans(self._h, bases.ctypes.data, bases.size, starts.ctypes.data,
                                              lens.ctypes.data, n_reads, bid))
        self._inflight[bid] = (bases, (starts, lens), n_reads)  # keep buffers alive
        return bid

    def submit_ends(self, bases: np.ndarray, starts: np.ndarray, lens: np.ndarray, true_lens: np.ndarray,
                    n_reads: int) -> int:
        """Submit an ends batch (step 1 only): read i is uploaded as head + tail, bases[starts[i] : starts[i] +
        lens[i]] with lens[i] = min(L, 2 * no_bp); true_lens[i] = L."""
        assert bases.dtype == np.uint8 and starts.dtype == np.uint64 and lens.dtype == np.uint32
        assert true_lens.dtype == np.uint32
        assert bases.flags.c_contiguous and starts.flags.c_contiguous and lens.flags.c_contiguous
        assert true_lens.flags.c_contiguous
        bid = next(_batch_ids)
        self._check(self.lib.tps_submit_ends(self._h, bases.ctypes.data, bases.size, starts.ctypes.data,
                                             lens.ctypes.data, true_lens.ctypes.data, n_reads, bid))
        self._inflight[bid] = (bases, (starts, lens, true_lens), n_reads)
        return bid

    def submit_regions(self, bases: np.ndarray, starts: np.ndarray, lens: np.ndarray, tails: np.ndarray,
                       n_reads: int) -> int:
        """Submit a region batch (steps 2/3 of reads that passed step 1 elsewhere): read i is the first
        (tails[i] = 0) or last (tails[i] = 1) min(L, maxlengthtelo) bases of the read."""
        assert bases.dtype == np.uint8 and starts.dtype == np.uint64 and lens.dtype == np.uint32
        assert tails.dtype == np.uint8
        assert bases.flags.c_contiguous and starts.flags.c_contiguous and lens.flags.c_contiguous
        assert tails.flags.c_contiguous
        bid = next(_batch_ids)
        self._check(self.lib.tps_submit_regions(self._h, bases.ctypes.data, bases.size, starts.ctypes.data,
                                                lens.ctypes.data, tails.ctypes.data, n_reads, bid))
        self._inflight[bid] = (bases, (starts, lens, tails), n_reads)
        return bid

    def submit_shared(self, owner: "ScanContext", bid: int) -> int:
        """Scan the batch `owner` has in flight as `bid` under this context's parameters, reusing the
        owner's upload and packed reads (no second H2D / K1)."""
        self._check(self.lib.tps_submit_shared(self._h, owner._h, bid))
        self._inflight[bid] = owner._inflight[bid]
        return bid

    def wait(self, bid: int, raw_view: bool = False):
        """Rows (and raw count tables) of a finished batch.  With `raw_view` the tables are returned as a view
        of this context's page-locked landing buffer, valid until its next wait()."""
        bases, offsets, n_reads = self._inflight[bid]
        rows = np.empty(n_reads, dtype=ROW_DTYPE)
        n_pass = C.c_uint32(0)
        elems = C.c_uint64(0)
        raw = None
        try:
            if self.want_rawcount:
                self._check(self.lib.tps_batch_info(self._h, bid, C.byref(n_pass), C.byref(elems)))
                need = min(int(elems.value), int(self.params.rawcount_capacity))
                if self._raw_pin is None or self._raw_pin.nbytes < need:      # page-locked landing zone for the
                    if self._raw_pin is not None:                             # count tables (D2H at PCIe speed)
                        self._raw_pin.free()
                    self._raw_pin = PinnedBuffer(max(need * 5 // 4, 1 << 20))
                raw = self._raw_pin.array[:need]
                self._check(self.lib.tps_wait(self._h, bid, rows.ctypes.data, C.byref(n_pass), raw.ctypes.data,
                                              raw.size, C.byref(elems)))
                raw = raw[:elems.value] if raw_view else raw[:elems.value].copy()
            else:
                self._check(self.lib.tps_wait(self._h, bid, rows.ctypes.data, C.byref(n_pass), None, 0,
                                              C.byref(elems)))
        finally:
            del self._inflight[bid]
        return rows, raw

    def wait_leased(self, bid: int):
        """wait() for pipelines: (rows, raw tables as a view of a leased landing buffer or None, RawLease or None).
        The view is valid until the lease is released; leases of several finished batches may be out at once."""
        if not self.want_rawcount:
            rows, _ = self.wait(bid)
            return rows, None, None
        bases, offsets, n_reads = self._inflight[bid]
        rows = np.empty(n_reads, dtype=ROW_DTYPE)
        n_pass = C.c_uint32(0)
        elems = C.c_uint64(0)
        buf = None
        try:
            self._check(self.lib.tps_batch_info(self._h, bid, C.byref(n_pass), C.byref(elems)))
            need = min(int(elems.value), int(self.params.rawcount_capacity))
            fit = [b for b in self._raw_pool if b.nbytes >= need]
            if fit:
                buf = min(fit, key=lambda b: b.nbytes)
                self._raw_pool.remove(buf)
            else:
                buf = PinnedBuffer(max(need * 5 // 4, 1 << 22))
            raw = buf.array[:need]
            self._check(self.lib.tps_wait(self._h, bid, rows.ctypes.data, C.byref(n_pass), raw.ctypes.data, raw.size,
                                          C.byref(elems)))
            lease, buf = RawLease(self, buf), None
            return rows, raw[:elems.value], lease
        finally:
            if buf is not None:
                self._raw_pool.append(buf)
            del self._inflight[bid]

    def scan(self, bases: np.ndarray, offsets: np.ndarray):
        """Synchronous scan of one host batch -> (rows, rawcounts or None)."""
        return self.wait(self.submit(bases, offsets))

    def scan_reads(self, seqs: Iterable):
        bases, offsets = pack_reads(seqs)
        return self.scan(bases, offsets)

    def scan_reads_ends_first(self, seqs: Iterable):
        """The ends-first protocol on in-memory reads: step 1 on head + tail of every read (tps_submit_ends),
        steps 2/3 on the first / last min(L, maxlengthtelo) bases of the TRC-pass reads (tps_submit_regions),
        rows merged.  Returns (rows, tables) with tables[i] = counts[w][p] of read i or None; equal to
        scan_reads() field by field (raw-count offsets aside).  pipeline.Scanner(ends_first=True) does the same
        from a file without ever copying the interior of a read."""
        bufs = [s.encode("ascii", "replace") if isinstance(s, str) else bytes(s) for s in seqs]
        h = int(self.params.no_bp)
        minis = [b if len(b) <= 2 * h else b[:h] + b[len(b) - h:] for b in bufs]
        bases, offsets = pack_reads(minis)
        lens = np.diff(offsets).astype(np.uint32)
        true_lens = np.array([len(b) for b in bufs], dtype=np.uint32)
        rows, _ = self.wait(self.submit_ends(bases, np.ascontiguousarray(offsets[:-1]), lens, true_lens, len(bufs)))
        tables = [None] * len(bufs)
        idx = np.nonzero(rows["status"] == ST_PASS)[0]
        if len(idx) == 0 or (self.params.flags & 1):       # TPS_FLAG_STEP1_ONLY
            return rows, tables
        m = int(self.params.maxlengthtelo)
        regs = []
        for i in idx:
            b = bufs[i]
            k = min(len(b), m)
            regs.append(b[:k] if rows["tail"][i] == 0 else b[len(b) - k:])
        rb, ro = pack_reads(regs)
        tails = np.ascontiguousarray(rows["tail"][idx].astype(np.uint8))
        rows2, raw2 = self.wait(self.submit_regions(rb, np.ascontiguousarray(ro[:-1]), np.diff(ro).astype(np.uint32),
                                                    tails, len(regs)))
        rows = rows.copy()
        for j, i in enumerate(idx):
            for f in ("status", "n_windows", "bkp", "telo_length"):
                rows[f][i] = rows2[f][j]
            tables[i] = self.rawcount_table(rows2, raw2, j)
            if tables[i] is not None:
                tables[i] = tables[i].copy()
        return rows, tables

    def rawcount_table(self, rows, raw, i) -> np.ndarray:
        """counts[w][p] (uint8) of read i, or None."""
        off = int(rows["rawcount_offset"][i])
        if raw is None or off == NO_RAWCOUNT:
            return None
        nw, npat = int(rows["n_windows"][i]), len(self.patterns)
        return raw[off:off + nw * npat].reshape(nw, npat)

    # -- device-resident path (kernel-only timing)
    def scan_device(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int, d_rows_ptr: int,
                    slot: int = 0):
        self._check(self.lib.tps_scan_device_slot(self._h, slot, d_bases_ptr, d_offsets_ptr, n_reads, n_bases,
                                                  d_rows_ptr))

    def sync(self):
        self._check(self.lib.tps_sync(self._h))

    def timings(self, back: int = 0):
        """CUDA-event stage times (ms) of the scan_device call `back` calls ago."""
        ms = (C.c_float * 4)()
        self._check(self.lib.tps_get_timings(self._h, back, C.byref(ms)))
        return dict(k1_pack=ms[0], k2_trc=ms[1], k3_windows_cp=ms[2], total=ms[3])

    def timeline(self, back: int, base_back: int):
        """Event times (ms) of the scan `back` calls ago relative to the start of the scan `base_back` calls ago:
        (K1 start, K1 end, K2 end, K4 end)."""
        ms = (C.c_float * 4)()
        self._check(self.lib.tps_get_timeline(self._h, back, base_back, C.byref(ms)))
        return tuple(float(x) for x in ms)

    def elapsed_to(self, back: int, event: int, other: "ScanContext", other_back: int, other_event: int) -> float:
        """Device time (ms) from event `event` (0 K1 start .. 3 scan end) of this context's scan_device call `back`
        calls ago to event `other_event` of `other`'s call `other_back` calls ago (tps_elapsed_between)."""
        ms = C.c_float()
        self._check(self.lib.tps_elapsed_between(self._h, back, event, other._h, other_back, other_event, C.byref(ms)))
        return float(ms.value)

    def kernel_launches(self) -> int:
        return int(self.lib.tps_kernel_launches(self._h))

    def debug_info(self) -> dict:
        v = self.debug_copy(5, 20).view(np.uint32)
        return dict(cw_stride=int(v[0]), k3_bitpar=bool(v[1]), max_pass=int(v[2]), k3_tile_bases=int(v[3]),
                    k2_kernel=("smem", "reg", "const")[int(v[4])])

    def window_sums(self, rows) -> dict:
        """Test hook: {read index: sums of c_w over the groups of five windows [5j, 5j+5)} of the TRC-pass reads
        of the last batch scanned on slot 0.  The bit-parallel kernel keeps exactly these (in shared memory; it
        writes them out only if the context was created under TPS_K3_DEBUG_GS=1); the plain kernel's per-window
        c_w are summed here."""
        n_pass = int((rows["status"] >= ST_PASS).sum())
        if n_pass == 0:
            return {}
        info = self.debug_info()
        stride, dt = info["cw_stride"], (np.uint16 if info["k3_bitpar"] else np.uint32)
        plist = self.debug_copy(3, n_pass * 4).view(np.uint32)
        cw = self.debug_copy(4, n_pass * stride * np.dtype(dt).itemsize).view(dt).reshape(n_pass, stride)
        out = {}
        for i, r in enumerate(plist):
            nw = int(rows["n_windows"][r])
            if nw < 7:
                continue
            out[int(r)] = cw[i, :(nw + 4) // 5].astype(np.int64) if info["k3_bitpar"] else group_sums(cw[i, :nw])
        return out

    def debug_copy(self, what: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        self._check(self.lib.tps_debug_copy(self._h, what, out.ctypes.data, nbytes))
        return out
