"""Streaming FASTQ / FASTA (.gz) input -- ctypes wrapper over csrc/tps_fastx.c.

Takes the place of `check_file_type` / `unzip_file` + `Bio.SeqIO.parse`
(Topsicle/allsteps.py:36-50, 127-149) on the scan path: records are indexed in C (several
threads on plain files, zlib on `.gz`) and their bases land back to back in a pinned batch
buffer that `tps_submit` uploads as is.  Python never materialises a per-read object;
ids / titles / qualities are sliced out of the raw text only for the reads that need them
(the TRC-pass reads written to the `_trc_over_` subset, Topsicle/main.py:68-86).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libtps_host.so")

FASTQ, FASTA = 1, 2
FORMAT_NAMES = {FASTQ: "fastq", FASTA: "fasta"}

REC_DTYPE = np.dtype([
    ("title_off", "<u8"), ("seq_off", "<u8"), ("qual_off", "<u8"), ("title_len", "<u4"), ("id_off", "<u4"),
    ("id_len", "<u4"), ("seq_len", "<u4"), ("seq_raw_len", "<u4"), ("flags", "<u4"),
])
assert REC_DTYPE.itemsize == 48


class FastxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fastx error {code}: {msg}")
        self.code = code


_lib = None


def host_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} not found: run __graft_entry__.build()")
        # the parser's OpenMP team shares the cores with the Python reader / worker threads (and, under
        # torchrun, with the other ranks): idle team members must sleep, not spin (read by libgomp when it loads)
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
        lib = C.CDLL(HOST_LIB_PATH)
        vp = C.c_void_p
        lib.tps_fastx_open.restype = C.c_int
        lib.tps_fastx_open.argtypes = [C.POINTER(vp), C.c_char_p, C.c_int]
        lib.tps_fastx_close.restype = None
        lib.tps_fastx_close.argtypes = [vp]
        lib.tps_fastx_format.restype = C.c_int
        lib.tps_fastx_format.argtypes = [vp]
        lib.tps_fastx_set_window.restype = None
        lib.tps_fastx_set_window.argtypes = [vp, C.c_uint64]
        lib.tps_fastx_set_clip.restype = None
        lib.tps_fastx_set_clip.argtypes = [vp, C.c_uint64]
        lib.tps_fastx_release.restype = None
        lib.tps_fastx_release.argtypes = [vp]
        lib.tps_fastx_last_error.restype = C.c_char_p
        lib.tps_fastx_last_error.argtypes = [vp]
        lib.tps_fastx_next.restype = C.c_int
        lib.tps_fastx_next.argtypes = [vp, C.c_uint64, C.c_uint32, vp, vp, vp, C.POINTER(C.c_uint32),
                                       C.POINTER(vp), C.POINTER(vp)]
        lib.tps_fastx_next_spans.restype = C.c_int
        lib.tps_fastx_next_spans.argtypes = [vp, C.c_uint64, C.c_uint32, vp, vp, vp, vp, C.POINTER(C.c_uint32),
                                             C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp)]
        lib.tps_fastx_next_ends.restype = C.c_int
        lib.tps_fastx_next_ends.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp,
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                            C.POINTER(vp), C.POINTER(vp)]
        lib.tps_fastx_set_two_pass.restype = None
        lib.tps_fastx_set_two_pass.argtypes = [vp, C.c_int]
        lib.tps_fastx_inflate_stats.restype = None
        lib.tps_fastx_inflate_stats.argtypes = [vp, vp]
        lib.tps_fastx_find_id.restype = C.c_uint32
        lib.tps_fastx_find_id.argtypes = [vp, vp, C.c_uint32, C.c_char_p, C.c_uint32, vp, C.c_uint32]
        lib.tps_fastx_gather_regions.restype = C.c_uint32
        lib.tps_fastx_gather_regions.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, vp]
        lib.tps_fastx_records_text.restype = C.c_int64
        lib.tps_fastx_records_text.argtypes = [vp, vp, vp, C.c_uint32, C.c_int, vp, C.c_uint64, vp]
        lib.tps_fastx_join_ids.restype = C.c_int64
        lib.tps_fastx_join_ids.argtypes = [vp, vp, vp, C.c_uint32, vp, C.c_uint64]
        lib.tps_format_rawcount.restype = C.c_int64
        lib.tps_format_rawcount.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p,
                                            C.POINTER(C.c_char_p), vp, C.c_uint64]
        _lib = lib
    return _lib


class Batch:
    """One batch of reads: `bases[offsets[i]:offsets[i+1]]` is read i; `recs[i]` indexes its raw
    text.  `first_read` = index of read 0 within its file.  Keep the batch (and its FastxFile)
    alive while slicing raw text; call `release()` when done."""

    def __init__(self, lib, n_reads, bases, offsets, recs, raw_base, raw_owner, first_read, fmt, lens=None,
                 span=None):
        self._lib = lib
        self.n_reads = n_reads
        self.bases = bases
        self.offsets = offsets          # read starts (n_reads + 1 entries when the reads are back to back)
        self.lens = lens                # uint32 read lengths of a span batch, None for a back-to-back batch
        self.span = span if span is not None else int(offsets[n_reads])   # bytes of `bases` in use
        self.recs = recs
        self._raw = raw_base
        self._owner = raw_owner
        self.first_read = first_read
        self.format = fmt

    @property
    def n_bases(self) -> int:
        if self.lens is not None:
            return int(self.lens[:self.n_reads].sum(dtype=np.uint64))
        return int(self.offsets[self.n_reads])

    def bounds(self, i):
        a = int(self.offsets[i])
        return (a, a + int(self.lens[i])) if self.lens is not None else (a, int(self.offsets[i + 1]))

    def _text(self, off, n) -> bytes:
        return C.string_at(self._raw + int(off), int(n))

    def title(self, i) -> str:
        r = self.recs[i]
        return self._text(r["title_off"], r["title_len"]).decode("utf-8", "replace")

    def read_id(self, i) -> str:
        r = self.recs[i].item()   # (title_off, seq_off, qual_off, title_len, id_off, id_len, ...)
        return C.string_at(self._raw + r[0] + r[4], r[5]).decode("utf-8", "replace")

    def read_ids(self, indices) -> list:
        """Ids of several reads in one call (the harvest of a batch asks for all its TRC-pass reads at once)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        if idx.size == 0:
            return []
        recs = np.ascontiguousarray(self.recs)
        cap = int(recs["id_len"][idx].sum(dtype=np.uint64)) + idx.size
        out = np.empty(cap, dtype=np.uint8)
        n = self._lib.tps_fastx_join_ids(self._raw, recs.ctypes.data, idx.ctypes.data, idx.size, out.ctypes.data, cap)
        if n < 0:
            raise FastxError(-4, "id buffer too small")
        return out[:n - 1].tobytes().decode("utf-8", "replace").split("\n")

    def sequence(self, i) -> bytes:
        a, b = self.bounds(i)
        return self.bases[a:b].tobytes()

    def quality(self, i) -> bytes:
        r = self.recs[i]
        return self._text(r["qual_off"], r["seq_len"])

    def record_text(self, i) -> bytes:
        """The record as `SeqIO.write(record, handle, fmt)` emits it (main.py:86): FASTQ
        `@title\\nseq\\n+\\nqual\\n`; FASTA `>title\\n` + sequence wrapped at 60 columns."""
        title_off, seq_off, qual_off, title_len, _, _, seq_len, seq_raw_len, flags = self.recs[i].item()
        if (self.format == FASTQ and not (flags & 1) and seq_raw_len == seq_len
                and seq_off == title_off + title_len + 1 and qual_off == seq_off + seq_len + 3):
            # the record already has SeqIO.write's shape in the file (bare '+' line, nothing stripped):
            # one slice of the raw text instead of four copies
            return self._text(title_off - 1, qual_off + seq_len - title_off + 1) + b"\n"
        title = self._text(title_off, title_len)
        seq = self.sequence(i)
        if self.format == FASTQ:
            return b"@" + title + b"\n" + seq + b"\n+\n" + self.quality(i) + b"\n"
        lines = [seq[j:j + 60] for j in range(0, len(seq), 60)]
        return b">" + title + b"\n" + b"".join(ln + b"\n" for ln in lines)

    def records_text(self, indices) -> list:
        """`record_text` of several reads in one call: a list of memoryviews into one buffer (what the subset file
        writer needs for all TRC-pass reads of a batch)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        if idx.size == 0:
            return []
        recs = np.ascontiguousarray(self.recs)
        sel = recs[idx]
        L = sel["seq_len"].astype(np.uint64)
        cap = int((sel["title_len"].astype(np.uint64) + 2 * L + L // 60 + 8).sum())
        out = np.empty(cap, dtype=np.uint8)
        ends = np.empty(idx.size, dtype=np.uint64)
        n = self._lib.tps_fastx_records_text(self._raw, recs.ctypes.data, idx.ctypes.data, idx.size, self.format,
                                             out.ctypes.data, cap, ends.ctypes.data)
        if n < 0:
            raise FastxError(-4, "record text buffer too small")
        view = memoryview(out)
        cuts = [0] + ends.tolist()
        return [view[cuts[j]:cuts[j + 1]] for j in range(idx.size)]

    def find_id(self, read_id: str) -> list:
        """Indices of the records whose id equals `read_id` (the reference's `seq.id != read` test)."""
        key = read_id.encode("utf-8")
        recs = np.ascontiguousarray(self.recs)
        out = np.empty(max(1, self.n_reads), dtype=np.uint32)
        n = self._lib.tps_fastx_find_id(self._raw, recs.ctypes.data, self.n_reads, key, len(key), out.ctypes.data,
                                        out.size)
        return [int(i) for i in out[:n]]

    def release(self):
        if self._owner:
            self._lib.tps_fastx_release(self._owner)
            self._owner = None
        self._raw = None


class EndsBatch(Batch):
    """A batch of the ends-first reader: `bases` holds head + tail of every read (the whole read when it is no
    longer than 2 * end_len), `true_lens[i]` its real length; the full sequence is read from the raw text."""

    def __init__(self, *a, true_lens=None, true_bases=0, end_len=0, **kw):
        super().__init__(*a, **kw)
        self.true_lens = true_lens
        self.true_bases = int(true_bases)
        self.end_len = end_len

    @property
    def n_bases(self) -> int:          # bases of the reads themselves (what was scanned), not what was uploaded
        return self.true_bases

    def sequence(self, i) -> bytes:
        r = self.recs[i]
        if not (int(r["flags"]) & 1):
            return self._text(r["seq_off"], r["seq_len"])
        raw = self._text(r["seq_off"], r["seq_raw_len"])
        return raw.translate(None, b" \r\n\t\x0b\x0c")

    def gather_regions(self, indices, tails, maxlengthtelo: int, dst: np.ndarray, starts: np.ndarray,
                       lens: np.ndarray) -> int:
        """Copy the regions (see `region`) of reads `indices` back to back into `dst`, filling `starts` / `lens`;
        returns how many fitted (a prefix).  One C call (threads) instead of two copies per read in Python."""
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        tl = np.ascontiguousarray(tails, dtype=np.uint8)
        assert dst.dtype == np.uint8 and starts.dtype == np.uint64 and lens.dtype == np.uint32
        assert len(starts) >= idx.size and len(lens) >= idx.size and tl.size == idx.size
        recs = np.ascontiguousarray(self.recs)
        return int(self._lib.tps_fastx_gather_regions(self._raw, recs.ctypes.data, idx.ctypes.data, tl.ctypes.data,
                                                      idx.size, int(maxlengthtelo), dst.size, dst.ctypes.data,
                                                      starts.ctypes.data, lens.ctypes.data))

    def region(self, i, tail: int, maxlengthtelo: int) -> bytes:
        """The bases steps 2/3 look at: first (tail 0) or last (tail 1) min(L, maxlengthtelo) bases."""
        r = self.recs[i]
        L = int(r["seq_len"])
        k = min(L, int(maxlengthtelo))
        if not (int(r["flags"]) & 1):
            return self._text(int(r["seq_off"]) + (0 if tail == 0 else L - k), k)
        s = self.sequence(i)
        return s[:k] if tail == 0 else s[L - k:]


class FastxFile:
    """An open FASTQ / FASTA file (gzip if the name ends in `.gz`, as the reference decides)."""

    def __init__(self, path: str, threads: int = 0):
        self._lib = host_library()
        self.path = path
        self._h = C.c_void_p()
        threads = threads or len(os.sched_getaffinity(0))
        rc = self._lib.tps_fastx_open(C.byref(self._h), os.fsencode(path), threads)
        if rc != 0:
            raise FastxError(rc, self._lib.tps_fastx_last_error(None).decode())
        self.format = self._lib.tps_fastx_format(self._h)
        self.format_name = FORMAT_NAMES[self.format]
        self.reads_delivered = 0

    def set_window(self, nbytes: int):
        self._lib.tps_fastx_set_window(self._h, nbytes)

    def set_clip(self, clip_bases: int):
        """A record longer than a whole batch is delivered as its first + last `clip_bases` bases (a batch of its
        own) instead of ending the file with error -4; see include/topsicle_host.h."""
        self._lib.tps_fastx_set_clip(self._h, int(clip_bases))

    def next_batch(self, bases: np.ndarray, offsets: np.ndarray, max_reads: int | None = None,
                   max_bases: int | None = None, recs: np.ndarray | None = None) -> Batch | None:
        """Fill `bases` (uint8) / `offsets` (uint64, >= max_reads + 1) with the next reads of the
        file; returns None at end of file."""
        assert bases.dtype == np.uint8 and offsets.dtype == np.uint64
        reads_cap = min(len(offsets) - 1, max_reads if max_reads is not None else 1 << 31)
        bases_cap = min(bases.size, max_bases if max_bases is not None else 1 << 62)
        if recs is None:
            recs = np.empty(reads_cap, dtype=REC_DTYPE)
        else:
            reads_cap = min(reads_cap, len(recs))
        n = C.c_uint32(0)
        raw, owner = C.c_void_p(), C.c_void_p()
        rc = self._lib.tps_fastx_next(self._h, bases_cap, reads_cap, bases.ctypes.data, offsets.ctypes.data,
                                      recs.ctypes.data, C.byref(n), C.byref(raw), C.byref(owner))
        if rc != 0:
            raise FastxError(rc, self._lib.tps_fastx_last_error(self._h).decode())
        if n.value == 0:
            return None
        b = Batch(self._lib, n.value, bases, offsets, recs[:n.value], raw.value, owner.value,
                  self.reads_delivered, self.format)
        self.reads_delivered += n.value
        return b

    def next_spans(self, bases: np.ndarray, starts: np.ndarray, lens: np.ndarray, max_reads: int | None = None,
                   max_span: int | None = None, recs: np.ndarray | None = None) -> Batch | None:
        """Next batch as a span batch (`tps_submit_spans`): read i = bases[starts[i] : starts[i] + lens[i]].
        FASTQ takes the one-pass reader (reads land at half their file offset, small gaps in between);
        FASTA falls back to back-to-back packing.  `starts` needs max_reads + 1 entries."""
        assert bases.dtype == np.uint8 and starts.dtype == np.uint64 and lens.dtype == np.uint32
        reads_cap = min(len(starts) - 1, len(lens), max_reads if max_reads is not None else 1 << 31)
        span_cap = min(bases.size, max_span if max_span is not None else 1 << 62)
        if recs is None:
            recs = np.empty(reads_cap, dtype=REC_DTYPE)
        else:
            reads_cap = min(reads_cap, len(recs))
        n, span = C.c_uint32(0), C.c_uint64(0)
        raw, owner = C.c_void_p(), C.c_void_p()
        rc = self._lib.tps_fastx_next_spans(self._h, span_cap, reads_cap, bases.ctypes.data, starts.ctypes.data,
                                            lens.ctypes.data, recs.ctypes.data, C.byref(n), C.byref(span),
                                            C.byref(raw), C.byref(owner))
        if rc != 0:
            raise FastxError(rc, self._lib.tps_fastx_last_error(self._h).decode())
        if n.value == 0:
            return None
        b = Batch(self._lib, n.value, bases, starts, recs[:n.value], raw.value, owner.value, self.reads_delivered,
                  self.format, lens=lens, span=span.value)
        self.reads_delivered += n.value
        return b

    def next_ends(self, bases: np.ndarray, starts: np.ndarray, lens: np.ndarray, true_lens: np.ndarray, end_len: int,
                  raw_cap: int = 1 << 30, max_reads: int | None = None, max_bases: int | None = None,
                  recs: np.ndarray | None = None) -> EndsBatch | None:
        """Next batch of the ends-first reader (`tps_submit_ends`): read i contributes its first and last
        `end_len` bases, packed back to back; at most `raw_cap` bytes of file text per batch."""
        assert bases.dtype == np.uint8 and starts.dtype == np.uint64 and lens.dtype == np.uint32
        assert true_lens.dtype == np.uint32
        reads_cap = min(len(starts), len(lens), len(true_lens), max_reads if max_reads is not None else 1 << 31)
        bases_cap = min(bases.size, max_bases if max_bases is not None else 1 << 62)
        if recs is None:
            recs = np.empty(reads_cap, dtype=REC_DTYPE)
        else:
            reads_cap = min(reads_cap, len(recs))
        n, span, tb = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        raw, owner = C.c_void_p(), C.c_void_p()
        rc = self._lib.tps_fastx_next_ends(self._h, raw_cap, bases_cap, reads_cap, end_len, bases.ctypes.data,
                                           starts.ctypes.data, lens.ctypes.data, true_lens.ctypes.data,
                                           recs.ctypes.data, C.byref(n), C.byref(span), C.byref(tb), C.byref(raw),
                                           C.byref(owner))
        if rc != 0:
            raise FastxError(rc, self._lib.tps_fastx_last_error(self._h).decode())
        if n.value == 0:
            return None
        b = EndsBatch(self._lib, n.value, bases, starts, recs[:n.value], raw.value, owner.value, self.reads_delivered,
                      self.format, lens=lens, span=span.value, true_lens=true_lens, true_bases=tb.value,
                      end_len=end_len)
        self.reads_delivered += n.value
        return b

    def inflate_stats(self) -> dict:
        """Counters of the parallel gzip inflater (all zero unless the file is plain gzip): stretches, segments,
        chain_breaks, members, text_bytes, parallel_text_bytes."""
        out = (C.c_uint64 * 6)()
        self._lib.tps_fastx_inflate_stats(self._h, out)
        return dict(zip(("stretches", "segments", "chain_breaks", "members", "text_bytes", "parallel_text_bytes"),
                        (int(x) for x in out)))

    def set_two_pass(self, on: bool = True):
        """Force the index-then-gather reader (the one-pass FASTQ reader's validating fallback)."""
        self._lib.tps_fastx_set_two_pass(self._h, 1 if on else 0)

    def close(self):
        if self._h and self._h.value:
            self._lib.tps_fastx_close(self._h)
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


def sniff_format(path: str) -> str | int:
    """`check_file_type` (allsteps.py:36-50): 'fastq' / 'fasta', or 0 if it cannot be told."""
    try:
        with FastxFile(path, threads=1) as fx:
            return fx.format_name
    except (FastxError, OSError):
        return 0


def format_rawcount_csv(counts: np.ndarray, slide: int, tail: str, patterns) -> bytes:
    """`rawCountPattern(...).to_csv()` text (allsteps.py:401-416,464; main.py:150) from a
    uint8 [n_windows][n_patterns] count table."""
    lib = host_library()
    counts = np.ascontiguousarray(counts, dtype=np.uint8)
    nw, npat = (counts.shape if counts.ndim == 2 else (0, len(patterns)))
    pats = (C.c_char_p * len(patterns))(*[p.encode() for p in patterns])
    cap = 64 + nw * npat * (48 + len(tail) + max((len(p) for p in patterns), default=0))
    out = C.create_string_buffer(cap)
    n = lib.tps_format_rawcount(counts.ctypes.data, nw, npat, slide, tail.encode(), pats, out, cap)
    if n < 0:
        raise FastxError(-4, f"rawcount buffer too small ({cap} < {-n})")
    return out.raw[:n]
