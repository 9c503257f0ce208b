"""Drop-in for `Topsicle.allsteps` -- same names, argument meaning and return shapes, with the
per-read arithmetic done by the sm_100a kernels behind `libtopsicle_b200.so`.

    from topsicle_b200.allsteps import *      # instead of: from Topsicle.allsteps import *

Function by function (reference file:line in Topsicle/allsteps.py):
  pattern_scramble_telo   :57-82    host (patterns.py)
  patterns_to_search      :84-125   host (patterns.py); '|' patterns are rejected
  check_file_type         :36-50    C reader sniff
  unzip_file              :127-149  C reader, yields lightweight records
  patternTRC_count        :152-204  GPU: K1 pack + K2 TRC (step 1 only)
  seq_cut_windows         :207-225  host helper (the kernels never build window strings)
  bound_detect            :227-338  GPU: K1 + K2 (tail forced by the caller) + K3 + K4
  rawCountPattern         :359-464  GPU: same scan, returns the count table as a DataFrame
  fit_quadratic_and_find_vertex :467-502  host NumPy (a handful of points)
There is no CPU implementation of the scan in this package: without the CUDA library or a
GPU these functions raise.
"""
from __future__ import annotations

import logging
import os

import numpy as np

from . import engine, fastx, pipeline
from .patterns import pattern_scramble_telo, patterns_to_search, validate_literals  # noqa: F401

__all__ = ["check_file_type", "pattern_scramble_telo", "patterns_to_search", "unzip_file", "patternTRC_count",
           "seq_cut_windows", "bound_detect", "rawCountPattern", "fit_quadratic_and_find_vertex", "plot_patterns",
           "version_number"]

version_number = "1.0.0"


def _devices():
    """GPUs used by the per-call API: TOPSICLE_DEVICES="0,1,..." or device 0."""
    env = os.environ.get("TOPSICLE_DEVICES")
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    return [0]


# ----------------------------------------------------------------------------------- input
def check_file_type(filepath):
    """'fastq' / 'fasta' by the first character of the file, 0 if neither (allsteps.py:36-50)."""
    kind = fastx.sniff_format(filepath)
    if not kind:
        logging.warning("Format cannot be identified. Check the input.")
    return kind


class _Seq(str):
    def __getitem__(self, key):
        out = str.__getitem__(self, key)
        return _Seq(out) if isinstance(key, slice) else out

    def upper(self):
        return _Seq(str.upper(self))


class Record:
    """What the reference reads off a Bio.SeqRecord: id, name, description, seq, len()."""
    __slots__ = ("id", "name", "description", "seq", "quality")

    def __init__(self, title, rid, seq, quality=None):
        self.description = title
        self.id = self.name = rid
        self.seq = _Seq(seq)
        self.quality = quality

    def __len__(self):
        return len(self.seq)


def unzip_file(filepath):
    """Yield the records of a FASTQ / FASTA file, gzip-compressed or not (allsteps.py:127-149)."""
    if not isinstance(filepath, str):
        logging.error("Input must be a string representing the file path.")
        return None
    try:
        fx = fastx.FastxFile(filepath)
    except (fastx.FastxError, OSError) as e:
        logging.error(f"File type could not be determined or is unsupported: {e}")
        return None
    bases = np.empty(1 << 26, dtype=np.uint8)
    offsets = np.empty((1 << 16) + 1, dtype=np.uint64)
    try:
        while True:
            try:
                b = fx.next_batch(bases, offsets)
            except fastx.FastxError as e:
                if e.code != -4:
                    raise
                bases = np.empty(bases.size * 4, dtype=np.uint8)   # a read longer than the buffer
                continue
            if b is None:
                return
            for i in range(b.n_reads):
                qual = b.quality(i).decode("ascii", "replace") if b.format == fastx.FASTQ else None
                yield Record(b.title(i), b.read_id(i), b.sequence(i).decode("ascii", "replace"), qual)
            b.release()
    finally:
        fx.close()


# ----------------------------------------------------------------------------------- step 1
def patternTRC_count(filepath, telopattern, read_length=0, kmer=4, no_bp=1000, cutoff=0.5):
    """Telomere-like repeat count of both read ends (allsteps.py:152-204).

    Returns `[[read_id, literal, 'forward'|'reverse', trc], ...]` in file order for the reads
    with `len > read_length` whose TRC exceeds `cutoff`."""
    if isinstance(filepath, list):
        print("Can only process 1 file path at the time, please loop paths through the list")
        return None
    literals = patterns_to_search(telopattern, cut_length=kmer)
    validate_literals(literals)
    if not check_file_type(filepath):
        logging.error("File type could not be determined or is unsupported.")
        return []
    cfg = pipeline.ScanConfig(patterns=literals, len_telopattern=len(telopattern), phrase=kmer, cutoff=cutoff,
                              min_seq_length=read_length, no_bp=no_bp, step1_only=True)
    # step 1 reads nothing but the first / last `no_bp` bases of a read: the ends-first reader uploads just those
    _, per_cfg = pipeline.collect_file(filepath, [cfg], devices=_devices(), ends_first=True)
    return [[p.read_id, p.literal, p.tail, p.trc] for p in per_cfg[0]]


# ----------------------------------------------------------------------------------- step 2
def seq_cut_windows(s, window_size, step):
    """Window starts `range(0, len(s) - window_size + 1, step)` with the text `s[i : i+window_size-1]`
    (W-1 characters, allsteps.py:207-225)."""
    return [(i, s[i:min(i + window_size - 1, len(s))]) for i in range(0, len(s) - window_size + 1, step)]


def _scan_one_read(filepath, read, pattern_telo, windowSize, slide, trimfirst, maxlengthtelo, cut_length, tail,
                   want_rawcount):
    """Scan the record(s) named `read` with the caller-chosen tail(s): reverse first, then forward,
    the order in which bound_detect / rawCountPattern report them (allsteps.py:335-336, 418-419)."""
    literals = patterns_to_search(pattern_telo, cut_length=cut_length)
    validate_literals(literals)
    tails = [tail] if tail in ("forward", "reverse") else ["reverse", "forward"]
    cfgs = [pipeline.ScanConfig(patterns=literals, len_telopattern=max(1, len(literals[0])), phrase=cut_length,
                                min_seq_length=0, count_threshold_override=0, window_size=windowSize, slide=slide,
                                trimfirst=trimfirst, maxlengthtelo=maxlengthtelo, want_rawcount=want_rawcount,
                                force_tail=t) for t in tails]
    per_cfg = pipeline.scan_named_read(filepath, cfgs, read, device=_devices()[0])
    return literals, tails, per_cfg


def bound_detect(filepath, read, pattern_telo, windowSize, slide, trimfirst, maxlengthtelo, cut_length, tail=None,
                 plot_yes_no=None, plotcp_range=None):
    """Telomere / subtelomere boundary of one read (allsteps.py:227-338): `[[read, boundary]]`.

    boundary = trimfirst + slide * b, b the single l2 change point of the mean window count
    (ruptures Binseg, jump 5, min_size 2).  With `tail=None` both ends are reported, reverse
    first.  Fewer than 7 windows is an error (the reference raises inside ruptures)."""
    if not isinstance(read, str):
        print("can only read in 1 read at a time")
        return None
    if windowSize is None:
        return []
    if not check_file_type(filepath):
        print("Problem in filepath, please double check")
        return ["didn't run", filepath, 0]
    want_plot = bool(plot_yes_no)
    literals, tails, per_cfg = _scan_one_read(filepath, read, pattern_telo, windowSize, slide, trimfirst,
                                              maxlengthtelo, cut_length, tail, want_rawcount=want_plot)
    by_index = {}
    for t, passes in zip(tails, per_cfg):
        for p in passes:
            by_index.setdefault(p.index, []).append((t, p))
    boundary = []
    for idx in sorted(by_index):
        for t, p in by_index[idx]:
            if p.n_windows == 0:
                continue                      # `if not x: return` (allsteps.py:307-308)
            if p.status != engine.ST_PASS:
                raise ValueError(f"read {read}: {p.n_windows} windows are too few for a change point "
                                 "(ruptures.BadSegmentationParameters in the reference)")
            m = min(maxlengthtelo, p.length)
            point = int(p.telo_length)
            if want_plot and p.counts is not None:
                _plot_boundary(read, p, len(literals), trimfirst, slide, m, point, plotcp_range)
            boundary.append([read, point] if (point <= m and point != 0) else [read, 0])
    return boundary


def _plot_boundary(read, p, n_pat, trimfirst, slide, maxlengthtelo, point, plotcp_range):
    """The --plot figure of bound_detect (allsteps.py:317-328); needs matplotlib."""
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        logging.error("matplotlib is not installed: --plot figure skipped")
        return
    y = p.counts.astype(np.int64).sum(axis=1) / n_pat
    x = trimfirst + slide * np.arange(len(y))
    plt.figure(figsize=(7.5, 3), dpi=300)
    plt.plot(x, y, color="#000000", linestyle="-", linewidth=2)
    plt.axvline(x=point, color="#FF2C2C", linewidth=2, linestyle="--", label=f"x = boundary point: {point}")
    plt.title(f"mean window + boundary point of {read}")
    plt.xlabel("base pair (bp)")
    plt.ylabel("mean window value")
    plt.xlim(0, plotcp_range if plotcp_range else maxlengthtelo)
    plt.tight_layout()
    plt.grid(True)


def plot_patterns(seq, patterns, read_ids, added_labels, ax, direction):
    """Scatter helper of the (dead) raw-count plot branch (allsteps.py:340-357): only reachable
    with plot_raw=True, which rawCountPattern forces to False (allsteps.py:421)."""
    raise NotImplementedError("plot_patterns is unreachable in the reference (plot_raw is forced to False)")


# ----------------------------------------------------------------------------------- step 3
def rawcount_frame(counts: np.ndarray, tail: str, slide: int, literals):
    """DataFrame(tail, position, pattern, count) rows, window-major / pattern-minor
    (allsteps.py:401-416, 464), from a [n_windows][n_patterns] count table."""
    import pandas as pd
    nw, npat = counts.shape
    return pd.DataFrame({
        "tail": np.repeat(np.array([tail], dtype=object), nw * npat),
        "position": np.repeat(np.arange(nw, dtype=np.int64) * slide, npat),
        "pattern": np.tile(np.array(list(literals), dtype=object), nw),
        "count": counts.astype(np.int64).reshape(-1),
    }, columns=["tail", "position", "pattern", "count"])


def rawCountPattern(filepath, read, pattern_telo, windowSize, slide, trimfirst, cut_length, minSeqLength,
                    maxlengthtelo, tail=None, plot_raw=False):
    """Per-window, per-literal match counts of one read (allsteps.py:359-464) as a DataFrame with
    columns tail / position / pattern / count.  `position` is the window start inside the trimmed
    region (trimfirst not added).  With `tail=None`: forward rows, then reverse rows."""
    import pandas as pd
    if not isinstance(read, str):
        print("can only read in 1 read at a time")
        return None
    if not check_file_type(filepath):
        print("Problem in filepath, please double check")
        return ["didn't run", filepath, 0]
    cols = ["tail", "position", "pattern", "count"]
    if windowSize is None:
        return pd.DataFrame([], columns=cols)
    literals, tails, per_cfg = _scan_one_read(filepath, read, pattern_telo, windowSize, slide, trimfirst,
                                              maxlengthtelo, cut_length, tail, want_rawcount=True)
    by_index = {}
    for t, passes in zip(tails, per_cfg):
        for p in passes:
            by_index.setdefault(p.index, {})[t] = p
    frames = []
    for idx in sorted(by_index):
        for t in ("forward", "reverse"):           # rawpattern_s then rawpattern_e (allsteps.py:418-419)
            p = by_index[idx].get(t)
            if p is not None and p.counts is not None and p.n_windows:
                frames.append(rawcount_frame(p.counts, t, slide, literals))
    if not frames:
        return pd.DataFrame([], columns=cols)
    return frames[0] if len(frames) == 1 else pd.concat(frames, ignore_index=True)


# ----------------------------------------------------------------------------------- summary
def fit_quadratic_and_find_vertex(trc_list, telo_length_list, inputtrc, median_trc, save_path=None):
    """Quadratic fit of telomere length over TRC and its vertex (allsteps.py:467-502):
    returns (vertex_x, vertex_y, coeffs) with the reference's clamps (vertex > 1 -> median TRC,
    vertex < inputtrc -> inputtrc).  The PNG is written only if matplotlib is importable."""
    trc_arr = np.array(trc_list)
    telo_arr = np.array(telo_length_list)
    coeffs = np.polyfit(trc_arr, telo_arr, 2)
    a, b, c = coeffs
    vertex_x = -b / (2 * a)
    if vertex_x > 1.0:
        vertex_x = median_trc
    if vertex_x < inputtrc:
        vertex_x = inputtrc
    vertex_y = a * vertex_x ** 2 + b * vertex_x + c
    if save_path:
        try:
            import matplotlib
            matplotlib.use("Agg", force=False)
            import matplotlib.pyplot as plt
        except ImportError:
            plt = None
        if plt is not None:
            x_fit = np.linspace(min(trc_arr), max(trc_arr), 100)
            plt.figure(figsize=(7, 5))
            plt.scatter(trc_arr, telo_arr, color="blue", label="Topsicle results")
            plt.plot(x_fit, a * x_fit ** 2 + b * x_fit + c, color="red", label="Fit line")
            plt.scatter([vertex_x], [vertex_y], color="green", label="Vertex")
            plt.xlabel("TRC values")
            plt.ylabel("Telomere length, each read (bp)")
            plt.title("Quadratic fit plot")
            plt.legend()
            plt.tight_layout()
            plt.savefig(save_path, dpi=300)
            plt.close()
    return vertex_x, vertex_y, coeffs
