"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of Topsicle's per-read telomere scan.

This file is the *checker*: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import it.  The product
(`topsicle_b200/`) never imports anything under `oracle/`.

Every function cites the reference lines (`/root/reference/...`) it restates.
Parity status: PINNED -- `tests/test_oracle_golden.py` checks this restatement
against golden vectors produced by running the *unmodified* reference
(`oracle/make_golden.py`, reference imported through `oracle/shims/`), including
the reference's own golden `Topsicle_demo/telolengths_all.csv`.

Two flavours of the change-point step are provided:
  * `change_point_float`  -- what the reference computes: float64 `numpy.var`
    costs exactly as `ruptures==1.1.9` Binseg/CostL2 does (allsteps.py:310-311).
  * `change_point_exact`  -- the equivalent exact-rational argmax that the CUDA
    path implements: argmax_b (n*S_b - b*T)^2 / (b*(n-b)), ties -> larger b.
"""
from __future__ import annotations

import gzip
import re
from fractions import Fraction

import numpy as np

_COMP = str.maketrans("ACGT", "TGCA")


# --------------------------------------------------------------------------- patterns
def pattern_scramble_telo(pattern: str, cut_length) -> list[str]:
    """allsteps.py:57-82 -- unique cut_length-mers of pattern+pattern, sorted."""
    doubled = (pattern + pattern).upper()
    lengths = cut_length if isinstance(cut_length, list) else [cut_length]
    cuts = set()
    for k in lengths:
        for i in range(len(doubled) - k + 1):
            cuts.add(doubled[i:i + k])
    return sorted(cuts)


def patterns_to_search(telopattern, cut_length) -> list[str]:
    """allsteps.py:84-125 -- origin k-mers followed by their complements (not reversed).

    A list argument is returned upper-cased unchanged (allsteps.py:122-123).
    `|` patterns yield a broken string in the reference; rejected here.
    """
    if isinstance(telopattern, list):
        return [p.upper() for p in telopattern]
    if "|" in telopattern:
        raise ValueError("'|' patterns are undefined behaviour in the reference")
    origin = pattern_scramble_telo(telopattern, [cut_length])
    return [p.upper() for p in origin + [p.translate(_COMP) for p in origin]]


# --------------------------------------------------------------------------- matching
def greedy_count(text: str, literal: str) -> int:
    """len(list(re.compile(literal).finditer(text))) -- allsteps.py:182-183, 281, 288.

    Leftmost non-overlapping matches; `str.count` has the identical definition for a
    literal (checked against `re` in tests/test_oracle_golden.py).
    """
    if not literal:
        return len(text) + 1
    return text.count(literal)


def greedy_count_re(text: str, literal: str) -> int:
    """Same value through `re`, the way the reference does it."""
    return sum(1 for _ in re.compile(literal).finditer(text))


# --------------------------------------------------------------------------- step 1
def trc_read(seq: str, patterns: list[str], len_telopattern: int, no_bp: int = 1000):
    """allsteps.py:176-198 for one read -> (tail, best_idx, count, ms, me, head_cnt, tail_cnt).

    `count / (no_bp / len_telopattern)` is the reference's TRC value; the forward /
    reverse decision compares the maxima strictly (tie -> reverse) and `max()` keeps
    the FIRST maximum in pattern order.
    """
    head = seq[:no_bp].upper()
    tail = seq[-no_bp:][::-1].upper()
    hc = [greedy_count(head, p) for p in patterns]
    tc = [greedy_count(tail, p) for p in patterns]
    ms, me = max(hc), max(tc)
    if ms > me:
        return "forward", hc.index(ms), ms, ms, me, hc, tc
    return "reverse", tc.index(me), me, ms, me, hc, tc


def trc_value(count: int, len_telopattern: int, no_bp: int = 1000) -> float:
    """allsteps.py:178,185-186 -- float64 `matches / (no_bp / len(telopattern))`."""
    return count / (no_bp / len_telopattern)


def pattern_trc_count(records, telopattern, read_length=0, kmer=4, no_bp=1000, cutoff=0.5):
    """allsteps.py:152-204 on an iterable of (id, seq) -> [[id, literal, tail, trc], ...]."""
    patterns = patterns_to_search(telopattern, kmer)
    out = []
    for rid, seq in records:
        if len(seq) > read_length:
            tail, bi, cnt, *_ = trc_read(seq, patterns, len(telopattern), no_bp)
            trc = trc_value(cnt, len(telopattern), no_bp)
            if trc > cutoff:
                out.append([rid, patterns[bi], tail, trc])
    return out


# --------------------------------------------------------------------------- step 2 / 3
def oriented_region(seq: str, tail: str, trimfirst: int, maxlengthtelo: int) -> str:
    """allsteps.py:263-271 -- `[trimfirst:min(maxlengthtelo, len)]` of the read or its reversal."""
    m = min(maxlengthtelo, len(seq))
    if tail == "forward":
        return seq[trimfirst:m].upper()
    return seq[::-1].upper()[trimfirst:m]


def window_starts(region_len: int, window_size: int, step: int) -> range:
    """allsteps.py:219 -- `range(0, len(s) - window_size + 1, step)`."""
    return range(0, region_len - window_size + 1, step)


def window_counts(region: str, patterns: list[str], window_size: int, step: int) -> np.ndarray:
    """allsteps.py:219-224 + 279-283 -- counts[w][p] = greedy count in `z[i:i+W-1]` floored at 1."""
    starts = window_starts(len(region), window_size, step)
    out = np.empty((len(starts), len(patterns)), dtype=np.int64)
    for w, i in enumerate(starts):
        text = region[i:i + window_size - 1]
        for p, lit in enumerate(patterns):
            out[w, p] = greedy_count(text, lit) or 1
    return out


def change_point_float(c_w: np.ndarray, n_patterns: int, jump: int = 5, min_size: int = 2) -> int:
    """allsteps.py:283 (mean) + 310-311 -> ruptures 1.1.9 Binseg(l2).predict(n_bkps=1)[0].

    y = sum/len as Python floats (int / int), costs via numpy float64 var exactly as
    CostL2.error does on an (n, 1) array.  Raises ValueError for n < 7 the way
    ruptures raises BadSegmentationParameters.
    """
    y = np.array([int(c) / n_patterns for c in c_w]).reshape(-1, 1)
    n = y.shape[0]
    if n // jump < 1 or ((min_size + jump - 1) // jump) * jump + min_size > n:
        raise ValueError("BadSegmentationParameters")

    def cost(a, b):
        return y[a:b].var(axis=0).sum() * (b - a)

    total = cost(0, n)
    best = None
    for b in range(0, n, jump):
        if b >= min_size and n - b >= min_size:
            cand = (total - cost(0, b) - cost(b, n), b)
            if best is None or cand > best:
                best = cand
    return best[1]


def change_point_exact(c_w, jump: int = 5, min_size: int = 2) -> int:
    """Exact form of the same argmax: gain(b) ~ (n*S_b - b*T)^2 / (b*(n-b)); ties -> larger b."""
    c = [int(v) for v in c_w]
    n = len(c)
    if n // jump < 1 or ((min_size + jump - 1) // jump) * jump + min_size > n:
        raise ValueError("BadSegmentationParameters")
    total = sum(c)
    prefix = 0
    pos = 0
    best = None
    for b in range(0, n, jump):
        while pos < b:
            prefix += c[pos]
            pos += 1
        if b >= min_size and n - b >= min_size:
            num = (n * prefix - b * total) ** 2
            cand = (Fraction(num, b * (n - b)), b)
            if best is None or cand >= best:
                best = cand
    return best[1]


def bound_detect_read(seq: str, tail: str, patterns: list[str], window_size: int, slide: int,
                      trimfirst: int, maxlengthtelo: int, exact: bool = False):
    """allsteps.py:263-315 for one read -> (telo_length, c_w, counts).

    telo_length = x[bkp] = trimfirst + slide * bkp  (allsteps.py:304, 312-315).
    """
    region = oriented_region(seq, tail, trimfirst, maxlengthtelo)
    counts = window_counts(region, patterns, window_size, slide)
    c_w = counts.sum(axis=1)
    bkp = change_point_exact(c_w) if exact else change_point_float(c_w, len(patterns))
    return trimfirst + slide * bkp, c_w, counts


# --------------------------------------------------------------------------- whole path
def scan_records(records, pattern: str, telophrase: int, cutoff: float, min_seq_length: int,
                 window_size: int, slide: int, trimfirst: int, maxlengthtelo: int,
                 exact: bool = False, want_counts: bool = False):
    """process_file (main.py:52-154) minus file I/O: rows in file order.

    Returns a list of dicts {id, trc, tail, best, count, telo_length[, counts]} for the
    reads that pass step 1 (L > minSeqLength and trc > cutoff).
    """
    patterns = patterns_to_search(pattern, telophrase)
    rows = []
    for rid, seq in records:
        if len(seq) <= min_seq_length:
            continue
        tail, bi, cnt, ms, me, _, _ = trc_read(seq, patterns, len(pattern))
        trc = trc_value(cnt, len(pattern))
        if not trc > cutoff:
            continue
        telo, c_w, counts = bound_detect_read(seq, tail, patterns, window_size, slide,
                                              trimfirst, maxlengthtelo, exact=exact)
        row = dict(id=rid, trc=trc, tail=tail, best=bi, count=cnt, ms=ms, me=me,
                   telo_length=telo, c_w=c_w)
        if want_counts:
            row["counts"] = counts
        rows.append(row)
    return rows


def csv_text(file_stem: str, phrase: int, rows) -> str:
    """main.py:198-200 header + :138 rows, csv.writer default dialect (CRLF)."""
    lines = ["file_number,phrase,trc,readID,telo_length"]
    for r in rows:
        lines.append(f"{file_stem},{phrase},{r['trc']:.3f},{r['id']},{r['telo_length']}")
    return "\r\n".join(lines) + "\r\n"


# --------------------------------------------------------------------------- input
# ------------------------------------------------------------------ overview heat map (descriptive_plot.py)
def heatmap_matches(records, telopattern: str, telophrase: int, min_seq_length: int):
    """descriptive_plot.py:233-313 `patterns_vs_match_heatmap`, the data half: for every read longer than
    min_seq_length, every origin k-mer (NOT its complement) followed by `len(telopattern) - telophrase` more
    characters, leftmost non-overlapping (`re.finditer(pattern(.{finding}))`), in `seq[100:2000]` and in the
    complement of `reversed(seq)[100:2000]`.  Returns (forward_rows, reverse_rows) of (pattern, match, name)
    in the order the reference appends them: read-major, pattern-minor, by position."""
    pats = pattern_scramble_telo(telopattern, telophrase)
    finding = int(len(telopattern) - telophrase)
    trans = str.maketrans("ACGT", "TGCA")
    fwd, rev = [], []
    for name, seq in records:
        if not len(seq) > min_seq_length:
            continue
        s1 = seq[100:2000].upper()
        s2 = seq[::-1][100:2000].upper().translate(trans)
        for pat in pats:
            rx = re.compile(rf"{re.escape(pat)}(.{{{finding}}})")
            fwd += [(pat, m.group(1), name) for m in rx.finditer(s1)]
            rev += [(pat, m.group(1), name) for m in rx.finditer(s2)]
    return fwd, rev


def heatmap_csv_text(fwd, rev) -> str:
    """`allstrands.to_csv(index=False)` of the reference (overview_plot.py:104-108): forward rows, then reverse
    rows; the read id column holds a one-element list."""
    out = ["Pattern,Match,read id\n"]
    for pat, match, name in list(fwd) + list(rev):
        out.append(f"{pat},{match},['{name}']\n")
    return "".join(out)


def read_fastx(path: str):
    """allsteps.py:36-50,127-149 -- (id, seq) records; id = title up to first whitespace.

    4-line FASTQ and multi-line FASTA, optionally gzip (by `.gz` suffix).
    """
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt", encoding="utf-8") as fh:
        first = fh.readline()
        if first.startswith("@"):
            title = first
            while title:
                seq = fh.readline().rstrip()
                fh.readline()
                fh.readline()
                t = title[1:].rstrip()
                yield (t.split(None, 1)[0] if t.split() else ""), seq
                title = fh.readline()
                while title in ("\n", "\r\n"):
                    title = fh.readline()
        elif first.startswith(">"):
            title, chunks = first[1:].rstrip(), []
            for line in fh:
                if line.startswith(">"):
                    yield (title.split(None, 1)[0] if title.split() else ""), "".join(chunks)
                    title, chunks = line[1:].rstrip(), []
                else:
                    chunks.append(line.strip())
            yield (title.split(None, 1)[0] if title.split() else ""), "".join(chunks)
        else:
            raise ValueError("Format cannot be identified")
