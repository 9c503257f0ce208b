"""The data half of the reference's overview heat map (Topsicle/descriptive_plot.py:233-313
`patterns_vs_match_heatmap`; `overview_plot.py --recfindingpattern --rawcount` writes it to
heatmap_rawcount_{i}.csv): which characters follow each telomere k-mer in the first and last 2 kb of the reads.

The matching (every origin k-mer followed by `len(pattern) - telophrase` characters, leftmost non-overlapping, on
`seq[100:2000]` and on the complement of `reversed(seq)[100:2000]`) runs on the GPU (`tps_follow_scan`, kernel K5);
the host only turns the match positions into the reference's rows.  The plots themselves (seaborn heat map,
`descriptive_plot`) are not rebuilt: SURVEY 2 rows 10-11 are visual only.
"""
from __future__ import annotations

import numpy as np

from . import engine
from .patterns import pattern_scramble_telo

_COMPLEMENT = bytes.maketrans(b"ACGT", b"TGCA")
SKIP, UPTO = 100, 2000        # descriptive_plot.py:266,268


def heatmap_rows(records, telopattern: str, telophrase: int, minSeqLength: int, device: int = 0):
    """(forward_rows, reverse_rows) of (pattern, match, read_name), in the reference's order (read-major,
    pattern-minor, by position).  `records` = iterable of (name, sequence)."""
    recs = [(n, s.encode("ascii", "replace") if isinstance(s, str) else bytes(s)) for n, s in records]
    kmers = pattern_scramble_telo(telopattern, telophrase)
    k = int(telophrase)
    match_len = len(telopattern)
    if match_len < k:
        raise ValueError("telophrase longer than the pattern")
    sel = engine.follow_scan([s for _, s in recs], kmers, match_len, minSeqLength, SKIP, UPTO, device=device)
    fwd, rev = [], []
    for r, (name, seq) in enumerate(recs):
        if not len(seq) > minSeqLength:
            continue
        s1 = seq[SKIP:UPTO].upper()
        s2 = seq[::-1][SKIP:UPTO].upper().translate(_COMPLEMENT)
        for p, kmer in enumerate(kmers):
            for text, bits, out in ((s1, sel[r, 0, p], fwd), (s2, sel[r, 1, p], rev)):
                for pos in np.nonzero(bits)[0].tolist():
                    out.append((kmer, text[pos + k:pos + match_len].decode("ascii", "replace"), name))
    return fwd, rev


def heatmap_csv_text(fwd, rev) -> str:
    """`allstrands.to_csv(index=False)` as overview_plot.py:104-108 writes it."""
    out = ["Pattern,Match,read id\n"]
    out += [f"{p},{m},['{n}']\n" for p, m, n in list(fwd) + list(rev)]
    return "".join(out)


def patterns_vs_match_heatmap(filepath, telopattern, telophrase, minSeqLength, device: int = 0):
    """Drop-in for descriptive_plot.py:233-313 without the figure: returns the `allstrands` DataFrame
    (Pattern, Match as an ordered categorical, read id as a one-element list), forward rows then reverse rows."""
    import pandas as pd
    from .allsteps import unzip_file
    gen = unzip_file(filepath)
    if gen is None:
        print("problem in filepath, can not have heatmap")
        return None
    kmers = pattern_scramble_telo(telopattern, telophrase)
    print(kmers)
    fwd, rev = heatmap_rows(((r.name, str(r.seq)) for r in gen), telopattern, telophrase, minSeqLength, device)
    cols = ["Pattern", "Match", "read id"]
    allstrands = pd.concat([pd.DataFrame([(p, m, [n]) for p, m, n in fwd], columns=cols),
                            pd.DataFrame([(p, m, [n]) for p, m, n in rev], columns=cols)], ignore_index=True)
    order = sorted(allstrands["Match"].dropna().unique())
    allstrands["Match"] = pd.Categorical(allstrands["Match"], categories=order, ordered=True)
    return allstrands
