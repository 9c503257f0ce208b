"""Telophrase expansion: the list of literals the scan searches for.

Mirrors `pattern_scramble_telo` / `patterns_to_search` of the reference
(Topsicle/allsteps.py:57-82, 84-125): every distinct `cut_length`-long window of the
doubled motif in sorted order, followed by the base-wise complement (A<->T, C<->G, NOT
reversed) of each.  Order matters: step 1 keeps the first maximum (allsteps.py:190-191)
and the rawcount table's columns follow it (allsteps.py:401-411).
"""
from __future__ import annotations

_COMPLEMENT = {"A": "T", "C": "G", "G": "C", "T": "A"}


def pattern_scramble_telo(pattern: str, cut_length) -> list[str]:
    motif2 = (pattern * 2).upper()
    lengths = cut_length if isinstance(cut_length, list) else [cut_length]
    seen = {motif2[start:start + k] for k in lengths for start in range(len(motif2) - k + 1)}
    return sorted(seen)


def _complement(kmer: str) -> str:
    return "".join(_COMPLEMENT.get(ch, ch) for ch in kmer)


def patterns_to_search(telopattern, cut_length) -> list[str]:
    if isinstance(telopattern, list):
        return [lit.upper() for lit in telopattern]          # allsteps.py:122-123
    if "|" in telopattern:
        # The reference builds a single string here that its callers then iterate
        # character by character (allsteps.py:90-102,168): undefined behaviour, rejected.
        raise ValueError("'|' alternation patterns are not supported (broken in the reference)")
    kmers = pattern_scramble_telo(telopattern, [cut_length])
    return [lit.upper() for lit in kmers + [_complement(k) for k in kmers]]


def validate_literals(literals) -> None:
    """Only ACGT literals have defined behaviour (regex metacharacters are UB upstream)."""
    for lit in literals:
        if not lit or any(ch not in "ACGT" for ch in lit.upper()):
            raise ValueError(f"pattern literal {lit!r} must be a non-empty ACGT string")
