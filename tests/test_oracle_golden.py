"""CPU: pin the oracle (oracle/topsicle_oracle.py) to golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py) -- SURVEY 8c."""
import hashlib
import os
import random

import numpy as np
import pytest

from oracle import topsicle_oracle as orc
from tests.conftest import GOLD, load_json


def test_greedy_count_equals_re():
    rnd = random.Random(7)
    lits = ["AA", "CCC", "CTAAC", "GATTG", "ACA", "CCCTAA", "TAACCCT", "A", "ACAC"]
    for _ in range(300):
        n = rnd.randint(0, 120)
        alphabet = rnd.choice(["ACGT", "AC", "A", "ACGTN", "CTA"])
        s = "".join(rnd.choice(alphabet) for _ in range(n))
        for lit in lits:
            assert orc.greedy_count(s, lit) == orc.greedy_count_re(s, lit)


def test_patterns_table():
    for e in load_json("edge.json")["patterns"]:
        assert orc.patterns_to_search(e["motif"], e["k"]) == e["patterns"], e


def test_window_starts():
    for e in load_json("edge.json")["windows"]:
        starts = list(orc.window_starts(e["len"], e["W"], e["step"]))
        assert starts == e["starts"]
        assert all(l == e["W"] - 1 for l in e["lens"])


def _gain(c_w, b):
    from fractions import Fraction
    c = [int(v) for v in c_w]
    n, tot = len(c), sum(c)
    return Fraction((n * sum(c[:b]) - b * tot) ** 2, b * (n - b))


def _step1(records, motif, k, read_length, no_bp):
    rows = orc.pattern_trc_count(records, motif, read_length=read_length, kmer=k, no_bp=no_bp, cutoff=-1.0)
    return [[r[0], r[1], r[2], repr(float(r[3]))] for r in rows]


def test_step1_demo(demo_records):
    for e in load_json("demo_step1.json"):
        assert _step1(demo_records, e["pattern"], e["kmer"], e["read_length"], 1000) == e["rows"], e["pattern"]


def test_step1_edge(edge_records):
    fasta = list(orc.read_fastx(os.path.join(GOLD, "edge.fasta")))
    assert fasta == edge_records
    for e in load_json("edge.json")["step1"]:
        recs = edge_records if e["file"].endswith("fastq") else fasta
        assert _step1(recs, e["motif"], e["k"], e["read_length"], e["no_bp"]) == e["rows"], e


@pytest.mark.parametrize("exact", [False, True])
def test_step2_edge(edge_records, exact):
    seqs = dict(edge_records)
    n_checked = n_err = n_tie = 0
    for e in load_json("edge.json")["step2"]:
        pats = orc.patterns_to_search(e["motif"], e["k"])
        region = orc.oriented_region(seqs[e["read"]], e["tail"], e["trimfirst"], e["maxlengthtelo"])
        counts = orc.window_counts(region, pats, e["W"], e["slide"])
        assert counts.shape[0] == e["n_windows"]
        assert [int(v) for v in counts.sum(axis=1)] == e["c_w"]
        assert hashlib.md5(counts.astype(np.int16).tobytes()).hexdigest() == e["counts_md5"]
        if e["n_windows"] == 0:
            assert e.get("telo_length") is None
            continue
        c_w = counts.sum(axis=1)
        if "error" in e:
            with pytest.raises(ValueError):
                orc.change_point_float(c_w, len(pats))
            n_err += 1
            continue
        bkp = orc.change_point_exact(c_w) if exact else orc.change_point_float(c_w, len(pats))
        telo = e["trimfirst"] + e["slide"] * bkp
        if exact and telo != e["telo_length"]:
            # The exact-rational argmax may differ from the reference's float64 argmax only when
            # the rational gains tie exactly (then the reference's answer is summation noise).
            b_ref = (e["telo_length"] - e["trimfirst"]) // e["slide"]
            assert _gain(c_w, bkp) == _gain(c_w, b_ref), e
            n_tie += 1
            continue
        assert telo == e["telo_length"], e
        n_checked += 1
    assert n_checked > 400 and n_err > 0
    assert n_tie == (9 if exact else 0)   # constant signals only (all gains exactly 0)


def test_demo_rawcount_tables(demo_records):
    seqs = dict(demo_records)
    tabs = np.load(os.path.join(GOLD, "demo_rawcount.npz"))
    for m in load_json("demo_rawcount.json"):
        region = orc.oriented_region(seqs[m["read"]], m["tail"], m["trimfirst"], m["maxlengthtelo"])
        counts = orc.window_counts(region, m["patterns"], m["windowSize"], m["slide"])
        assert np.array_equal(counts, tabs[m["key"]].astype(np.int64)), m["key"]
        assert counts.shape[0] == m["n_windows"]
        for exact in (False, True):
            telo, _, _ = orc.bound_detect_read(seqs[m["read"]], m["tail"], m["patterns"], m["windowSize"],
                                               m["slide"], m["trimfirst"], m["maxlengthtelo"], exact=exact)
            assert telo == m["telo_length"]


def _cli_args(argv):
    import argparse
    p = argparse.ArgumentParser()
    p.add_argument("--pattern")
    p.add_argument("--minSeqLength", type=int, default=9000)
    p.add_argument("--telophrase", nargs="+", type=int)
    p.add_argument("--cutoff", nargs="+", type=float, default=0.7)
    p.add_argument("--windowSize", type=int, default=100)
    p.add_argument("--slide", type=int)
    p.add_argument("--trimfirst", type=int, default=100)
    p.add_argument("--maxlengthtelo", type=int, default=20000)
    p.add_argument("--rawcountpattern", action="store_true")
    return p.parse_args(argv)


@pytest.mark.parametrize("exact", [False, True])
def test_demo_cli_csv(demo_records, exact):
    """Whole path on the demo: CSV text byte-identical to the reference's (8 flag sets,
    the first one being the reference's own committed golden telolengths_all.csv)."""
    gold = load_json("demo_cli.json")
    assert gold["reference_golden"]["csv_md5"] == "92c042b7c7e13ae38ba5823370adc6a4"
    for case in gold["cases"]:
        a = _cli_args(case["argv"])
        phrases = a.telophrase or [len(a.pattern) - 2]
        cutoff = min(a.cutoff) if isinstance(a.cutoff, list) else a.cutoff
        slide = a.slide or len(a.pattern)
        body = []
        for k in phrases:
            rows = orc.scan_records(demo_records, a.pattern, k, cutoff, a.minSeqLength, a.windowSize, slide,
                                    a.trimfirst, a.maxlengthtelo, exact=exact)
            body.append(orc.csv_text("demo.fastq", k, rows))
        text = body[0] + "".join(b.split("\r\n", 1)[1] for b in body[1:])
        assert text == case["csv"], case["name"]
        assert hashlib.md5(text.encode()).hexdigest() == case["csv_md5"]


def test_heatmap_oracle_matches_reference_csv():
    """oracle.heatmap_matches / heatmap_csv_text == `patterns_vs_match_heatmap(...).to_csv(index=False)` of the
    unmodified reference, including its own golden Topsicle_demo/result_justone/heatmap_rawcount_1.csv."""
    import hashlib
    for case in load_json("demo_heatmap.json"):
        src = "demo.fastq.gz" if case["input"].endswith(".gz") else case["input"]
        recs = list(orc.read_fastx(os.path.join(GOLD, src)))
        if case["mode"] == "subset":       # overview_plot.py:63-84
            keep = {r[0] for r in orc.pattern_trc_count(recs, case["pattern"], read_length=case["minSeqLength"],
                                                        kmer=case["telophrase"], no_bp=1000, cutoff=0.7)}
            recs = [(i, s) for i, s in recs if i in keep]
        fwd, rev = orc.heatmap_matches(recs, case["pattern"], case["telophrase"], case["minSeqLength"])
        txt = orc.heatmap_csv_text(fwd, rev)
        assert len(fwd) + len(rev) == case["rows"], case
        assert hashlib.md5(txt.encode()).hexdigest() == case["md5"], case
