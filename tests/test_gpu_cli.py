"""GPU tier: the drop-in CLI and the allsteps-compatible API on the real kernels, against the outputs
of the unmodified reference (tests/golden, made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import topsicle_oracle as orc
from tests.cli_cases import golden_cases, run_case
from tests.conftest import GOLD, load_json

pytestmark = pytest.mark.gpu
DEMO = os.path.join(GOLD, "demo.fastq.gz")


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_cli_matches_reference_outputs(case, tmp_path):
    run_case(case, tmp_path)


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_cli_ends_first_matches_reference_outputs(case, tmp_path):
    """--ends-first on the real kernels: tps_submit_ends + tps_submit_regions, same bytes out."""
    run_case(case, tmp_path, extra_argv=["--ends-first"])


def test_cli_ends_first_small_batches(tmp_path, monkeypatch):
    monkeypatch.setenv("TOPSICLE_BATCH_READS", "5")
    monkeypatch.setenv("TOPSICLE_ENDS_FIRST", "1")
    case = [c for c in golden_cases() if c["name"] == "CCCTAA_sweep_456_cut04_07"][0]
    run_case(case, tmp_path)


def test_cli_small_batches(tmp_path, monkeypatch):
    monkeypatch.setenv("TOPSICLE_BATCH_READS", "5")
    case = [c for c in golden_cases() if c["name"] == "CCCTAA_sweep_456_cut04_07"][0]
    run_case(case, tmp_path)


def test_patternTRC_count_golden():
    """allsteps.patternTRC_count rows == the unmodified reference's, including the float TRC."""
    from topsicle_b200.allsteps import patternTRC_count
    for e in load_json("demo_step1.json"):
        rows = patternTRC_count(DEMO, e["pattern"], read_length=e["read_length"], kmer=e["kmer"], no_bp=1000,
                                cutoff=-1.0)
        assert [[r[0], r[1], r[2], repr(float(r[3]))] for r in rows] == e["rows"], (e["pattern"], e["kmer"])


def test_bound_detect_and_rawcount_golden():
    from topsicle_b200.allsteps import bound_detect, rawCountPattern
    tabs = np.load(os.path.join(GOLD, "demo_rawcount.npz"))
    for m in load_json("demo_rawcount.json"):
        bd = bound_detect(DEMO, m["read"], m["patterns"], m["windowSize"], m["slide"], m["trimfirst"],
                          m["maxlengthtelo"], m["kmer"], tail=m["tail"])
        assert bd == [[m["read"], m["telo_length"]]]
        df = rawCountPattern(DEMO, m["read"], m["patterns"], m["windowSize"], m["slide"], m["trimfirst"], m["kmer"],
                             9000, m["maxlengthtelo"], tail=m["tail"])
        n_pat = len(m["patterns"])
        assert list(df.columns) == ["tail", "position", "pattern", "count"]
        assert np.array_equal(df["count"].to_numpy().reshape(-1, n_pat), tabs[m["key"]].astype(np.int64))
        assert list(df["pattern"][:n_pat]) == m["patterns"] and set(df["tail"]) == {m["tail"]}
        pos = df["position"].to_numpy().reshape(-1, n_pat)[:, 0]
        assert (int(pos[0]), int(pos[-1]), len(pos)) == (m["first_pos"], m["last_pos"], m["n_windows"])


def test_bound_detect_both_tails_and_errors(tmp_path):
    from topsicle_b200.allsteps import bound_detect, rawCountPattern, patterns_to_search
    recs = dict(orc.read_fastx(DEMO))
    rid = "ERR11436636.60645"
    pats = patterns_to_search("CCCTAAA", 5)
    both = bound_detect(DEMO, rid, pats, 100, 6, 100, 20000, 5)          # tail=None: reverse first, then forward
    want = []
    for tail in ("reverse", "forward"):
        t, _, _ = orc.bound_detect_read(recs[rid], tail, pats, 100, 6, 100, 20000, exact=True)
        want.append([rid, t])
    assert both == want
    df = rawCountPattern(DEMO, rid, pats, 100, 6, 100, 5, 9000, 20000)
    assert list(df["tail"].unique()) == ["forward", "reverse"]
    assert bound_detect(DEMO, "no_such_read", pats, 100, 6, 100, 20000, 5, tail="forward") == []
    assert bound_detect(DEMO, 7, pats, 100, 6, 100, 20000, 5) is None
    short = tmp_path / "short.fasta"
    short.write_text(">s1\n" + "CCCTAAA" * 17 + "\n")          # 119 bases -> 4 windows
    with pytest.raises(ValueError):
        bound_detect(str(short), "s1", pats, 100, 6, 0, 20000, 5, tail="forward")   # 7 windows needed


def test_unzip_file_and_check_file_type():
    from topsicle_b200.allsteps import check_file_type, unzip_file
    assert check_file_type(DEMO) == "fastq"
    assert check_file_type(os.path.join(GOLD, "edge.fasta")) == "fasta"
    got = [(r.id, str(r.seq)) for r in unzip_file(DEMO)]
    assert got == list(orc.read_fastx(DEMO))
