"""GPU tier: the overview heat-map data path (K5 `tps_follow_kernel` through `tps_follow_scan`) against the CSV
the unmodified reference writes (`patterns_vs_match_heatmap(...).to_csv(index=False)`, descriptive_plot.py:233-313),
including the reference's own golden Topsicle_demo/result_justone/heatmap_rawcount_1.csv."""
import hashlib
import os

import numpy as np
import pytest

from oracle import topsicle_oracle as orc
from tests.conftest import GOLD, load_json

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", load_json("demo_heatmap.json"), ids=lambda c: f"{c['pattern']}_{c['telophrase']}_{c['input'][:4]}")
def test_heatmap_csv_equals_reference(case):
    from topsicle_b200 import descriptive
    src = "demo.fastq.gz" if case["input"].endswith(".gz") else case["input"]
    recs = list(orc.read_fastx(os.path.join(GOLD, src)))
    if case["mode"] == "subset":        # overview_plot.py:63-84: the reads with TRC > 0.7 first
        from topsicle_b200.allsteps import patternTRC_count
        keep = {r[0] for r in patternTRC_count(os.path.join(GOLD, src), case["pattern"], read_length=case["minSeqLength"],
                                               kmer=case["telophrase"], no_bp=1000, cutoff=0.7)}
        recs = [(i, s) for i, s in recs if i in keep]
    fwd, rev = descriptive.heatmap_rows(recs, case["pattern"], case["telophrase"], case["minSeqLength"])
    txt = descriptive.heatmap_csv_text(fwd, rev)
    assert len(fwd) + len(rev) == case["rows"]
    assert hashlib.md5(txt.encode()).hexdigest() == case["md5"]


def test_heatmap_dataframe_drop_in():
    from topsicle_b200 import descriptive
    case = load_json("demo_heatmap.json")[1]
    df = descriptive.patterns_vs_match_heatmap(os.path.join(GOLD, "demo.fastq.gz"), case["pattern"], case["telophrase"],
                                               case["minSeqLength"])
    assert list(df.columns) == ["Pattern", "Match", "read id"] and len(df) == case["rows"]
    assert hashlib.md5(df.to_csv(index=False).encode()).hexdigest() == case["md5"]
    assert list(df["Match"].cat.categories) == sorted(df["Match"].unique())


@pytest.mark.parametrize("motif,k", [("CCCTAA", 4), ("CCCTAA", 6), ("CCCTAAA", 5), ("AAACCCT", 3), ("TTAGGG", 1),
                                     ("AACCGGTT", 8), ("CCCTAA", 2)])
def test_heatmap_random_reads_equal_oracle(motif, k):
    """Random reads around every boundary (shorter than 100, between 100 and 2000, telomeric, N, lower case):
    match positions and following characters equal the regex restatement."""
    from topsicle_b200 import descriptive
    rng = np.random.default_rng(len(motif) * 10 + k)
    B = np.array(list("ACGT"))
    recs = []
    for i in range(160):
        L = int(rng.choice([rng.integers(0, 120), rng.integers(100, 2100), rng.integers(2000, 9000)]))
        s = B[rng.integers(0, 4, L)]
        if i % 3 == 0 and L > 50:
            tl = int(rng.integers(20, L))
            rep = np.array(list((motif * (tl // len(motif) + 2))[int(rng.integers(0, len(motif))):][:tl]))
            err = rng.random(tl) < 0.05
            rep[err] = B[rng.integers(0, 4, int(err.sum()))]
            if i % 2:
                s[:tl] = rep
            else:
                s[L - tl:] = np.array(list("".join(rep)[::-1].translate(str.maketrans("ACGT", "TGCA"))))
        if i % 5 == 0:
            s[rng.random(L) < 0.03] = "N"
        seq = "".join(s)
        if i % 7 == 0:
            seq = seq[:L // 3] + seq[L // 3:].lower()
        recs.append((f"r{i}", seq))
    for minlen in (0, 150):
        want = orc.heatmap_matches(recs, motif, k, minlen)
        got = descriptive.heatmap_rows(recs, motif, k, minlen)
        assert got[0] == want[0] and got[1] == want[1], (motif, k, minlen)
        assert len(want[0]) + len(want[1]) > 100


def test_overview_plot_cli_writes_the_reference_csv(tmp_path):
    """`overview_plot --recfindingpattern --rawcount` on the demo == the reference's own
    Topsicle_demo/result_justone/heatmap_rawcount_1.csv (md5 28ad064f...)."""
    import shutil
    from topsicle_b200 import overview_plot
    indir, out = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    shutil.copy(os.path.join(GOLD, "demo.fastq.gz"), indir / "demo.fastq.gz")
    overview_plot.main(["--inputDir", str(indir), "--outputDir", str(out), "--pattern", "CCCTAAA",
                        "--recfindingpattern", "--rawcount"])
    got = hashlib.md5(open(out / "heatmap_rawcount_1.csv", "rb").read()).hexdigest()
    assert got == load_json("demo_heatmap.json")[0]["md5"] == "28ad064f247aa236af6f0fedddc63ed4"


def test_heatmap_empty_inputs():
    from topsicle_b200 import descriptive, engine
    assert descriptive.heatmap_rows([], "CCCTAA", 4, 0) == ([], [])
    assert descriptive.heatmap_rows([("a", ""), ("b", "")], "CCCTAA", 4, 0) == ([], [])
    sel = engine.follow_scan(["", "ACGT"], ["CCCT"], 6, 0)
    assert sel.shape == (2, 2, 1, 1900) and not sel.any()
