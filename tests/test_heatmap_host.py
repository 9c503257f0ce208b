"""CPU tier: the host side of the overview heat map (row order, CSV text, DataFrame, `overview_plot` CLI flow) with
the GPU matching answered by the regex oracle (TEST stand-in for `engine.follow_scan`).  The real-kernel version
is tests/test_gpu_heatmap.py."""
import hashlib
import os
import re
import shutil

import numpy as np
import pytest

from oracle import topsicle_oracle as orc
from tests import fake_engine
from tests.conftest import GOLD, load_json


def _oracle_follow_scan(seqs, kmers, match_len, min_seq_length, skip=100, upto=2000, device=0):
    """What tps_follow_scan returns, computed with `re` (descriptive_plot.py:262-289)."""
    seqs = [s.decode("ascii", "replace") if isinstance(s, (bytes, bytearray)) else s for s in seqs]
    k = len(kmers[0])
    out = np.zeros((len(seqs), 2, len(kmers), upto - skip), dtype=bool)
    trans = str.maketrans("ACGT", "TGCA")
    for r, seq in enumerate(seqs):
        if not len(seq) > min_seq_length:
            continue
        texts = (seq[skip:upto].upper(), seq[::-1][skip:upto].upper().translate(trans))
        for p, kmer in enumerate(kmers):
            rx = re.compile(rf"{re.escape(kmer)}(.{{{match_len - k}}})")
            for strand, text in enumerate(texts):
                for m in rx.finditer(text):
                    out[r, strand, p, m.start()] = True
    return out


@pytest.fixture
def fake_gpu(monkeypatch):
    from topsicle_b200 import engine
    fake_engine.install(monkeypatch)
    monkeypatch.setattr(engine, "follow_scan", _oracle_follow_scan)


@pytest.mark.parametrize("case", load_json("demo_heatmap.json"), ids=lambda c: f"{c['pattern']}_{c['telophrase']}_{c['input'][:4]}")
def test_heatmap_rows_and_csv_text(case, fake_gpu):
    from topsicle_b200 import descriptive
    src = "demo.fastq.gz" if case["input"].endswith(".gz") else case["input"]
    recs = list(orc.read_fastx(os.path.join(GOLD, src)))
    if case["mode"] == "subset":
        keep = {r[0] for r in orc.pattern_trc_count(recs, case["pattern"], read_length=case["minSeqLength"],
                                                    kmer=case["telophrase"], no_bp=1000, cutoff=0.7)}
        recs = [(i, s) for i, s in recs if i in keep]
    fwd, rev = descriptive.heatmap_rows(recs, case["pattern"], case["telophrase"], case["minSeqLength"])
    txt = descriptive.heatmap_csv_text(fwd, rev)
    assert len(fwd) + len(rev) == case["rows"]
    assert hashlib.md5(txt.encode()).hexdigest() == case["md5"]


def test_heatmap_dataframe_and_overview_cli(tmp_path, fake_gpu):
    from topsicle_b200 import descriptive, overview_plot
    cases = load_json("demo_heatmap.json")
    df = descriptive.patterns_vs_match_heatmap(os.path.join(GOLD, "demo.fastq.gz"), cases[1]["pattern"],
                                               cases[1]["telophrase"], cases[1]["minSeqLength"])
    assert list(df.columns) == ["Pattern", "Match", "read id"]
    assert hashlib.md5(df.to_csv(index=False).encode()).hexdigest() == cases[1]["md5"]
    assert list(df["Match"].cat.categories) == sorted(df["Match"].unique())
    indir, out = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    shutil.copy(os.path.join(GOLD, "demo.fastq.gz"), indir / "demo.fastq.gz")
    overview_plot.main(["--inputDir", str(indir), "--outputDir", str(out), "--pattern", "CCCTAAA",
                        "--recfindingpattern", "--rawcount"])
    got = hashlib.md5(open(out / "heatmap_rawcount_1.csv", "rb").read()).hexdigest()
    assert got == cases[0]["md5"] == "28ad064f247aa236af6f0fedddc63ed4"   # the reference's own golden file
    # a bad path: unzip_file logs and yields nothing (the reference's `== None` test is dead code as well)
    assert len(descriptive.patterns_vs_match_heatmap(["a", "b"], "CCCTAA", 4, 0)) == 0
