"""CPU tier: the host side of the drop-in (reader -> batch pipeline -> CLI writers) with the GPU
answered by the oracle stand-in (tests/fake_engine.py).  The real-kernel version of the same
checks is tests/test_gpu_cli.py."""
import os

import numpy as np
import pytest

from tests import fake_engine
from tests.cli_cases import golden_cases, run_case
from tests.conftest import GOLD


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_cli_matches_reference_outputs(case, tmp_path, monkeypatch):
    fake_engine.install(monkeypatch)
    run_case(case, tmp_path)


def test_cli_small_batches_and_two_devices(tmp_path, monkeypatch):
    """Tiny batches dealt over two (stand-in) devices: same bytes out, file order restored."""
    fake_engine.install(monkeypatch)
    monkeypatch.setenv("TOPSICLE_BATCH_READS", "3")
    case = [c for c in golden_cases() if c["name"] == "CCCTAA_sweep_456_cut04_07"][0]
    run_case(case, tmp_path, extra_argv=["--devices", "0", "1"])


def test_cli_refuses_to_overwrite(tmp_path, monkeypatch):
    fake_engine.install(monkeypatch)
    case = golden_cases()[0]
    run_case(case, tmp_path)
    from topsicle_b200 import main as tmain
    argv = ["--inputDir", str(tmp_path / "in"), "--outputDir", str(tmp_path / "out"), "--pattern", "CCCTAAA"]
    with pytest.raises(SystemExit) as e:
        tmain.main(argv)
    assert e.value.code == 1
    tmain.main(argv + ["--override", "--slide", "6"])
    assert open(tmp_path / "out" / "telolengths_all.csv", newline="").read() == case["csv"]


def test_cli_phrase_longer_than_pattern_exits(tmp_path, monkeypatch):
    fake_engine.install(monkeypatch)
    from topsicle_b200 import main as tmain
    with pytest.raises(SystemExit):
        tmain.main(["--inputDir", os.path.join(GOLD, "demo.fastq.gz"), "--outputDir", str(tmp_path / "o"),
                    "--pattern", "CCCTAA", "--telophrase", "7"])


def test_pass_capacity_overflow_is_split_not_lost(tmp_path, monkeypatch):
    """More TRC-pass reads in a batch than the context can hold -> the batch is re-scanned in halves."""
    fake_engine.install(monkeypatch)
    from topsicle_b200 import pipeline
    from topsicle_b200.patterns import patterns_to_search
    cfg = pipeline.ScanConfig(patterns=patterns_to_search("CCCTAAA", 5), len_telopattern=7, phrase=5, slide=6)
    path = os.path.join(GOLD, "demo.fastq.gz")
    _, big = pipeline.collect_file(path, [cfg], max_pass_reads=1 << 10)
    _, small = pipeline.collect_file(path, [cfg], max_pass_reads=2, max_batch_reads=1 << 10)
    assert len(big[0]) == 17
    assert [(p.index, p.read_id, p.telo_length) for p in small[0]] == [(p.index, p.read_id, p.telo_length)
                                                                      for p in big[0]]


@pytest.mark.parametrize("ends_first", ["0", "1"])
def test_read_check_only_that_read(tmp_path, monkeypatch, ends_first):
    """--read_check: one CSV row, but the subset file still holds every TRC-pass read (main.py:64-87
    runs before the read_check branch)."""
    fake_engine.install(monkeypatch)
    monkeypatch.setenv("TOPSICLE_ENDS_FIRST", ends_first)
    import hashlib
    from topsicle_b200 import main as tmain
    case = golden_cases()[0]
    out = tmp_path / "o"
    if hasattr(tmain.tprint, "logfile"):
        del tmain.tprint.logfile
    tmain.main(["--inputDir", os.path.join(GOLD, "demo.fastq.gz"), "--outputDir", str(out), "--pattern", "CCCTAAA",
                "--slide", "6", "--read_check", "ERR11436636.163616"])
    rows = open(out / "telolengths_all.csv", newline="").read().split("\r\n")
    assert rows[1:] == ["demo.fastq,5,0.791,ERR11436636.163616,3010", ""]
    sub = hashlib.md5(open(out / "demo.fastq_trc_over_0.7.fastq", "rb").read()).hexdigest()
    assert sub == case["files"]["demo.fastq_trc_over_0.7.fastq"]


@pytest.mark.parametrize("ends_first", ["0", "1"])
def test_directory_of_files_scanned_concurrently(tmp_path, monkeypatch, ends_first):
    """--inputDir with several files (.fastq.gz, .fastq, .fasta): files are read concurrently, every
    file's rows stay in file order, subset files are per file with the reference's naming rule."""
    fake_engine.install(monkeypatch)
    monkeypatch.setenv("TOPSICLE_ENDS_FIRST", ends_first)
    import gzip
    import shutil
    from oracle import topsicle_oracle as orc
    from topsicle_b200 import main as tmain
    recs = list(orc.read_fastx(os.path.join(GOLD, "demo.fastq.gz")))
    indir, out = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    shutil.copy(os.path.join(GOLD, "demo.fastq.gz"), indir / "a.fastq.gz")
    with open(indir / "b.fq", "w") as fh:
        for rid, s in recs[10:40]:
            fh.write(f"@{rid} x\n{s}\n+\n{'I' * len(s)}\n")
    with open(indir / "c.fasta", "w") as fh:
        for rid, s in recs[:25]:
            fh.write(f">{rid}\n" + "\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + "\n")
    if hasattr(tmain.tprint, "logfile"):
        del tmain.tprint.logfile
    monkeypatch.setenv("TOPSICLE_BATCH_READS", "7")
    tmain.main(["-i", str(indir), "-o", str(out), "--pattern", "CCCTAAA", "--slide", "6", "--devices", "0", "1"])
    rows = [ln.split(",") for ln in open(out / "telolengths_all.csv", newline="").read().split("\r\n")[1:] if ln]
    want = {}
    for stem, sub in (("a.fastq", recs), ("b", recs[10:40]), ("c", recs[:25])):
        want[stem] = [[stem, "5", f"{r['trc']:.3f}", r["id"], str(r["telo_length"])]
                      for r in orc.scan_records(sub, "CCCTAAA", 5, 0.7, 9000, 100, 6, 100, 20000, exact=True)]
    for stem in want:
        assert [r for r in rows if r[0] == stem] == want[stem], stem
    assert len(rows) == sum(len(v) for v in want.values())
    names = set(os.listdir(out))
    assert {"a.fastq_trc_over_0.7.fastq", "b_trc_over_0.7.fastq", "c_trc_over_0.7.fasta"} <= names
    sub_c = open(out / "c_trc_over_0.7.fasta").read()
    assert [ln[1:] for ln in sub_c.split("\n") if ln.startswith(">")] == [r[3] for r in want["c"]]
    assert all(len(ln) <= 60 for ln in sub_c.split("\n"))


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_cli_ends_first_matches_reference_outputs(case, tmp_path, monkeypatch):
    """--ends-first (head + tail upload, then the regions of the TRC-pass reads): the same bytes out."""
    fake_engine.install(monkeypatch)
    run_case(case, tmp_path, extra_argv=["--ends-first"])


def test_ends_first_small_batches_split_regions(tmp_path, monkeypatch):
    """Tiny batches, two stand-in devices, region batches cut by max_pass_reads: same reads, same order."""
    fake_engine.install(monkeypatch)
    from topsicle_b200 import pipeline
    from topsicle_b200.patterns import patterns_to_search
    cfgs = [pipeline.ScanConfig(patterns=patterns_to_search("CCCTAA", k), len_telopattern=6, phrase=k, slide=6,
                                cutoff=0.4, want_rawcount=(k == 5)) for k in (4, 5)]
    path = os.path.join(GOLD, "demo.fastq.gz")
    _, want = pipeline.collect_file(path, cfgs)
    _, got = pipeline.collect_file(path, cfgs, ends_first=True, max_pass_reads=3, max_batch_reads=7, devices=(0, 1),
                                   ends_raw_bytes=200_000)
    for w, g in zip(want, got):
        assert len(w) > 10
        assert [(p.index, p.read_id, p.tail, p.count, p.status, p.n_windows, p.telo_length, p.length) for p in g] == \
               [(p.index, p.read_id, p.tail, p.count, p.status, p.n_windows, p.telo_length, p.length) for p in w]
        for a, b in zip(w, g):
            assert (a.counts is None) == (b.counts is None)
            if a.counts is not None:
                assert np.array_equal(a.counts, b.counts)


@pytest.mark.parametrize("ends_first", ["0", "1"])
def test_directory_with_bad_files_goes_on(tmp_path, monkeypatch, ends_first):
    """One unreadable file of --inputDir ends THAT file, not the run: the reference walks every file of the
    directory (main.py:224-229) and only logs a parse error (allsteps.py:137-149), keeping the reads parsed before
    a bad record and processing the other files.  A stray text file, an empty file and a FASTQ with a truncated
    record next to two good inputs: both good files give their 17 rows, the reads before the bad record count."""
    import gzip
    import shutil
    fake_engine.install(monkeypatch)
    monkeypatch.setenv("TOPSICLE_ENDS_FIRST", ends_first)
    from topsicle_b200 import main as tmain
    ind, out = tmp_path / "in", tmp_path / "out"
    ind.mkdir()
    demo = os.path.join(GOLD, "demo.fastq.gz")
    shutil.copy(demo, ind / "a.fastq.gz")
    shutil.copy(demo, ind / "b.fastq.gz")
    (ind / "notes.txt").write_text("these are not reads\n")
    (ind / "empty.fastq").write_text("")
    text = gzip.open(demo, "rt").read().split("\n")
    # 30 good records, then one whose quality line is shorter than its sequence line, then more good ones
    bad = text[:4 * 30] + [text[120], text[121], "+", text[123][:10]] + text[124:4 * 40] + [""]
    (ind / "trunc.fastq").write_text("\n".join(bad))
    if hasattr(tmain.tprint, "logfile"):
        del tmain.tprint.logfile
    tmain.main(["--inputDir", str(ind), "--outputDir", str(out), "--pattern", "CCCTAAA", "--slide", "6", "--threads", "2"])
    rows = [r.split(",") for r in open(out / "telolengths_all.csv", newline="").read().split("\r\n")[1:] if r]
    by_file = {}
    for r in rows:
        by_file.setdefault(r[0], []).append(r)
    assert len(by_file["a.fastq"]) == 17 and len(by_file["b.fastq"]) == 17
    assert [r[3:] for r in by_file["a.fastq"]] == [r[3:] for r in by_file["b.fastq"]]
    ids30 = {ln.split()[0][1:] for ln in text[:120:4]}
    assert {r[3] for r in by_file.get("trunc", [])} == {r[3] for r in by_file["a.fastq"] if r[3] in ids30}
    log = open(out / "topsicle_run.log").read()
    for name in ("notes.txt", "trunc.fastq"):
        assert f"Error occurred while parsing file {ind / name}" in log
    assert "All telomere found" in log


def test_reader_error_does_not_hang(tmp_path, monkeypatch):
    """A reader thread that dies (not a parse error: those end only their file) must wake the other readers
    that wait for a free slot; scan_files raises instead of hanging in join()."""
    import threading
    fake_engine.install(monkeypatch)
    from topsicle_b200 import fastx, pipeline
    from topsicle_b200.patterns import patterns_to_search
    cfg = pipeline.ScanConfig(patterns=patterns_to_search("CCCTAAA", 5), len_telopattern=7, phrase=5, slide=6)
    demo = os.path.join(GOLD, "demo.fastq.gz")
    real = fastx.FastxFile.next_spans
    calls = {"n": 0}

    def flaky(self, *a, **kw):
        calls["n"] += 1
        if calls["n"] == 3:
            raise MemoryError("simulated reader failure")
        return real(self, *a, **kw)

    monkeypatch.setattr(fastx.FastxFile, "next_spans", flaky)
    jobs = [pipeline.FileJob(demo, lambda res: None) for _ in range(4)]
    done = {}

    def run():
        try:
            with pipeline.Scanner([cfg], devices=[0], depth=1, max_batch_reads=8, max_batch_bases=1 << 22) as sc:
                sc.scan_files(jobs, readers=4)
        except BaseException as e:  # noqa: BLE001
            done["error"] = e

    t = threading.Thread(target=run, daemon=True)
    t.start()
    t.join(60)
    assert not t.is_alive(), "scan_files hung after a reader error"
    assert isinstance(done.get("error"), MemoryError)


def test_parser_equals_reference():
    """Flag names, dest, nargs, defaults, types, required flags, metavars and help sentences of the `topsicle` CLI
    equal those of the reference's parser (main.py:319-334; tests/golden/cli_parser.json is captured from the
    unmodified Topsicle.main by oracle/make_golden.py), and so does the `--help` text of the shared flags."""
    import argparse
    from tests.conftest import load_json
    from topsicle_b200 import main as tmain
    want = load_json("cli_parser.json")
    p = tmain.build_parser()
    assert p.description == want["description"] and p.formatter_class.__name__ == want["formatter"]
    got = {a.dest: a for a in p._actions if not isinstance(a, argparse._HelpAction)}
    for o in want["options"]:
        a = got.pop(o["dest"])
        assert (list(a.option_strings), a.nargs, a.default, getattr(a.type, "__name__", None), bool(a.required),
                a.metavar, a.help, type(a).__name__) == (o["option_strings"], o["nargs"], o["default"], o["type"],
                                                        o["required"], o["metavar"], o["help"], o["action"]), o["dest"]
    assert sorted(got) == ["devices", "ends_first"]          # what this build adds; both optional
    assert all(not a.required for a in got.values())
    os.environ["COLUMNS"] = "100"
    assert tmain.build_parser(reference_only=True).format_help() == want["help_text"]


def test_console_script_and_package_exports():
    """`topsicle = topsicle_b200.main:main` (reference: setup.py:17-21) and the star-exports of the package
    (reference: Topsicle/__init__.py:1)."""
    import tomllib
    from tests.conftest import REPO
    meta = tomllib.load(open(os.path.join(REPO, "pyproject.toml"), "rb"))
    assert meta["project"]["scripts"] == {"topsicle": "topsicle_b200.main:main"}
    import topsicle_b200
    from topsicle_b200 import main as tmain
    assert callable(tmain.main)
    ns = {}
    exec("from topsicle_b200 import *", ns)
    for name in ("check_file_type", "pattern_scramble_telo", "patterns_to_search", "unzip_file", "patternTRC_count",
                 "seq_cut_windows", "bound_detect", "rawCountPattern", "fit_quadratic_and_find_vertex", "plot_patterns"):
        assert callable(ns[name]), name
    assert topsicle_b200.patternTRC_count is ns["patternTRC_count"]


@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
def test_record_longer_than_a_batch_is_scanned_by_its_ends(tmp_path, monkeypatch, fmt):
    """A FASTA of contigs: records far longer than a whole batch (chromosomes) are delivered as their first and
    last max(maxlengthtelo, 1000) bases, which is every base the scan looks at -- same rows and the same subset
    records as with a batch that holds them whole (the reference reads such files like any other)."""
    fake_engine.install(monkeypatch)
    from topsicle_b200 import pipeline
    from topsicle_b200.patterns import patterns_to_search
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)

    def rand(n):
        return bytes(rng.choice(acgt, n)).decode()
    seqs = [("short1", "CCCTAA" * 400 + rand(9000)),
            ("chrF", "CCCTAA" * 700 + rand(400_000)),                      # forward telomere, 404 kb
            ("chrN", rand(350_000)),                                        # no telomere
            ("chrR", rand(500_000) + "TTAGGG" * 900),                       # reverse telomere
            ("short2", rand(12000) + "TTAGGG" * 300)]
    path = tmp_path / f"contigs.{fmt}"
    with open(path, "w") as fh:
        for name, s in seqs:
            if fmt == "fasta":
                fh.write(f">{name} len={len(s)}\n")
                fh.writelines(s[i:i + 60] + "\n" for i in range(0, len(s), 60))
            else:
                fh.write(f"@{name}\n{s}\n+\n{'I' * len(s)}\n")
    cfg = pipeline.ScanConfig(patterns=patterns_to_search("CCCTAA", 4), len_telopattern=6, phrase=4, slide=6,
                              maxlengthtelo=20000, want_rawcount=True)
    _, whole = pipeline.collect_file(str(path), [cfg], max_batch_bases=1 << 20, records_cfg=0)
    _, clipped = pipeline.collect_file(str(path), [cfg], max_batch_bases=100_000, records_cfg=0)
    assert [p.read_id for p in whole[0]] == ["short1", "chrF", "chrR", "short2"]

    def key(p):
        return (p.index, p.read_id, p.tail, round(p.trc, 6), p.status, p.n_windows, p.telo_length, p.record,
                p.counts.tobytes())
    assert [key(p) for p in clipped[0]] == [key(p) for p in whole[0]]
    assert [p.length for p in clipped[0]] == [len(seqs[0][1]), 40000, 40000, len(seqs[4][1])]
    # a batch too small for the two ends: the file ends with the reader's capacity error, as before
    from topsicle_b200 import fastx
    with pytest.raises(fastx.FastxError) as e:
        pipeline.collect_file(str(path), [cfg], max_batch_bases=30_000)
    assert e.value.code == -4


def test_recommended_cutoff_clamps():
    """The reference's clamps on the fitted vertex (main.py:277-291): every branch, values and log wording."""
    from topsicle_b200.main import recommended_cutoff as rc
    assert rc(0.82, 1.1, 0.9, 0.7) == (0.82, [])                                   # inside the data: kept
    x, notes = rc(1.4, 1.1, 0.93, 0.7)                                              # right of the data: the median
    assert x == 0.93 and notes == ["Asymptotic TRC 1.400 is greater than max TRC, which is not expected. See plot.",
                                   "Using median TRC value (0.930) as asymptotic TRC instead."]
    x, notes = rc(1.4, 1.2, 1.05, 0.7)                                              # ... or 0.9 when the median is >= 1
    assert x == 0.9 and notes[1] == "Using 0.9 as asymptotic TRC instead, since asymptotic is greater than 1.0."
    x, notes = rc(0.35, 0.9, 0.8, 0.3)                                              # below 0.4 but above the input cutoff
    assert x == 0.35 and notes == ["Quadratic fit suggests asymptotic TRC less than 0.4. See plot with fit line"]
    x, notes = rc(0.2, 0.38, 0.3, 0.3)                                              # low data and below the input cutoff
    assert x == 0.3 and len(notes) == 3
    assert notes[1] == ("Maximum TRC value in data is 0.380, which is less than 0.4, indicating low confidence in "
                        "telomere detection.")
    assert notes[2] == ("Asymptotic TRC 0.200 is less than input cutoff 0.300. Topsicle declares input TRC (=0.3) as "
                        "asymptotic TRC.")
    x, notes = rc(0.9, 0.5, 0.35, 0.36)                                             # both clamps in a row
    assert x == 0.36 and len(notes) == 4 and notes[3].startswith("Asymptotic TRC 0.350 is less than input cutoff 0.360")
