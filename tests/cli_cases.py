"""Shared checker: run the `topsicle` CLI of this repo on the demo input and compare every output
with what the UNMODIFIED reference wrote for the same flags (tests/golden/demo_cli.json, produced
by oracle/make_golden.py)."""
import ast
import hashlib
import os
import shutil

from tests.conftest import GOLD, load_json

KEEP = ("patterns to search", "k-mer:", "asymptotic TRC", "Asymptotic TRC", "Median telomere", "Using ",
        "Quadratic fit", "Maximum TRC", "Not enough data", "No read has", "All telomere found", "No telophrase")


def golden_cases():
    return load_json("demo_cli.json")["cases"]


def _as_obj(v):
    return ast.literal_eval(v) if isinstance(v, str) and v[:1] in "[{" else v


def run_case(case, tmp_path, extra_argv=()):
    from topsicle_b200 import main as tmain
    indir, out = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    shutil.copy(os.path.join(GOLD, "demo.fastq.gz"), indir / "demo.fastq.gz")
    argv = ["--inputDir", str(indir), "--outputDir", str(out)] + list(_as_obj(case["argv"])) + list(extra_argv)
    if hasattr(tmain.tprint, "logfile"):
        del tmain.tprint.logfile
    tmain.main(argv)
    csv = open(out / "telolengths_all.csv", newline="").read()
    assert csv == case["csv"], case["name"]
    assert hashlib.md5(csv.encode()).hexdigest() == case["csv_md5"]
    files = _as_obj(case["files"])
    for name, md5 in files.items():
        if name.endswith(".png"):
            continue
        got = hashlib.md5(open(out / name, "rb").read()).hexdigest()
        assert got == md5, (case["name"], name)
    extra = set(os.listdir(out)) - set(files) - {"telolengths_all.csv", "topsicle_run.log"}
    assert not {e for e in extra if not e.endswith(".png")}, extra
    log = open(out / "topsicle_run.log").read().splitlines()
    summary = [ln.split("] ", 1)[1] for ln in log if any(k in ln for k in KEEP)]
    want = _as_obj(case["summary"])
    # the reference prints one "patterns to search" line per phrase inside its phrase loop; same lines here
    assert sorted(summary) == sorted(want), (case["name"], summary, want)
    assert summary[-1] == "All telomere found, have a nice day."
