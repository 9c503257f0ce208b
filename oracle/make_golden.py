#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference's own
modules are imported from where they lie; the four third-party imports that are
not installable here (Bio, ruptures, seaborn, matplotlib) come from
`oracle/shims/` (see the docstrings there; `ruptures` is the only shim that
carries arithmetic and is a restatement of ruptures==1.1.9).

Outputs (committed, small):
  tests/golden/demo.fastq.gz          copy of the reference's demo INPUT data
  tests/golden/demo_cli.json          CLI known answers (CSV text, md5s, log summary lines)
  tests/golden/demo_step1.json        patternTRC_count rows for every read (cutoff -1)
  tests/golden/demo_rawcount.npz      rawCountPattern count tables for a few demo reads
  tests/golden/demo_heatmap.json      patterns_vs_match_heatmap CSV md5s (overview_plot.py --recfindingpattern --rawcount)
  tests/golden/edge.fastq / edge.fasta   crafted edge-case reads
  tests/golden/edge.json              reference function outputs on them
  tests/golden/cli_parser.json        the reference CLI's argparse table (main.py:319-334) and --help text

Usage:  python oracle/make_golden.py
"""
import contextlib
import hashlib
import io
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(REPO, "tests", "golden")
DEMO_IN = os.path.join(REF, "Topsicle_demo", "data_col0_teloreg_chr",
                       "Col-0-6909_GWHBDNP00000001.1_nano_right.fastq.gz")
DEMO_GOLD_CSV = os.path.join(REF, "Topsicle_demo", "telolengths_all.csv")
DEMO_GOLD_SUBSET = os.path.join(
    REF, "Topsicle_demo", "result_justone",
    "Col-0-6909_GWHBDNP00000001.1_nano_right.fastq_trc_over_0.7.fastq")

sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(1, REF)

import numpy as np  # noqa: E402


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def run_cli(argv):
    """Run Topsicle.main.main() with argv; returns (outdir files dict, log summary lines)."""
    import Topsicle.main as tmain
    out = tempfile.mkdtemp(prefix="tps_gold_")
    full = ["topsicle", "--outputDir", out, "--threads", "1"] + argv
    old = sys.argv
    sys.argv = full
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            tmain.main()
    finally:
        sys.argv = old
    log = open(os.path.join(out, "topsicle_run.log")).read().splitlines()
    keep = ("patterns to search", "k-mer:", "asymptotic TRC", "Asymptotic TRC", "Median telomere",
            "Using ", "Quadratic fit", "Maximum TRC", "Not enough data", "No read has",
            "All telomere found", "No telophrase")
    summary = [ln.split("] ", 1)[1] for ln in log if any(k in ln for k in keep)]
    return out, summary


def demo_cli_cases():
    cases = [
        ("golden_CCCTAAA_slide6", ["--pattern", "CCCTAAA", "--slide", "6"]),
        ("config1_AAACCCT_defaults", ["--pattern", "AAACCCT"]),
        ("CCCTAAA_slide6_phrase4", ["--pattern", "CCCTAAA", "--slide", "6", "--telophrase", "4"]),
        ("readme_detailed", ["--pattern", "CCCTAAA", "--telophrase", "4", "--cutoff", "0.4",
                             "--slide", "6", "--trimfirst", "200"]),
        ("CCCTAA_sweep_456_cut04_07", ["--pattern", "CCCTAA", "--telophrase", "4", "5", "6",
                                       "--cutoff", "0.4", "0.7"]),
        ("stress_CCCTAAA_w50_s3", ["--pattern", "CCCTAAA", "--windowSize", "50", "--slide", "3"]),
        ("TTTAGGG_no_rows", ["--pattern", "TTTAGGG"]),
        ("CCCTAAA_phrase7_min0_max5000", ["--pattern", "CCCTAAA", "--telophrase", "7", "--cutoff", "0.3",
                                          "--minSeqLength", "0", "--maxlengthtelo", "5000",
                                          "--trimfirst", "0", "--slide", "1"]),
        ("CCCTAAA_rawcount", ["--pattern", "CCCTAAA", "--slide", "6", "--rawcountpattern",
                              "--cutoff", "0.88"]),
    ]
    res = []
    for name, argv in cases:
        indir = tempfile.mkdtemp(prefix="tps_in_")
        shutil.copy(DEMO_IN, os.path.join(indir, "demo.fastq.gz"))
        out, summary = run_cli(["--inputDir", indir] + argv)
        files = sorted(os.listdir(out))
        csv = open(os.path.join(out, "telolengths_all.csv"), newline="").read()
        entry = dict(name=name, argv=argv, csv=csv,
                     csv_md5=md5(os.path.join(out, "telolengths_all.csv")),
                     summary=summary, files={})
        for f in files:
            if f in ("telolengths_all.csv", "topsicle_run.log"):
                continue
            entry["files"][f] = md5(os.path.join(out, f))
        res.append(entry)
        print(f"  {name}: {csv.count(chr(10)) - 1} rows, md5 {entry['csv_md5']}, files {list(entry['files'])[:3]}")
        shutil.rmtree(out)
        shutil.rmtree(indir)
    return res


def check_reference_golden():
    """The reference's own golden CSV / subset FASTQ must come out byte-identical."""
    indir = os.path.dirname(DEMO_IN)
    out, summary = run_cli(["--inputDir", indir, "--pattern", "CCCTAAA", "--slide", "6"])
    got_csv = md5(os.path.join(out, "telolengths_all.csv"))
    got_sub = md5(os.path.join(out, os.path.basename(DEMO_GOLD_SUBSET)))
    assert got_csv == md5(DEMO_GOLD_CSV) == "92c042b7c7e13ae38ba5823370adc6a4", got_csv
    assert got_sub == md5(DEMO_GOLD_SUBSET) == "e6432c8562a283c3bf7e0dd535fde5bd", got_sub
    assert "k-mer: 5, with TRC >= 0.7, median telomere length is 2110.00 bp" in summary
    assert "asymptotic TRC, or recommended cutoff: 0.897" in summary
    assert "Median telomere length for reads with TRC cutoff >= 0.897: 2050.00 bp" in summary
    shutil.rmtree(out)
    print("  reference golden CSV + subset FASTQ + log medians reproduced byte-identically")
    return dict(csv_md5=got_csv, subset_md5=got_sub, summary=summary)


def demo_step1():
    from Topsicle.allsteps import patternTRC_count
    combos = [("CCCTAAA", 5), ("CCCTAAA", 4), ("AAACCCT", 5), ("CCCTAA", 4), ("CCCTAA", 5),
              ("CCCTAA", 6), ("TTTAGGG", 6), ("TTAGGG", 3), ("TTAGGG", 2), ("CCCTAAA", 7)]
    out = []
    for pat, k in combos:
        for minlen in (9000, 0):
            rows = patternTRC_count(DEMO_IN, pat, read_length=minlen, kmer=k, no_bp=1000, cutoff=-1.0)
            out.append(dict(pattern=pat, kmer=k, read_length=minlen,
                            rows=[[r[0], r[1], r[2], repr(float(r[3]))] for r in rows]))
    print(f"  step1: {len(out)} combos, {sum(len(o['rows']) for o in out)} rows")
    return out


def demo_rawcount():
    """rawCountPattern tables (int16 [nW][P]) + bound_detect for a few demo reads."""
    from Topsicle.allsteps import rawCountPattern, bound_detect, patternTRC_count, patterns_to_search
    arrays, meta = {}, []
    cfgs = [("CCCTAAA", 5, 100, 6, 100, 20000), ("CCCTAA", 5, 100, 6, 100, 20000),
            ("CCCTAA", 6, 50, 3, 0, 6000), ("TTAGGG", 3, 100, 7, 100, 3000)]
    for ci, (pat, k, W, s, t, M) in enumerate(cfgs):
        rows = patternTRC_count(DEMO_IN, pat, read_length=9000, kmer=k, cutoff=0.5)
        plist = patterns_to_search(pat, k)
        for r in rows[:3]:
            rid, tail = r[0], r[2]
            df = rawCountPattern(DEMO_IN, rid, plist, W, s, t, k, 9000, M, tail=tail)
            nP = len(plist)
            cnt = df["count"].to_numpy().reshape(-1, nP).astype(np.int16)
            pos = df["position"].to_numpy().reshape(-1, nP)[:, 0]
            assert list(df["pattern"][:nP]) == plist and set(df["tail"]) == {tail}
            bd = bound_detect(DEMO_IN, rid, plist, W, s, t, M, k, tail=tail)
            key = f"c{ci}_{rid}"
            arrays[key] = cnt
            meta.append(dict(key=key, pattern=pat, kmer=k, windowSize=W, slide=s, trimfirst=t,
                             maxlengthtelo=M, read=rid, tail=tail, patterns=plist,
                             first_pos=int(pos[0]), last_pos=int(pos[-1]), n_windows=int(len(pos)),
                             telo_length=int(bd[0][1])))
    print(f"  rawcount: {len(meta)} tables")
    return arrays, meta


# ------------------------------------------------------------------ crafted edge-case reads
def demo_heatmap():
    """`patterns_vs_match_heatmap(...).to_csv(index=False)` of the unmodified reference (descriptive_plot.py:233-313,
    overview_plot.py:98-108).  The first case is the reference's own golden
    Topsicle_demo/result_justone/heatmap_rawcount_1.csv: the reads of the demo with TRC > 0.7 (phrase 5)."""
    import contextlib
    import importlib
    import io
    importlib.import_module("Topsicle.descriptive_plot")
    dp = sys.modules["Topsicle.descriptive_plot"]
    from Topsicle.allsteps import patternTRC_count
    from Bio import SeqIO
    import gzip
    gold_md5 = md5(os.path.join(REF, "Topsicle_demo", "result_justone", "heatmap_rawcount_1.csv"))
    out = []
    edge_fq = os.path.join(GOLD, "edge.fastq")
    cases = [("CCCTAAA", 5, 9000, "subset", DEMO_IN), ("CCCTAA", 4, 9000, "all", DEMO_IN),
             ("CCCTAA", 6, 9000, "all", DEMO_IN), ("AAACCCT", 4, 20000, "all", DEMO_IN),
             ("TTTAGGG", 6, 0, "all", DEMO_IN), ("CCCTAA", 4, 0, "all", edge_fq), ("CCCTAAA", 5, 150, "all", edge_fq)]
    for pat, k, minlen, mode, src in cases:
        path = src
        tmp = None
        if mode == "subset":      # overview_plot.py:63-84: reads with TRC > 0.7 written to a temp file first
            with contextlib.redirect_stdout(io.StringIO()):
                keep = {r[0] for r in patternTRC_count(src, telopattern=pat, read_length=minlen, kmer=k, no_bp=1000,
                                                       cutoff=0.7)}
            tmp = tempfile.mktemp(suffix=".fastq")
            with gzip.open(src, "rt") as ih, open(tmp, "w") as oh:
                for rec in SeqIO.parse(ih, "fastq"):
                    if rec.id in keep:
                        SeqIO.write(rec, oh, "fastq")
            path = tmp
        with contextlib.redirect_stdout(io.StringIO()):
            df = dp.patterns_vs_match_heatmap(path, pat, k, minlen)
        txt = df.to_csv(index=False)
        if tmp:
            os.remove(tmp)
        e = dict(pattern=pat, telophrase=k, minSeqLength=minlen, input=os.path.basename(src), mode=mode,
                 rows=len(df), md5=hashlib.md5(txt.encode()).hexdigest())
        if mode == "subset":
            assert e["md5"] == gold_md5 == "28ad064f247aa236af6f0fedddc63ed4", e
            e["reference_golden"] = "Topsicle_demo/result_justone/heatmap_rawcount_1.csv"
        out.append(e)
        print(f"  heatmap {pat} k={k} {e['input']} {mode}: {e['rows']} rows, md5 {e['md5']}")
    return out


def build_edge_reads():
    rng = np.random.default_rng(20261017)
    B = np.array(list("ACGT"))

    def rnd(n):
        return "".join(B[rng.integers(0, 4, n)])

    def mutate(s, rate):
        a = np.array(list(s))
        hit = rng.random(len(a)) < rate
        a[hit] = B[rng.integers(0, 4, int(hit.sum()))]
        return "".join(a)

    def rc(s):
        return s[::-1].translate(str.maketrans("ACGT", "TGCA"))

    reads = []
    # forward / reverse telomeric reads for several motifs, with substitution noise
    for motif in ("CCCTAA", "CCCTAAA", "AAACCCT", "TTAGGG", "TTTAGGG"):
        for noise in (0.0, 0.03):
            tl = int(rng.integers(900, 2200))
            tel = (motif * (tl // len(motif) + 2))[int(rng.integers(0, len(motif))):][:tl]
            reads.append((f"fwd_{motif}_{noise}", mutate(tel, noise) + rnd(int(rng.integers(1500, 3000)))))
            reads.append((f"rev_{motif}_{noise}", rnd(int(rng.integers(1500, 3000))) + mutate(rc(tel), noise)))
    # lower-case / mixed-case telomere
    tel = "CCCTAA" * 300
    reads.append(("lower_fwd", tel.lower() + rnd(2500)))
    reads.append(("mixed_rev", rnd(2000) + "".join(c.lower() if i % 3 else c for i, c in enumerate(rc(tel)))))
    # N and IUPAC inside the telomere
    t2 = list("CCCTAAA" * 250)
    for i in rng.integers(0, len(t2), 40):
        t2[i] = "N"
    for i in rng.integers(0, len(t2), 10):
        t2[i] = "R"
    reads.append(("with_N_fwd", "".join(t2) + rnd(2600)))
    reads.append(("all_N", "N" * 1500))
    # homopolymers / low-complexity: self-overlapping k-mers (AA, AAA, CCC ...)
    reads.append(("polyA", "A" * 2400))
    reads.append(("polyAC", "AC" * 1300))
    reads.append(("polyC_then_rand", "C" * 1100 + rnd(1500)))
    reads.append(("ctaac_chain", "CTAA" * 400 + "C" + rnd(1300)))   # CTAAC overlaps itself at shift 4
    # short reads (< 1000 bp: head and tail are the whole read) and tiny reads
    reads.append(("short_700", ("CCCTAA" * 60)[:350] + rnd(350)))
    reads.append(("short_999", rnd(500) + rc("CCCTAAA" * 72)[:499]))
    reads.append(("len_1000", ("TTAGGG" * 100)[:600] + rnd(400)))
    reads.append(("len_1001", rnd(400) + ("TTAGGG" * 101)[:601]))
    reads.append(("tiny_150", "CCCTAA" * 25))
    reads.append(("tiny_5", "CCCTA"))
    # exact head/tail tie (tie -> reverse): read = X + reverse(X)
    x = "CCCTAA" * 170
    reads.append(("tie_palin", x + x[::-1]))
    # no telomere at all
    reads.append(("random_3000", rnd(3000)))
    reads.append(("random_12000", rnd(12000)))
    # telomere longer than maxlengthtelo used in tests; interior telomere-like block
    reads.append(("long_tel", mutate("CCCTAAA" * 900, 0.02) + rnd(1200)))
    reads.append(("interior_block", rnd(1500) + "CCCTAA" * 200 + rnd(1500)))
    # two change levels (partial degradation)
    reads.append(("two_level", "CCCTAA" * 150 + mutate("CCCTAA" * 150, 0.25) + rnd(2000)))
    return reads


def edge_cases(fastq_path, fasta_path):
    from Topsicle.allsteps import (patternTRC_count, bound_detect, rawCountPattern,
                                   patterns_to_search, seq_cut_windows)
    import ruptures
    out = dict(patterns=[], step1=[], step2=[], windows=[])
    # pattern expansion table
    for motif in ("CCCTAA", "AACCCT", "TTAGGG", "AAACCCT", "CCCTAAA", "TTTAGGG", "TTAGG", "AT", "ACGT", "cccTaa"):
        for k in range(1, len(motif) + 1):
            out["patterns"].append(dict(motif=motif, k=k, patterns=patterns_to_search(motif, k)))
    # seq_cut_windows
    for s_len, W, st in ((250, 100, 6), (100, 100, 7), (99, 100, 3), (136, 100, 6), (1000, 50, 3)):
        s = "".join("ACGT"[i % 4] for i in range(s_len))
        win = seq_cut_windows(s, W, st)
        out["windows"].append(dict(len=s_len, W=W, step=st, starts=[w[0] for w in win],
                                   lens=[len(w[1]) for w in win]))
    # step 1 on both files
    for path in (fastq_path, fasta_path):
        for motif, k, minlen, no_bp in (("CCCTAA", 4, 0, 1000), ("CCCTAA", 5, 0, 1000), ("CCCTAA", 6, 0, 1000),
                                        ("CCCTAAA", 5, 0, 1000), ("AAACCCT", 5, 1000, 1000),
                                        ("TTAGGG", 2, 0, 1000), ("TTAGGG", 3, 0, 1000),
                                        ("TTTAGGG", 7, 0, 1000), ("CCCTAA", 4, 2400, 1000),
                                        ("CCCTAA", 4, 0, 300), ("CCCTAAA", 5, 999, 2000)):
            rows = patternTRC_count(path, motif, read_length=minlen, kmer=k, no_bp=no_bp, cutoff=-1.0)
            out["step1"].append(dict(file=os.path.basename(path), motif=motif, k=k, read_length=minlen,
                                     no_bp=no_bp,
                                     rows=[[r[0], r[1], r[2], repr(float(r[3]))] for r in rows]))
    # step 2/3 on the fastq: every read x several parameter sets, both tails
    ids = [r[0] for r in build_edge_reads()]
    cfgs = [("CCCTAA", 4, 100, 6, 100, 20000), ("CCCTAA", 5, 100, 6, 100, 2000),
            ("CCCTAA", 6, 50, 3, 0, 1800), ("CCCTAAA", 5, 100, 7, 100, 20000),
            ("TTAGGG", 3, 100, 6, 50, 2500), ("TTAGGG", 2, 30, 1, 10, 400),
            ("TTTAGGG", 7, 64, 5, 33, 1500), ("AAACCCT", 5, 100, 7, 200, 20000)]
    for motif, k, W, s, t, M in cfgs:
        plist = patterns_to_search(motif, k)
        for rid in ids:
            for tail in ("forward", "reverse"):
                rec = dict(motif=motif, k=k, W=W, slide=s, trimfirst=t, maxlengthtelo=M, read=rid, tail=tail)
                try:
                    bd = bound_detect(fastq_path, rid, plist, W, s, t, M, k, tail=tail)
                    rec["telo_length"] = int(bd[0][1]) if bd else None
                except ruptures.BadSegmentationParameters:
                    rec["error"] = "BadSegmentationParameters"
                df = rawCountPattern(fastq_path, rid, plist, W, s, t, k, 0, M, tail=tail)
                cnt = df["count"].to_numpy().reshape(-1, len(plist))
                rec["n_windows"] = int(cnt.shape[0])
                rec["c_w"] = [int(v) for v in cnt.sum(axis=1)]
                rec["counts_md5"] = hashlib.md5(cnt.astype(np.int16).tobytes()).hexdigest()
                out["step2"].append(rec)
    print(f"  edge: {len(out['patterns'])} pattern sets, {len(out['step1'])} step1 runs, "
          f"{len(out['step2'])} step2 rows")
    return out


def cli_parser_spec():
    """The reference's argparse table (main.py:319-334), captured from the unmodified Topsicle.main.main(): every
    option with its strings, dest, nargs, default, type, required flag, metavar and help text, plus the parser's
    description and the text of `topsicle --help`."""
    import argparse
    import Topsicle.main as tmain
    seen = {}

    class Stop(Exception):
        pass

    def capture(self, *a, **kw):
        seen["parser"] = self
        raise Stop()

    old_parse, old_argv = argparse.ArgumentParser.parse_args, sys.argv
    argparse.ArgumentParser.parse_args = capture
    sys.argv = ["topsicle"]
    try:
        tmain.main()
    except Stop:
        pass
    finally:
        argparse.ArgumentParser.parse_args = old_parse
        sys.argv = old_argv
    parser = seen["parser"]
    parser.prog = "topsicle"
    opts = []
    for act in parser._actions:
        if isinstance(act, argparse._HelpAction):
            continue
        opts.append(dict(option_strings=list(act.option_strings), dest=act.dest, nargs=act.nargs, default=act.default,
                         type=getattr(act.type, "__name__", None), required=bool(act.required),
                         metavar=act.metavar, help=act.help, action=type(act).__name__))
    os.environ["COLUMNS"] = "100"
    return dict(description=parser.description, formatter=parser.formatter_class.__name__, options=opts,
                help_text=parser.format_help())


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--cli-parser" in sys.argv:   # only the argparse table of the reference CLI
        json.dump(cli_parser_spec(), open(os.path.join(GOLD, "cli_parser.json"), "w"), indent=1)
        return
    if "--heatmap" in sys.argv:      # only the overview heat-map fixtures (the others are left as committed)
        json.dump(demo_heatmap(), open(os.path.join(GOLD, "demo_heatmap.json"), "w"), indent=1)
        return
    print("[1] reference golden check")
    ref_check = check_reference_golden()
    shutil.copy(DEMO_IN, os.path.join(GOLD, "demo.fastq.gz"))
    os.chmod(os.path.join(GOLD, "demo.fastq.gz"), 0o644)
    print("[2] demo CLI cases")
    cli = demo_cli_cases()
    json.dump(dict(reference_golden=ref_check, cases=cli), open(os.path.join(GOLD, "demo_cli.json"), "w"), indent=1)
    print("[3] demo step 1")
    json.dump(demo_step1(), open(os.path.join(GOLD, "demo_step1.json"), "w"), indent=0)
    print("[4] demo rawcount tables")
    arrays, meta = demo_rawcount()
    np.savez_compressed(os.path.join(GOLD, "demo_rawcount.npz"), **arrays)
    json.dump(meta, open(os.path.join(GOLD, "demo_rawcount.json"), "w"), indent=1)
    print("[5] edge-case reads")
    reads = build_edge_reads()
    fq = os.path.join(GOLD, "edge.fastq")
    fa = os.path.join(GOLD, "edge.fasta")
    with open(fq, "w") as f:
        for rid, s in reads:
            f.write(f"@{rid} extra description\n{s}\n+\n{'I' * len(s)}\n")
    with open(fa, "w") as f:
        for rid, s in reads:
            f.write(f">{rid} extra description\n")
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + "\n")
    json.dump(edge_cases(fq, fa), open(os.path.join(GOLD, "edge.json"), "w"), indent=0)
    print("[6] overview heat map")
    json.dump(demo_heatmap(), open(os.path.join(GOLD, "demo_heatmap.json"), "w"), indent=1)
    print("[7] CLI argparse table")
    json.dump(cli_parser_spec(), open(os.path.join(GOLD, "cli_parser.json"), "w"), indent=1)
    print("done")


if __name__ == "__main__":
    main()
