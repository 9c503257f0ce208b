"""Drop-in for the data outputs of the reference's `overview_plot.py` (same flags): the reads of every input
file with TRC > 0.7 are kept (overview_plot.py:60-84), and `--recfindingpattern --rawcount` writes
`heatmap_rawcount_{i}.csv` -- which characters follow each telomere k-mer in the first / last 2 kb of those
reads (descriptive_plot.py:233-313) -- with the matching done on the GPU (`tps_follow_scan`).  The PNG figures
(`descriptive_plot_{i}.png`, `heatmap_{i}.png`) need matplotlib + seaborn and are out of this build's scope: they
are skipped with a note.

    python -m topsicle_b200.overview_plot --inputDir reads/ --outputDir out/ --pattern CCCTAAA \
        --recfindingpattern --rawcount
"""
from __future__ import annotations

import argparse
import os

from . import descriptive
from .allsteps import patternTRC_count, unzip_file


def plot_running(args):
    os.makedirs(args.outputDir, exist_ok=True)
    if os.path.isdir(args.inputDir):
        filenames = [os.path.join(root, f) for root, _, files in os.walk(args.inputDir) for f in files]
    else:
        filenames = [args.inputDir]
    if args.telophrase is None:
        telo_phrases = [len(args.pattern) - 2]
        print(f"No telophrase provided, use kmer: {telo_phrases}")
    else:
        telo_phrases = args.telophrase if isinstance(args.telophrase, list) else [args.telophrase]
    kept = []          # per input file with at least one TRC-pass read: its (name, sequence) records
    for seq_loc in filenames:
        rows = patternTRC_count(seq_loc, telopattern=args.pattern, read_length=args.minSeqLength,
                                kmer=telo_phrases[0], no_bp=1000, cutoff=0.7)
        if rows:
            keep = {r[0] for r in rows}
            kept.append([(r.name, str(r.seq)) for r in unzip_file(seq_loc) if r.id in keep])
    print("Loaded all data, start plotting")
    print("descriptive_plot_{i}.png / heatmap_{i}.png are not drawn by this build (figures only; no matplotlib path)")
    if args.recfindingpattern:
        for i, recs in enumerate(kept, start=1):
            for phrase in telo_phrases:
                fwd, rev = descriptive.heatmap_rows(recs, args.pattern, phrase, args.minSeqLength)
                if args.rawcount:
                    csv_path = f"{args.outputDir}/heatmap_rawcount_{i}.csv"
                    print(f"Saving raw count of heatmap to {csv_path}")
                    with open(csv_path, "w") as fh:
                        fh.write(descriptive.heatmap_csv_text(fwd, rev))
    print(f"Heatmap is in here: {args.outputDir}")
    return "plotted the plot"


def build_parser():
    """Same flags, types and defaults as the reference's overview_plot.py:122-134."""
    p = argparse.ArgumentParser(description="Overview outputs of the telomere scan (B200 build: CSV data only)")
    p.add_argument("--inputDir", type=str, help="input file or directory of FASTQ / FASTA(.gz) files")
    p.add_argument("--outputDir", type=str, help="directory for the outputs")
    p.add_argument("--pattern", metavar="CHAR", type=str, required=True,
                   help="telomere repeat, 5' to 3' (e.g. CCCTAA for human)")
    p.add_argument("--minSeqLength", type=int, default=9000, help="skip reads no longer than this (default 9000)")
    p.add_argument("--telophrase", nargs="+", type=int,
                   help="k-mer length(s) to search; default: len(pattern) - 2")
    p.add_argument("--recfindingpattern", action="store_true",
                   help="compute the k-mer vs following-characters table (the reference's heat map)")
    p.add_argument("--rawcount", action="store_true", help="write that table to heatmap_rawcount_{i}.csv")
    return p


def main(argv=None):
    return plot_running(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
