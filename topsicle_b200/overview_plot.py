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
    p = argparse.ArgumentParser(description="Command line input handling for run_analysis function")
    p.add_argument("--inputDir", type=str, help="Path to the input folder directory")
    p.add_argument("--outputDir", type=str, help="Path to the output folder directory")
    p.add_argument("--pattern", metavar="CHAR", type=str, required=True,
                   help="Required, Telomere repeat sequence (in 5' to 3' orientation). For e.g., in human use CCCTAA")
    p.add_argument("--minSeqLength", type=int, default=9000, help="Minimum of long read sequence, default = 9kbp")
    p.add_argument("--telophrase", nargs="+", type=int,
                   help="Length of telomere k-mer to search. By default will use telomere k-mer length minus 2")
    p.add_argument("--recfindingpattern", action="store_true",
                   help="Optional, use this to plot the heatmap of patterns vs match")
    p.add_argument("--rawcount", action="store_true",
                   help="Optional, save raw count results to CSV for flexibility of plotting")
    return p


def main(argv=None):
    return plot_running(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
