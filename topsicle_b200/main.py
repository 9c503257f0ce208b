"""`topsicle` command line -- drop-in for Topsicle/main.py on a B200 box.

Same flags, defaults and types as the reference parser (main.py:319-334), same outputs in
`--outputDir`: `telolengths_all.csv` (header + CRLF rows `file,phrase,trc,readID,telo_length`,
main.py:198-200,136-138), `{stem}_trc_over_{cutoff}.fastq|fasta` (main.py:64-87),
`topsicle_run.log` (main.py:31-45), optional `rawcount_{phrase}_{i}.csv` (main.py:146-150),
`plot_{phrase}_{i}.png`, `quadfit_{phrase}mer_{pattern}.png`, and the same log sentences
including the final "All telomere found, have a nice day." (main.py:309).

What differs is how the work is done: the reference forks one process per input FILE and
re-parses the file once per telophrase and once per TRC-pass read; here every file is parsed
once, all telophrases are scanned from that one pass, and the batches of reads are dealt to
all visible GPUs (`--devices`, default: every CUDA device).  `--threads` sets the host
parser threads.  Rows of one file are always in file order (the reference interleaves files
nondeterministically).
"""
from __future__ import annotations

import argparse
import csv
import datetime
import os
import sys
import threading
import time
from collections import defaultdict

import numpy as np

from . import engine, fastx, pipeline
from .allsteps import _plot_boundary, fit_quadratic_and_find_vertex
from .patterns import patterns_to_search, validate_literals

version_number = "1.0.0"
Topsicle_output_prefix = "Topsicle"
_csv_lock = threading.Lock()


def get_log_path(args):
    log_dir = getattr(args, "outputDir", ".")
    os.makedirs(log_dir, exist_ok=True)
    return os.path.join(log_dir, "topsicle_run.log")


def tprint(*args, **kwargs):
    """Timestamped line to stdout and to topsicle_run.log (main.py:37-45)."""
    msg = " ".join(str(a) for a in args)
    now = datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S")
    line = f"[{now}] {msg}"
    print(line)
    if hasattr(tprint, "logfile"):
        with open(tprint.logfile, "a") as f:
            f.write(line + "\n")


def visible_devices():
    """CUDA devices to use: TOPSICLE_DEVICES, else all devices the runtime reports."""
    env = os.environ.get("TOPSICLE_DEVICES")
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    n = engine.device_count()
    return list(range(n)) if n > 0 else [0]


def subset_path(args, seq_loc, min_cutoff):
    """Name and format of the TRC subset file (main.py:53-80): `.fastq` for .fastq/.fq(.gz)
    inputs, `.fasta` otherwise; an existing `.fasta`-named file is reused."""
    file_name = os.path.splitext(os.path.basename(seq_loc))[0]
    fasta_temp = os.path.join(args.outputDir, f"{file_name}_trc_over_{min_cutoff}.fasta")
    if os.path.exists(fasta_temp):
        return file_name, fasta_temp, True
    name = seq_loc[:-3] if seq_loc.endswith(".gz") else seq_loc
    if name.endswith(".fastq") or name.endswith(".fq"):
        fasta_temp = os.path.join(args.outputDir, f"{file_name}_trc_over_{min_cutoff}.fastq")
    return file_name, fasta_temp, False


class _FileWriter:
    """Ordered sink of one input file: subset records, CSV rows, rawcount tables, plots."""

    def __init__(self, args, seq_loc, cfgs, phrases, min_cutoff, sliding_val, csv_path):
        self.args = args
        self.cfgs = cfgs
        self.phrases = phrases
        self.sliding_val = sliding_val
        self.csv_path = csv_path
        self.file_name, self.subset, self.subset_exists = subset_path(args, seq_loc, min_cutoff)
        name = seq_loc[:-3] if seq_loc.endswith(".gz") else seq_loc
        # the subset holds the reads passing ONE phrase: with FASTQ input the reference rewrites it for
        # every phrase (last one stays), with FASTA input the first phrase's file is reused (main.py:64-66)
        is_fastq_name = name.endswith(".fastq") or name.endswith(".fq")
        self.records_cfg = None if self.subset_exists else (len(cfgs) - 1 if is_fastq_name else 0)
        # opened at the file's first batch, not here: a directory may hold more files than `ulimit -n`, and an
        # aborted run must not leave truncated subsets of files it never reached (main.py:82 opens per file)
        self.subset_handle = None
        self._closed = False
        self.rows = [[] for _ in cfgs]          # per phrase: CSV rows (held: the CSV is phrase-major)
        self.results = [[] for _ in cfgs]       # per phrase: (telolen, trc)
        self.image_num = [1 for _ in cfgs]
        self.n_badseg = 0
        self.stats_scanned = 0

    def __call__(self, res: pipeline.BatchResult):
        args = self.args
        self.stats_scanned += res.n_scanned
        new_first = []   # rows of the first phrase go to the CSV as they are found (main.py:136: "real time")
        for k, passes in enumerate(res.passes):
            cfg, phrase = self.cfgs[k], self.phrases[k]
            for p in passes:
                if self.records_cfg == k and p.record is not None:
                    self._subset().write(p.record)
                if args.read_check and p.read_id != args.read_check:
                    continue
                if p.status != engine.ST_PASS:
                    # the reference dies here: ruptures.BadSegmentationParameters (1..6 windows) or
                    # IndexError on an empty boundary list (0 windows), main.py:133
                    self.n_badseg += 1
                    tprint(f"WARNING: read {p.read_id} passes TRC but has only {p.n_windows} windows "
                           f"(needs >= 7 for a change point); skipped")
                    continue
                m = min(args.maxlengthtelo, p.length)
                telolen = p.telo_length if (p.telo_length <= m and p.telo_length != 0) else 0
                row = [self.file_name, phrase, f"{p.trc:.3f}", p.read_id, telolen]
                (new_first if k == 0 else self.rows[k]).append(row)
                self.results[k].append((float(telolen), float(p.trc)))
                i = self.image_num[k]
                if args.plot and p.counts is not None:
                    _plot_boundary(p.read_id, p, len(cfg.patterns), args.trimfirst, self.sliding_val, m, telolen,
                                   args.rangecp)
                    try:
                        import matplotlib.pyplot as plt
                        plt.savefig(f"{args.outputDir}/plot_{phrase}_{i}.png", format="png", dpi=300)
                        plt.close()
                    except ImportError:
                        pass
                if args.rawcountpattern and p.counts is not None:
                    text = fastx.format_rawcount_csv(p.counts, self.sliding_val, p.tail, cfg.patterns)
                    with open(f"{args.outputDir}/rawcount_{phrase}_{i}.csv", "wb") as fh:
                        fh.write(text)
                self.image_num[k] += 1
        if new_first:
            with _csv_lock, open(self.csv_path, mode="a", newline="") as file:   # files are scanned concurrently
                csv.writer(file).writerows(new_first)

    def _subset(self):
        if self.subset_handle is None:
            self.subset_handle = open(self.subset, "wb")
        return self.subset_handle

    def close(self, create=True):
        """`create`: the file was processed: the subset exists afterwards even if no read passed (main.py:82)."""
        if self._closed:
            return
        self._closed = True
        if self.subset_handle is None and create and not self.subset_exists:
            self._subset()
        if self.subset_handle is not None:
            self.subset_handle.close()
            self.subset_handle = None


def recommended_cutoff(vertex_x, max_trc, median_trc, inputtrc):
    """The clamps the reference puts on the vertex of its quadratic fit before it recommends a TRC cutoff
    (main.py:277-296), as a function -> (cutoff, log lines in the reference's wording).  A vertex to the right of
    every data point is not believed (the median TRC stands in, or 0.9 when even that is >= 1); a vertex below 0.4
    is reported, and one below the cutoff the run was started with is replaced by that cutoff."""
    notes = []
    x = vertex_x
    if x > max_trc:
        notes.append(f"Asymptotic TRC {x:.3f} is greater than max TRC, which is not expected. See plot.")
        if median_trc < 1.0:
            x = median_trc
            notes.append(f"Using median TRC value ({median_trc:.3f}) as asymptotic TRC instead.")
        else:
            x = 0.9
            notes.append("Using 0.9 as asymptotic TRC instead, since asymptotic is greater than 1.0.")
    if x < 0.4:
        notes.append("Quadratic fit suggests asymptotic TRC less than 0.4. See plot with fit line")
        if max_trc < 0.4:
            notes.append(f"Maximum TRC value in data is {max_trc:.3f}, which is less than 0.4, indicating low "
                         "confidence in telomere detection.")
        if x < inputtrc:
            notes.append(f"Asymptotic TRC {x:.3f} is less than input cutoff {inputtrc:.3f}. Topsicle "
                         f"declares input TRC (={inputtrc}) as asymptotic TRC.")
            x = inputtrc
    return x, notes


def scan_configs(args, telo_phrases, patterns, sliding_val):
    """One ScanConfig per telophrase == the arguments the reference passes down (main.py:57,129-130,147-148)."""
    min_cutoff = min(args.cutoff) if isinstance(args.cutoff, (list, tuple)) else args.cutoff
    want_counts = bool(args.rawcountpattern or args.plot)
    return [pipeline.ScanConfig(patterns=pats, len_telopattern=len(args.pattern), phrase=k, cutoff=min_cutoff,
                                min_seq_length=args.minSeqLength, no_bp=1000, window_size=args.windowSize,
                                slide=sliding_val, trimfirst=args.trimfirst, maxlengthtelo=args.maxlengthtelo,
                                want_rawcount=want_counts)
            for k, pats in zip(telo_phrases, patterns)]


def file_job(args, seq_loc, telo_phrases, scanner, sliding_val):
    """Writer + scan job of one input file, every telophrase in one pass (reference: process_file,
    main.py:52-154, once per phrase and per file).  Returns (writer, FileJob)."""
    tprint("subsetting raw dataset based on TRC cutoff")
    min_cutoff = min(args.cutoff) if isinstance(args.cutoff, (list, tuple)) else args.cutoff
    w = _FileWriter(args, seq_loc, scanner.cfgs, telo_phrases, min_cutoff, sliding_val,
                    f"{args.outputDir}/telolengths_all.csv")
    if w.subset_exists:
        tprint(f"Temporary fasta file already exists: {w.subset}. Using existing file.")

    def on_done(stats):
        w.close()
        w.stats = stats
        if not w.subset_exists:
            tprint(f"Temporary fasta file with TRC more than {min_cutoff}:", w.subset)

    def on_error(e):     # allsteps.py:147-149: the error is logged, the reads parsed before it are kept
        tprint(f"Error occurred while parsing file {seq_loc}: {e}")

    return w, pipeline.FileJob(seq_loc, w, records_cfg=w.records_cfg, on_done=on_done, on_error=on_error)


def analysis_run(args):
    print("---- Topsicle run parameters ---")
    for k, v in vars(args).items():
        tprint(f"{k}: {v}")
    print("---------------------")

    tprint("Starting Topsicle analysis")
    os.makedirs(args.outputDir, exist_ok=True)

    if args.threads is not None:
        num_cores = args.threads
        tprint(f"Specified number of cores are/is: {num_cores}")
    else:
        num_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        tprint(f"By default, Topsicle allocates nmber of cores: {num_cores}")
    explicit_devices = bool(getattr(args, "devices", None) or os.environ.get("TOPSICLE_DEVICES"))
    devices = args.devices if getattr(args, "devices", None) else visible_devices()

    output_csv = f"{args.outputDir}/telolengths_all.csv"
    tprint(f"Output will be here: {output_csv}")
    if os.path.exists(output_csv) and os.path.getsize(output_csv) > 0:
        if args.override:
            tprint(f"Output file {output_csv} already exists and will be overridden becuz having --override flag.")
            os.remove(output_csv)
        else:
            tprint(f"Output file {output_csv} already exists and is not empty. Exiting to avoid overwrite. "
                   "Use --override to force overwrite.")
            sys.exit(1)

    if args.telophrase is None:
        telo_phrases = [len(args.pattern) - 2]
        tprint(f"No telophrase provided, use kmer: {telo_phrases}")
    else:
        telo_phrases = args.telophrase if isinstance(args.telophrase, list) else [args.telophrase]

    print("---------------------")

    with open(output_csv, mode="w", newline="") as file:
        csv.writer(file).writerow(["file_number", "phrase", "trc", "readID", "telo_length"])

    for telo_phrase in telo_phrases:
        if telo_phrase > len(args.pattern):
            tprint("Cannot have length of subset larger than length of pattern")
            tprint(f"Cannot get {telo_phrase}-bp cut from {len(args.pattern)}-bp pattern")
            sys.exit()
    sliding_val = args.slide if args.slide else len(args.pattern)

    patterns = []
    for telo_phrase in telo_phrases:
        pats = patterns_to_search(telopattern=args.pattern, cut_length=telo_phrase)
        validate_literals(pats)
        patterns.append(pats)
        tprint("patterns to search:", pats)

    filenames = []
    if os.path.isdir(args.inputDir):
        for root, dirs, files in os.walk(args.inputDir):
            for filename in files:
                filenames.append(os.path.join(root, filename))
    else:
        filenames.append(args.inputDir)

    tprint("begin processing reads")
    phrase_to_telo = defaultdict(list)
    phrase_to_trc = defaultdict(list)
    csv_rows = [[] for _ in telo_phrases]
    t0 = time.time()
    total_bases = total_reads = 0
    # batch size: 256 MiB of bases for large inputs, less otherwise: three slots per device are page-locked at
    # about 1.4 GB/s, which for a 10 GB input would otherwise cost more than the scan itself
    try:
        total_bytes = sum(os.path.getsize(f) * (4 if f.endswith(".gz") else 1) for f in filenames)
    except OSError:
        total_bytes = 1 << 40
    auto_bases = min(1 << 28, max(1 << 24, 1 << max(0, (total_bytes // 48).bit_length() - 1)))
    if not explicit_devices and len(devices) > 1:
        # a GPU scans ~30 Gbases/s from FASTQ, but bringing one up (CUDA context, device buffers, page-locked
        # staging) and tearing it down costs about a second of process time: measured on an 8-GPU box, 12 GB of
        # FASTQ take 0.14 s to scan and 12 s of wall time when all eight devices are initialised
        # (profiles/README.md).  One more device per 32 GiB of input text; --devices / TOPSICLE_DEVICES override.
        devices = devices[:max(1, min(len(devices), 1 + total_bytes // (32 << 30)))]
    tprint(f"CUDA devices used for the scan: {devices} ({engine.load_library().tps_build_info().decode()})")
    scanner = pipeline.Scanner(scan_configs(args, telo_phrases, patterns, sliding_val), devices=devices,
                               threads=args.threads or 0,
                               max_batch_bases=int(os.environ.get("TOPSICLE_BATCH_BASES", auto_bases)),
                               max_batch_reads=int(os.environ.get("TOPSICLE_BATCH_READS", 1 << 17)),
                               ends_first=bool(getattr(args, "ends_first", False)
                                               or os.environ.get("TOPSICLE_ENDS_FIRST", "0") not in ("", "0")))
    t_created = time.time()
    writers = []
    try:
        if args.read_check:
            tprint("checking specific read:", args.read_check)
        jobs = []
        for seq_loc in filenames:
            w, job = file_job(args, seq_loc, telo_phrases, scanner, sliding_val)
            writers.append(w)
            jobs.append(job)
        # files are scanned concurrently (the reference's Pool is over files, main.py:232-235): one reader
        # thread per file in flight, every GPU takes batches of any file
        for st in scanner.scan_files(jobs):
            total_bases += st.n_bases
            total_reads += st.n_reads
        t_scanned = time.time()
        for w in writers:
            for k, telo_phrase in enumerate(telo_phrases):
                for telolen, trc_val in w.results[k]:
                    phrase_to_telo[telo_phrase].append(telolen)
                    phrase_to_trc[telo_phrase].append(trc_val)
                if k:
                    csv_rows[k].extend(w.rows[k])
    finally:
        for w in writers:
            w.close(create=False)     # files the run never reached leave no empty subset behind
        scanner.close()
    if os.environ.get("TOPSICLE_TIMING"):
        print(f"[timing] contexts + pinned staging {t_created - t0:.3f} s, scan {t_scanned - t_created:.3f} s, "
              f"teardown {time.time() - t_scanned:.3f} s", file=sys.stderr)
    # the CSV is phrase-major (the reference's outer loop is over telo_phrases, main.py:206-235): rows of
    # the first phrase were appended as they were found, those of the other phrases follow here
    with open(output_csv, mode="a", newline="") as file:
        writer = csv.writer(file)
        for rows in csv_rows[1:]:
            writer.writerows(rows)
    dt = max(time.time() - t0, 1e-9)
    tprint("finished processing all reads")
    tprint(f"scanned {total_reads} reads / {total_bases} bases x {len(telo_phrases)} phrase(s) in {dt:.2f} s "
           f"({total_bases * len(telo_phrases) / dt / 1e9:.3f} Gbases/s)")
    print("---------------------")

    inputtrc = args.cutoff[0] if isinstance(args.cutoff, (list, tuple)) else args.cutoff
    for phrase in sorted(phrase_to_telo):
        median_telo = np.median(phrase_to_telo[phrase])
        median_trc = np.median(phrase_to_trc[phrase])
        tprint(f"k-mer: {phrase}, with TRC >= {inputtrc}, median telomere length is {median_telo:.2f} bp")

        if len(phrase_to_telo[phrase]) >= 3:
            max_trc = max(phrase_to_trc[phrase])
            plot_path = os.path.join(args.outputDir, f"quadfit_{phrase}mer_{args.pattern}.png")
            vertex_x, vertex_y, coeffs = fit_quadratic_and_find_vertex(
                phrase_to_trc[phrase], phrase_to_telo[phrase], inputtrc=inputtrc, median_trc=median_trc,
                save_path=plot_path)
            vertex_x, notes = recommended_cutoff(vertex_x, max_trc, median_trc, inputtrc)
            for note in notes:
                tprint(note)
            tprint(f"asymptotic TRC, or recommended cutoff: {vertex_x:.3f}")
            filtered_telolen = [telo for trc, telo in zip(phrase_to_trc[phrase], phrase_to_telo[phrase])
                                if trc >= vertex_x]
            if filtered_telolen:
                tprint(f"Median telomere length for reads with TRC cutoff >= {vertex_x:.3f}: "
                       f"{np.median(filtered_telolen):.2f} bp")
            else:
                tprint(f"No read has TRC >= {vertex_x:.3f}, please double check the data or submit log to GitHub.")
        else:
            tprint("Not enough data points to recommend TRC cutoff.")

    return tprint("All telomere found, have a nice day.")


# The reference's flag table (main.py:319-334) as data: (option strings, metavar, type, default, nargs, help).
# A drop-in keeps names, types, defaults and help sentences to the letter -- `tests/test_cli_host.py::
# test_parser_equals_reference` compares every field and the `--help` text with the reference's own parser
# (tests/golden/cli_parser.json, captured from the unmodified Topsicle.main) -- so the sentences, typos included,
# are the reference's.  type None = a store_true switch; required = the three the reference requires.
_REQ = ("inputDir", "outputDir", "pattern")
_REFERENCE_FLAGS = (
    (("--inputDir", "-i"), "FILE/FOLDER", str, None, None, "Required, Path to the input file or directory"),
    (("--outputDir", "-o"), "FOLDER", str, None, None, "Required, Path to the output directory"),
    (("--pattern",), "CHAR", str, None, None,
     "Required, Telomere repeat sequence (in 5' to 3' orientation). For e.g., in human use CCCTAA"),
    (("--minSeqLength",), "INT", int, 9000, None, "Minimum length of a long read sequence that will be analyzed"),
    (("--rawcountpattern",), None, None, False, None, "Output raw count of the k-mer for each window"),
    (("--telophrase",), "INT", int, None, "+",
     "Length of telomere k-mer to search. By default will use telomere k-mer length minus 2"),
    (("--cutoff",), "FLOAT", float, 0.7, "+", "TRC statistics threshold"),
    (("--windowSize",), "INT", int, 100, None, "Sliding window size"),
    (("--slide",), "INT", int, None, None, "Window sliding step. Default is telomere k-mer length"),
    (("--trimfirst",), "INT", int, 100, None, "Length of intial number of base pairs to trim"),
    (("--maxlengthtelo",), "INT", int, 20000, None, "Longest possible length of telomere for any given read"),
    (("--plot",), None, None, False, None,
     "Optional, generate plot showing for each telomere read the abundance across the sequencing reead and the "
     "changepoint"),
    (("--rangecp",), "INT", int, None, None,
     "Optional, set range of changepoint plot for visualization, default is maxlengthtelo"),
    (("--read_check",), "STR", str, None, None, "Optional, get telomere of a specific read"),
    (("--override", "-ov"), None, None, False, None, "Override telolengths_all.csv file but keep subset fastq"),
    (("--threads", "-t"), "INT", int, None, None, "Number of CPU cores to use (by default, all available cores)"),
)


def build_parser(reference_only: bool = False):
    """The `topsicle` argument parser: the reference's flags, then (unless `reference_only`) the two this build adds."""
    p = argparse.ArgumentParser(prog="topsicle", description="Topsicle - Telomere length estimation from long reads",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for names, metavar, typ, default, nargs, text in _REFERENCE_FLAGS:
        if typ is None:
            p.add_argument(*names, action="store_true", help=text)
            continue
        kw = dict(type=typ, metavar=metavar, help=text)
        if nargs:
            kw["nargs"] = nargs
        if names[0].lstrip("-") in _REQ:
            kw["required"] = True
        else:
            kw["default"] = default
        p.add_argument(*names, **kw)
    if reference_only:
        return p
    p.add_argument("--ends-first", dest="ends_first", action="store_true",
                   help="B200 only, optional: upload just the first/last 1000 bases of every read, then the "
                        "telomere regions of the reads that pass TRC (same outputs; about a tenth of the bytes "
                        "cross PCIe; also TOPSICLE_ENDS_FIRST=1)")
    p.add_argument("--devices", nargs="+", metavar="INT", type=int,
                   help="(B200 build) CUDA devices to scan on; default: all visible devices")
    return p


def main(argv=None):
    start_time = time.time()
    args = build_parser().parse_args(argv)
    tprint.logfile = get_log_path(args)
    analysis_run(args)
    elapsed = time.time() - start_time
    print(f"Elapsed time(s): {elapsed:.2f} seconds")


if __name__ == "__main__":
    main()
