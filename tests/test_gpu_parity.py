"""GPU parity tests: CUDA path (through the C-ABI / ctypes) vs the oracle and the golden
vectors of the unmodified reference.  Integer results are required to be bit-exact."""
import argparse
import hashlib
import os

import numpy as np
import pytest

from oracle import topsicle_oracle as orc
from tests.conftest import GOLD, load_json
from tests.test_gpu_random_sweep import make_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from topsicle_b200 import engine
    engine.load_library()
    return engine


def _ctx(eng, patterns, **kw):
    kw.setdefault("max_batch_reads", 4096)
    kw.setdefault("max_batch_bases", 1 << 24)
    return eng.ScanContext(patterns, **kw)


# ------------------------------------------------------------------------------- K1
def test_k1_pack_matches_numpy(eng):
    rng = np.random.default_rng(11)
    n = 512 * 37 + 123
    alphabet = np.frombuffer(b"ACGTacgtNnRY\n@+-*", dtype=np.uint8)
    bases = alphabet[rng.integers(0, 8, n)]
    hit = rng.random(n) < 0.01
    bases = np.where(hit, alphabet[rng.integers(0, len(alphabet), n)], bases).astype(np.uint8)
    bases[5000:5200] = rng.integers(0, 256, 200).astype(np.uint8)
    offsets = np.array([0, n], dtype=np.uint64)
    with _ctx(eng, ["CCCTA"], min_seq_length=0) as ctx:
        ctx.scan(bases, offsets)
        ng = (n + 15) // 16
        codes = ctx.debug_copy(0, ng * 4).view(np.uint32)
        flags = ctx.debug_copy(1, ((n + 511) // 512) * 4).view(np.uint32)
        masks = ctx.debug_copy(2, ng * 2).view(np.uint16)
    pad = np.concatenate([bases, np.full(ng * 16 - n, ord("N"), np.uint8)]).reshape(ng, 16)
    valid = np.isin(pad, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    code = (pad >> 1) & 3
    g = np.arange(16)
    shift = (8 * (g & 3) + 2 * (g >> 2)).astype(np.uint32)
    want = (code.astype(np.uint32) << shift).sum(axis=1).astype(np.uint32)
    full = n // 16  # groups entirely inside the uploaded bytes
    assert np.array_equal(codes[:full], want[:full])
    want_flag = ~valid.all(axis=1)
    got_flag = ((flags[np.arange(ng) >> 5] >> (np.arange(ng) & 31).astype(np.uint32)) & 1).astype(bool)
    assert np.array_equal(got_flag[:full], want_flag[:full])
    want_mask = (valid.astype(np.uint32) << g.astype(np.uint32)).sum(axis=1).astype(np.uint16)
    assert want_flag[:full].sum() > 50
    assert np.array_equal(masks[:full], want_mask[:full])


# ------------------------------------------------------------------------------- step 1
def _check_step1(eng, records, motif, k, minlen, no_bp):
    pats = orc.patterns_to_search(motif, k)
    with _ctx(eng, pats, len_telopattern=len(motif), min_seq_length=minlen, no_bp=no_bp,
              count_threshold_override=0, maxlengthtelo=3000) as ctx:
        rows, _ = ctx.scan_reads([s for _, s in records])
    assert len(rows) == len(records)
    for (rid, seq), row in zip(records, rows):
        assert row["length"] == len(seq)
        if len(seq) <= minlen:
            assert row["status"] == eng.ST_FILTERED
            continue
        tail, bi, cnt, ms, me, _, _ = orc.trc_read(seq, pats, len(motif), no_bp)
        got = (eng.TAIL_NAMES[row["tail"]], int(row["best_pattern"]), int(row["match_count"]),
               int(row["head_max"]), int(row["tail_max"]))
        assert got == (tail, bi, cnt, ms, me), (rid, motif, k)
        assert row["status"] in (eng.ST_PASS, eng.ST_BADSEG)


@pytest.mark.parametrize("motif,k", [("CCCTAA", 4), ("CCCTAA", 5), ("CCCTAA", 6), ("CCCTAAA", 5), ("AAACCCT", 5),
                                     ("TTAGGG", 2), ("TTAGGG", 3), ("TTTAGGG", 7), ("CCCTAAA", 7)])
def test_step1_edge_and_demo(eng, edge_records, demo_records, motif, k):
    _check_step1(eng, edge_records, motif, k, 0, 1000)
    _check_step1(eng, demo_records, motif, k, 9000, 1000)


def test_step1_other_no_bp(eng, edge_records):
    _check_step1(eng, edge_records, "CCCTAA", 4, 0, 300)
    _check_step1(eng, edge_records, "CCCTAAA", 5, 999, 2000)
    _check_step1(eng, edge_records, "CCCTAA", 5, 2400, 1000)


def test_step1_golden_rows(eng, demo_records):
    """Against rows produced by the unmodified reference's patternTRC_count."""
    for e in load_json("demo_step1.json"):
        pats = orc.patterns_to_search(e["pattern"], e["kmer"])
        with _ctx(eng, pats, len_telopattern=len(e["pattern"]), min_seq_length=e["read_length"],
                  count_threshold_override=0) as ctx:
            rows, _ = ctx.scan_reads([s for _, s in demo_records])
        got = []
        for (rid, _), row in zip(demo_records, rows):
            if row["status"] == eng.ST_FILTERED:
                continue
            trc = eng.trc_value(row["match_count"], len(e["pattern"]))
            got.append([rid, pats[row["best_pattern"]], eng.TAIL_NAMES[row["tail"]], repr(float(trc))])
        assert got == e["rows"], (e["pattern"], e["kmer"])


# ------------------------------------------------------------------------------- step 2/3
CFGS = [("CCCTAA", 4, 100, 6, 100, 20000), ("CCCTAA", 5, 100, 6, 100, 2000), ("CCCTAA", 6, 50, 3, 0, 1800),
        ("CCCTAAA", 5, 100, 7, 100, 20000), ("TTAGGG", 3, 100, 6, 50, 2500), ("TTAGGG", 2, 30, 1, 10, 400),
        ("TTTAGGG", 7, 64, 5, 33, 1500), ("AAACCCT", 5, 100, 7, 200, 20000), ("CCCTAA", 4, 300, 1, 0, 9000)]


@pytest.mark.parametrize("cfg", CFGS)
def test_step2_counts_and_changepoint(eng, edge_records, demo_records, cfg):
    motif, k, W, s, t, M = cfg
    pats = orc.patterns_to_search(motif, k)
    records = edge_records + [(rid + "_rev", seq[::-1]) for rid, seq in edge_records] + demo_records[:12]
    with _ctx(eng, pats, len_telopattern=len(motif), min_seq_length=0, count_threshold_override=0,
              window_size=W, slide=s, trimfirst=t, maxlengthtelo=M, want_rawcount=True,
              rawcount_capacity=1 << 27) as ctx:
        rows, raw = ctx.scan_reads([sq for _, sq in records])
        n_pass = n_bad = 0
        for i, ((rid, seq), row) in enumerate(zip(records, rows)):
            if len(seq) == 0:
                assert row["status"] == eng.ST_FILTERED
                continue
            tail = eng.TAIL_NAMES[row["tail"]]
            assert tail == orc.trc_read(seq, pats, len(motif))[0]
            region = orc.oriented_region(seq, tail, t, M)
            counts = orc.window_counts(region, pats, W, s)
            assert row["n_windows"] == counts.shape[0], rid
            tab = ctx.rawcount_table(rows, raw, i)
            if counts.shape[0]:
                assert np.array_equal(tab.astype(np.int64), counts), rid
            if counts.shape[0] < 7:
                assert row["status"] == eng.ST_BADSEG and row["telo_length"] == -1
                n_bad += 1
                continue
            bkp = orc.change_point_exact(counts.sum(axis=1))
            assert row["status"] == eng.ST_PASS
            assert (int(row["bkp"]), int(row["telo_length"])) == (bkp, t + s * bkp), rid
            n_pass += 1
    assert n_pass > 20


def test_step2_golden_edge(eng, edge_records):
    """telo_length vs the unmodified reference's bound_detect (float64 ruptures): must match
    except on exactly tied gains (constant signals), where the reference's answer is rounding noise."""
    seqs = dict(edge_records)
    gold = load_json("edge.json")["step2"]
    by_cfg = {}
    for e in gold:
        by_cfg.setdefault((e["motif"], e["k"], e["W"], e["slide"], e["trimfirst"], e["maxlengthtelo"]), []).append(e)
    n_eq = n_tie = 0
    for (motif, k, W, s, t, M), entries in by_cfg.items():
        pats = orc.patterns_to_search(motif, k)
        with _ctx(eng, pats, len_telopattern=len(motif), min_seq_length=0, count_threshold_override=0,
                  window_size=W, slide=s, trimfirst=t, maxlengthtelo=M) as ctx:
            ids = list(seqs)
            rows, _ = ctx.scan_reads([seqs[i] for i in ids])
        rowof = dict(zip(ids, rows))
        for e in entries:
            row = rowof[e["read"]]
            if eng.TAIL_NAMES[row["tail"]] != e["tail"]:
                continue
            assert row["n_windows"] == e["n_windows"]
            if "telo_length" not in e or e["telo_length"] is None:
                assert row["status"] == eng.ST_BADSEG
                continue
            if int(row["telo_length"]) == e["telo_length"]:
                n_eq += 1
            else:
                assert len(set(e["c_w"])) == 1, e["read"]  # constant signal: every gain is exactly 0
                n_tie += 1
    assert n_eq > 200 and n_tie <= 9


# ------------------------------------------------------------------------------- whole path
def _cli_args(argv):
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


def test_demo_csv_byte_identical(eng, demo_records):
    gold = load_json("demo_cli.json")
    for case in gold["cases"]:
        a = _cli_args(case["argv"])
        phrases = a.telophrase or [len(a.pattern) - 2]
        cutoff = min(a.cutoff) if isinstance(a.cutoff, list) else a.cutoff
        slide = a.slide or len(a.pattern)
        text = "file_number,phrase,trc,readID,telo_length\r\n"
        for k in phrases:
            pats = orc.patterns_to_search(a.pattern, k)
            with _ctx(eng, pats, len_telopattern=len(a.pattern), cutoff=cutoff, min_seq_length=a.minSeqLength,
                      window_size=a.windowSize, slide=slide, trimfirst=a.trimfirst,
                      maxlengthtelo=a.maxlengthtelo) as ctx:
                rows, _ = ctx.scan_reads([s for _, s in demo_records])
            for (rid, _), row in zip(demo_records, rows):
                if row["status"] == eng.ST_PASS:
                    trc = eng.trc_value(row["match_count"], len(a.pattern))
                    text += f"demo.fastq,{k},{trc:.3f},{rid},{row['telo_length']}\r\n"
                assert row["status"] != eng.ST_BADSEG
        assert text == case["csv"], case["name"]
        assert hashlib.md5(text.encode()).hexdigest() == case["csv_md5"]


def test_demo_rawcount_tables(eng, demo_records):
    seqs = dict(demo_records)
    tabs = np.load(os.path.join(GOLD, "demo_rawcount.npz"))
    for m in load_json("demo_rawcount.json"):
        with _ctx(eng, m["patterns"], len_telopattern=len(m["pattern"]), min_seq_length=9000, cutoff=0.5,
                  window_size=m["windowSize"], slide=m["slide"], trimfirst=m["trimfirst"],
                  maxlengthtelo=m["maxlengthtelo"], want_rawcount=True, rawcount_capacity=1 << 22) as ctx:
            rows, raw = ctx.scan_reads([seqs[m["read"]]])
            assert eng.TAIL_NAMES[rows[0]["tail"]] == m["tail"]
            assert np.array_equal(ctx.rawcount_table(rows, raw, 0), tabs[m["key"]].astype(np.uint8))
            assert rows[0]["telo_length"] == m["telo_length"]


def test_random_reads_property(eng):
    """Seeded synthetic reads (telomeric / near-threshold / N / lower-case), many read lengths
    and alignments: every integer output equals the oracle's."""
    rng = np.random.default_rng(2026)
    B = np.array(list("ACGT"))
    reads = []
    for i in range(300):
        L = int(rng.integers(1, 6000))
        s = B[rng.integers(0, 4, L)]
        kind = i % 5
        motif = "CCCTAA"
        if kind in (0, 1) and L > 200:
            tl = int(rng.integers(50, max(51, L - 50)))
            tel = np.array(list((motif * (tl // 6 + 2))[int(rng.integers(0, 6)):][:tl]))
            err = rng.random(tl) < 0.04
            tel[err] = B[rng.integers(0, 4, int(err.sum()))]
            if kind == 0:
                s[:tl] = tel
            else:
                rc = np.array(list("".join(tel)[::-1].translate(str.maketrans("ACGT", "TGCA"))))
                s[L - tl:] = rc
        if kind == 2:
            s[rng.random(L) < 0.02] = "N"
        seq = "".join(s)
        if kind == 3 and L > 600:
            seq = seq[:100] + seq[100:600].lower() + seq[600:]
        reads.append(seq)
    for motif, k, W, sl, t, M in [("CCCTAA", 4, 100, 6, 100, 20000), ("CCCTAA", 5, 50, 3, 17, 2500)]:
        pats = orc.patterns_to_search(motif, k)
        with _ctx(eng, pats, len_telopattern=6, min_seq_length=150, cutoff=0.3, window_size=W, slide=sl,
                  trimfirst=t, maxlengthtelo=M) as ctx:
            rows, _ = ctx.scan_reads(reads)
        thr = eng.count_threshold(0.3, 6)
        npass = 0
        for seq, row in zip(reads, rows):
            if len(seq) <= 150:
                assert row["status"] == eng.ST_FILTERED
                continue
            tail, bi, cnt, ms, me, _, _ = orc.trc_read(seq, pats, 6)
            assert (eng.TAIL_NAMES[row["tail"]], row["best_pattern"], row["match_count"]) == (tail, bi, cnt)
            assert (orc.trc_value(cnt, 6) > 0.3) == (cnt >= thr)
            if cnt < thr:
                assert row["status"] == eng.ST_BELOW
                continue
            region = orc.oriented_region(seq, tail, t, M)
            c_w = orc.window_counts(region, pats, W, sl).sum(axis=1)
            assert row["n_windows"] == len(c_w)
            if len(c_w) < 7:
                assert row["status"] == eng.ST_BADSEG
            else:
                assert row["telo_length"] == t + sl * orc.change_point_exact(c_w)
                npass += 1
        assert npass > 30


def test_empty_and_capacity(eng):
    with _ctx(eng, ["CCCTA"], max_batch_reads=4, max_batch_bases=4096) as ctx:
        rows, _ = ctx.scan(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
        assert len(rows) == 0
        rows, _ = ctx.scan_reads(["", "ACGT", ""])
        assert list(rows["length"]) == [0, 4, 0]
        with pytest.raises(eng.TpsError):
            ctx.scan_reads(["A"] * 5)
        with pytest.raises(eng.TpsError):
            ctx.scan_reads(["A" * 5000])
    with pytest.raises(eng.TpsError):
        _ctx(eng, ["CCNTA"])
    with pytest.raises(eng.TpsError):
        _ctx(eng, ["CC|TA"])


def test_generic_literal_lists(eng, edge_records, demo_records):
    """Literal lists of mixed lengths / longer than 8 (the reference accepts a list for pattern_telo,
    allsteps.py:122-123) take the generic (non-templated) matching path."""
    lists = [["CCCTAA", "TTAGGGTTA", "AA", "CCCTAACCCTAA", "G"], ["CCCTAAACCCTAAACCCTAAACCCTAAACCC", "AAACCCT"],
             ["CCCTAAACC", "GGGATTTGG", "TTTAGGGTT"]]
    records = edge_records + demo_records[:8]
    for pats in lists:
        W, s, t, M = 100, 6, 50, 4000
        with _ctx(eng, pats, len_telopattern=7, min_seq_length=0, count_threshold_override=0, window_size=W,
                  slide=s, trimfirst=t, maxlengthtelo=M, want_rawcount=True, rawcount_capacity=1 << 26) as ctx:
            rows, raw = ctx.scan_reads([sq for _, sq in records])
            for i, ((rid, seq), row) in enumerate(zip(records, rows)):
                tail, bi, cnt, ms, me, _, _ = orc.trc_read(seq, pats, 7)
                assert (eng.TAIL_NAMES[row["tail"]], int(row["best_pattern"]), int(row["match_count"]),
                        int(row["head_max"]), int(row["tail_max"])) == (tail, bi, cnt, ms, me), (rid, pats)
                counts = orc.window_counts(orc.oriented_region(seq, tail, t, M), pats, W, s)
                assert row["n_windows"] == counts.shape[0]
                if counts.shape[0]:
                    assert np.array_equal(ctx.rawcount_table(rows, raw, i).astype(np.int64), counts), rid
                if counts.shape[0] >= 7:
                    assert row["telo_length"] == t + s * orc.change_point_exact(counts.sum(axis=1))


def test_max_pass_overflow_is_loud(eng, demo_records):
    pats = orc.patterns_to_search("CCCTAAA", 5)
    with _ctx(eng, pats, len_telopattern=7, cutoff=0.7, slide=6, max_pass_reads=4) as ctx:
        with pytest.raises(eng.TpsError):
            ctx.scan_reads([s for _, s in demo_records])
    with _ctx(eng, pats, len_telopattern=7, cutoff=0.7, slide=6, max_pass_reads=17) as ctx:
        rows, _ = ctx.scan_reads([s for _, s in demo_records])
        assert int((rows["status"] == eng.ST_PASS).sum()) == 17


def test_span_batch_equals_back_to_back(eng, edge_records, demo_records):
    """tps_submit_spans: reads separated by gaps of junk give the same rows as the packed batch, also for a
    second context sharing the upload (tps_submit_shared)."""
    rng = np.random.default_rng(4)
    records = edge_records + demo_records[:10]
    seqs = [s.encode() for _, s in records]
    starts, lens, chunks, at = [], [], [], 0
    for s in seqs:
        gap = bytes(rng.choice(np.frombuffer(b"ACGTN@+\n", np.uint8), int(rng.integers(0, 40))))
        chunks.append(gap)
        at += len(gap)
        starts.append(at)
        lens.append(len(s))
        chunks.append(s)
        at += len(s)
    chunks.append(b"CCCTAA" * 7)
    buf = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    starts = np.array(starts, dtype=np.uint64)
    lens = np.array(lens, dtype=np.uint32)
    p4, p5 = orc.patterns_to_search("CCCTAA", 4), orc.patterns_to_search("CCCTAA", 5)
    kw = dict(len_telopattern=6, min_seq_length=0, cutoff=0.3, slide=6, want_rawcount=True, rawcount_capacity=1 << 26)
    with _ctx(eng, p4, **kw) as a, _ctx(eng, p5, max_batch_bases=1, **kw) as b:
        want4 = a.scan_reads(seqs)
        bid = a.submit_spans(buf, starts, lens, len(seqs))
        b.submit_shared(a, bid)
        got4, got5 = a.wait(bid), b.wait(bid)
        with _ctx(eng, p5, **kw) as c:
            want5 = c.scan_reads(seqs)
    fields = [f for f in eng.ROW_DTYPE.names if f != "rawcount_offset"]   # offsets are handed out by atomics
    for ctx_p, got, want in ((p4, got4, want4), (p5, got5, want5)):
        for f in fields:
            assert np.array_equal(got[0][f], want[0][f]), f
        for i in np.nonzero(want[0]["n_windows"] > 0)[0]:
            nw, npat = int(want[0]["n_windows"][i]), len(ctx_p)
            ga, wa = int(got[0]["rawcount_offset"][i]), int(want[0]["rawcount_offset"][i])
            assert np.array_equal(got[1][ga:ga + nw * npat], want[1][wa:wa + nw * npat]), i
    assert int((got4[0]["status"] == eng.ST_PASS).sum()) > 20
    with _ctx(eng, p4, **kw) as a:
        with pytest.raises(eng.TpsError):
            a.submit_spans(buf[:100], starts, lens, len(seqs))     # last read ends beyond the uploaded bytes


# ------------------------------------------------------------- kernel / stream variants
VARIANTS = [
    {},                                                     # bulk-copy K1, pack / tail streams (defaults)
    {"TPS_K1_TMA": "0"},                                    # register-staged K1
    {"TPS_SPLIT_STREAMS": "0"},                             # whole batch on its slot's stream
    {"TPS_K1_STAGES": "2", "TPS_K1_STAGE_KB": "8", "TPS_K1_CTAS_PER_SM": "1"},
    {"TPS_K1_STAGES": "6", "TPS_K1_CTAS_PER_SM": "3"},
    {"TPS_K2_SMEM_PATH": "1"},                              # shared-memory staged K2
    {"TPS_K2_NO_PAIRS": "1"},                               # literal-by-literal match instead of complement pairs
]


def test_kernel_and_stream_variants_agree(eng, monkeypatch):
    """Every tuning knob (read at tps_create) must give bit-identical codes, flags, rows and raw counts:
    the knobs select how the bytes move, never what is computed."""
    rng = np.random.default_rng(77)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    lens = rng.integers(1, 9000, 700)
    lens[::50] = rng.integers(16384 * 3, 16384 * 5, len(lens[::50]))  # several 16-KiB chunks per read
    offsets = np.zeros(len(lens) + 1, np.uint64)
    offsets[1:] = np.cumsum(lens)
    n = int(offsets[-1])
    bases = B[rng.integers(0, 4, n)].copy()
    tel = np.frombuffer((b"CCCTAA" * (n // 6 + 1))[:n], dtype=np.uint8)
    for i in range(0, len(lens), 7):                        # telomeric heads, some with errors
        a, e = int(offsets[i]), int(offsets[i]) + min(int(lens[i]), 3000)
        bases[a:e] = tel[:e - a]
    bases[rng.random(n) < 0.002] = ord("N")
    bases[rng.random(n) < 0.0005] = ord("r")
    pats = orc.patterns_to_search("CCCTAA", 5)              # two self-overlapping literals
    ref = None
    for env in VARIANTS:
        for k in ("TPS_K1_TMA", "TPS_SPLIT_STREAMS", "TPS_K1_STAGES", "TPS_K1_STAGE_KB", "TPS_K1_CTAS_PER_SM",
                  "TPS_K2_SMEM_PATH", "TPS_K2_NO_PAIRS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with _ctx(eng, pats, len_telopattern=6, min_seq_length=500, cutoff=0.4, slide=6, want_rawcount=True,
                  rawcount_capacity=1 << 26, max_batch_bases=n + 4096, n_slots=3) as ctx:
            bids = [ctx.submit(bases, offsets) for _ in range(3)]      # three batches in flight on three slots
            got = [_rows_without_offsets(eng, ctx.wait(b)[0]) for b in bids]
            rows, raw = ctx.scan(bases, offsets)
            ng = n // 16
            codes = ctx.debug_copy(0, ng * 4).tobytes()
            flags = ctx.debug_copy(1, (n // 512) * 4).tobytes()
            passing = np.nonzero(rows["status"] == eng.ST_PASS)[0]
            # raw-count offsets come from an atomic cursor: compare the tables per read, not the layout
            tables = [ctx.rawcount_table(rows, raw, int(i)).tobytes() for i in passing[:40]]
        cur = (codes, flags, _rows_without_offsets(eng, rows), tables)
        assert got[0] == got[1] == got[2] == cur[2], env
        if ref is None:
            ref = cur
            assert len(passing) > 20
        else:
            for a_, b_ in zip(cur, ref):
                assert a_ == b_, env


@pytest.mark.parametrize("motif,k", [("CCCTAA", 4), ("TTAGGG", 4), ("CCCTAA", 3), ("CCCTAAA", 5), ("TTTAGGG", 5),
                                     ("AAACCCT", 5), ("TTTAGGG", 6), ("TTTTAGGG", 6), ("CCCTAAAA", 8), ("CCTAA", 3)])
def test_k2_with_literals_in_the_instructions_equals_table_k2(eng, edge_records, demo_records, monkeypatch, motif, k):
    """Complement-paired literal sets of 5..8 literals without self-overlap run tps_trc_const_kernel<K, U> (masks as
    constant operands, unrolled); TPS_K2_CONST=0 keeps tps_trc_reg_kernel<K>.  Same rows, whole reads and ends
    batches, and both equal the oracle's step 1 (allsteps.py:175-198)."""
    pats = orc.patterns_to_search(motif, k)
    rng = np.random.default_rng(len(motif) * 31 + k)
    reads = [sq for _, sq in edge_records] + [sq for _, sq in demo_records[:30]] + make_reads(rng, motif, 120, 20000)
    out = []
    for const in (True, False):
        if const:
            monkeypatch.delenv("TPS_K2_CONST", raising=False)
        else:
            monkeypatch.setenv("TPS_K2_CONST", "0")
        with _ctx(eng, pats, len_telopattern=len(motif), min_seq_length=0, cutoff=0.3, slide=len(motif)) as ctx:
            bordered = any(p[:i] == p[-i:] for p in pats for i in range(1, len(p)))
            eligible = not bordered and 5 <= len(pats) // 2 <= 8 and 3 <= k <= 8
            assert ctx.debug_info()["k2_kernel"] == ("const" if const and eligible else "reg")
            rows, _ = ctx.scan_reads(reads)
            out.append(rows)
    assert out[0].tobytes() == out[1].tobytes()
    n_pass = 0
    for i, seq in enumerate(reads):
        if len(seq) == 0:
            continue
        tail, bi, cnt, ms, me, _, _ = orc.trc_read(seq, pats, len(motif), 1000)
        row = out[0][i]
        got = (eng.TAIL_NAMES[row["tail"]], int(row["best_pattern"]), int(row["match_count"]), int(row["head_max"]),
               int(row["tail_max"]))
        assert got == (tail, bi, cnt, ms, me), (motif, k, i)
        n_pass += row["status"] >= eng.ST_PASS
    assert n_pass > 10


def _rows_without_offsets(eng, rows):
    r = rows.copy()
    r["rawcount_offset"] = 0
    return r.tobytes()


# ------------------------------------------------------------------------- ends-first protocol
def _random_reads(seed, n=260, lmax=30000):
    rng = np.random.default_rng(seed)
    B = np.array(list("ACGT"))
    reads = []
    for i in range(n):
        L = int(rng.integers(1, lmax if i % 4 else 2500))   # many reads shorter than 2 * no_bp
        s = B[rng.integers(0, 4, L)]
        if i % 3 == 0 and L > 300:
            tl = int(rng.integers(100, L))
            tel = np.array(list(("CCCTAA" * (tl // 6 + 2))[int(rng.integers(0, 6)):][:tl]))
            err = rng.random(tl) < 0.03
            tel[err] = B[rng.integers(0, 4, int(err.sum()))]
            if i % 2:
                s[:tl] = tel
            else:
                s[L - tl:] = np.array(list("".join(tel)[::-1].translate(str.maketrans("ACGT", "TGCA"))))
        if i % 7 == 0:
            s[rng.random(L) < 0.01] = "N"
        reads.append("".join(s))
    return reads


@pytest.mark.parametrize("cfg", [
    dict(motif="CCCTAA", k=4, W=100, slide=6, trim=100, maxlen=20000, minlen=500, cutoff=0.3, no_bp=1000),
    dict(motif="CCCTAA", k=5, W=50, slide=3, trim=17, maxlen=2500, minlen=0, cutoff=0.2, no_bp=1000),
    dict(motif="CCCTAAA", k=5, W=100, slide=7, trim=100, maxlen=700, minlen=9000, cutoff=0.4, no_bp=1000),
    dict(motif="CCCTAA", k=6, W=80, slide=6, trim=0, maxlen=20000, minlen=100, cutoff=0.1, no_bp=300),
])
def test_ends_first_equals_whole_read_scan(eng, demo_records, cfg):
    """tps_submit_ends + tps_submit_regions (head/tail upload, then the regions of the TRC-pass reads) give the
    rows and raw-count tables of the whole-read scan, field by field."""
    reads = _random_reads(31) + [s for _, s in demo_records]
    pats = orc.patterns_to_search(cfg["motif"], cfg["k"])
    with _ctx(eng, pats, len_telopattern=len(cfg["motif"]), min_seq_length=cfg["minlen"], cutoff=cfg["cutoff"],
              window_size=cfg["W"], slide=cfg["slide"], trimfirst=cfg["trim"], maxlengthtelo=cfg["maxlen"],
              no_bp=cfg["no_bp"], want_rawcount=True, rawcount_capacity=1 << 27, max_batch_bases=1 << 25) as ctx:
        full, raw = ctx.scan_reads(reads)
        got, tables = ctx.scan_reads_ends_first(reads)
        assert (full["status"] >= eng.ST_PASS).sum() > 15
        for f in ("length", "status", "tail", "best_pattern", "match_count", "head_max", "tail_max", "n_windows",
                  "bkp", "telo_length"):
            assert np.array_equal(full[f], got[f]), f
        for i in np.nonzero(full["status"] >= eng.ST_PASS)[0]:
            want = ctx.rawcount_table(full, raw, int(i))
            if want is None:
                assert tables[i] is None or tables[i].size == 0
            else:
                assert np.array_equal(want, tables[i]), i
