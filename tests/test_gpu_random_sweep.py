"""GPU tier: seeded random sweep over motifs, telophrases, window geometry, orientation forcing and
read content -- every integer the scan returns equals the oracle's (which is pinned to the reference)."""
import numpy as np
import pytest

from oracle import topsicle_oracle as orc

pytestmark = pytest.mark.gpu

MOTIFS = ["CCCTAA", "TTAGGG", "CCCTAAA", "TTTAGGG", "AAACCCT", "TTAGG", "TTTTAGGG", "CCCTAAAA", "AACCT", "TTGGGG",
          "ACACAC", "AAAAAA", "CCCGAA"]


def make_reads(rng, motif, n_reads, max_len):
    B = np.array(list("ACGT"))
    rc = str.maketrans("ACGT", "TGCA")
    reads = []
    for i in range(n_reads):
        L = int(rng.integers(1, max_len))
        s = B[rng.integers(0, 4, L)]
        kind = int(rng.integers(0, 6))
        if kind in (0, 1, 2) and L > 60:
            tl = int(rng.integers(20, L))
            ph = int(rng.integers(0, len(motif)))
            tel = np.array(list((motif * (tl // len(motif) + 2))[ph:ph + tl]))
            err = rng.random(tl) < rng.choice([0.0, 0.02, 0.1])
            tel[err] = B[rng.integers(0, 4, int(err.sum()))]
            if kind == 0:
                s[:tl] = tel
            elif kind == 1:
                s[L - tl:] = np.array(list("".join(tel)[::-1].translate(rc)))
            else:                                   # telomere-like repeats at BOTH ends: exercises the tie rules
                s[:tl] = tel
                s[L - tl:] = np.array(list("".join(tel)[::-1]))
        if kind == 3:
            s[rng.random(L) < 0.05] = rng.choice(list("NnRYKM-*"))
        seq = "".join(s)
        if kind == 4 and L > 10:
            a = int(rng.integers(0, L - 5))
            seq = seq[:a] + seq[a:a + L // 2].lower() + seq[a + L // 2:]
        reads.append(seq)
    return reads


@pytest.mark.parametrize("seed", range(60))
def test_random_configuration(seed):
    from topsicle_b200 import engine
    rng = np.random.default_rng(1000 + seed)
    motif = MOTIFS[seed % len(MOTIFS)]
    k = int(rng.integers(2, len(motif) + 1))
    pats = orc.patterns_to_search(motif, k)
    W = int(rng.choice([20, 50, 64, 100, 101, 150, 333]))
    s = int(rng.choice([1, 2, 3, 5, 6, 7, 13, 40]))
    t = int(rng.choice([0, 1, 17, 100, 200]))
    M = int(rng.choice([300, 1000, 2500, 5000, 20000]))
    no_bp = int(rng.choice([1000, 1000, 1000, 100, 640, 999]))
    minlen = int(rng.choice([0, 0, 50, 400]))
    cutoff = float(rng.choice([0.0, 0.1, 0.3, 0.7]))
    force = [None, None, None, "forward", "reverse"][int(rng.integers(0, 5))]
    reads = make_reads(rng, motif, 70, 7000) + ["", "A", motif, motif * 200, (motif * 200)[::-1]]
    thr = engine.count_threshold(cutoff, len(motif), no_bp)
    with engine.ScanContext(pats, len_telopattern=len(motif), cutoff=cutoff, min_seq_length=minlen, no_bp=no_bp,
                            window_size=W, slide=s, trimfirst=t, maxlengthtelo=M, want_rawcount=True,
                            rawcount_capacity=1 << 27, max_batch_reads=256, max_batch_bases=1 << 20,
                            force_tail=force) as ctx:
        rows, raw = ctx.scan_reads(reads)
        n_checked = 0
        for i, (seq, row) in enumerate(zip(reads, rows)):
            assert row["length"] == len(seq)
            if len(seq) <= minlen:
                assert row["status"] == engine.ST_FILTERED
                continue
            tail, bi, cnt, ms, me, hc, tc = orc.trc_read(seq, pats, len(motif), no_bp)
            if force:
                tail = force
                cs = hc if force == "forward" else tc
                cnt = max(cs)
                bi = cs.index(cnt)
            assert (engine.TAIL_NAMES[row["tail"]], int(row["best_pattern"]), int(row["match_count"]),
                    int(row["head_max"]), int(row["tail_max"])) == (tail, bi, cnt, ms, me), (seed, i)
            assert (orc.trc_value(cnt, len(motif), no_bp) > cutoff) == (cnt >= thr)
            if cnt < thr:
                assert row["status"] == engine.ST_BELOW
                continue
            counts = orc.window_counts(orc.oriented_region(seq, tail, t, M), pats, W, s)
            assert row["n_windows"] == counts.shape[0], (seed, i)
            if counts.shape[0]:
                assert np.array_equal(ctx.rawcount_table(rows, raw, i).astype(np.int64), counts), (seed, i)
            if counts.shape[0] < 7:
                assert row["status"] == engine.ST_BADSEG and row["telo_length"] == -1
            else:
                b = orc.change_point_exact(counts.sum(axis=1))
                assert row["status"] == engine.ST_PASS
                assert (int(row["bkp"]), int(row["telo_length"])) == (b, t + s * b), (seed, i)
            n_checked += 1
        assert n_checked > 10 or cutoff >= 0.3
