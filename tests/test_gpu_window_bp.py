"""GPU tier: the bit-parallel window kernel (tps_window_bp_kernel, the default K3 with the change point fused;
a read per CTA, or -- few passing reads -- a read's tiles dealt over 2 or 4 CTAs, TPS_K3_SPLIT=0 to forbid)
against the oracle and against the plain per-literal kernel (tps_window_kernel + tps_changepoint_kernel,
TPS_K3_BITPAR=0): the sums of c_w over the groups of five windows (what the change point reads), n_windows,
status, bkp and telo_length must be bit-identical; TPS_K3_NO_GROUPS=1 (every window on its own) likewise.
Reference semantics: /root/reference/Topsicle/allsteps.py:207-225, 279-291 (W-1 text, `or 1`, sum over all P),
:304-315 + ruptures 1.1.9 Binseg."""
import numpy as np
import pytest

from oracle import topsicle_oracle as orc
from tests.test_gpu_random_sweep import MOTIFS, make_reads

pytestmark = pytest.mark.gpu

# (motif, k, W, slide, trimfirst, maxlengthtelo): D = W - k >= 32 selects the bit-parallel kernel; phrases 5/6 of
# CCCTAA and 6/7 of the 7-mers have self-overlapping literals; W 100 / 50 are the BASELINE configs' geometries,
# 36 + k is the smallest eligible window (D = 32: a = 1, r = 0), 333 / 1000 give a = 10 / 31 dilation words
CFGS = [("CCCTAA", 4, 100, 6, 100, 20000), ("TTAGGG", 4, 50, 3, 100, 20000), ("CCCTAA", 5, 100, 6, 100, 20000),
        ("CCCTAA", 6, 100, 6, 100, 20000), ("CCCTAA", 6, 50, 3, 0, 1800), ("CCCTAAA", 5, 100, 7, 100, 20000),
        ("TTTAGGG", 7, 64, 5, 33, 1500), ("AAACCCT", 6, 101, 7, 200, 20000), ("CCCTAA", 4, 36, 1, 0, 9000),
        ("CCCTAA", 4, 300, 1, 0, 9000), ("TTAGGG", 3, 333, 13, 17, 5000), ("TTAGGG", 2, 1000, 40, 1, 20000),
        ("ACACAC", 4, 100, 2, 0, 3000), ("AAAAAA", 3, 64, 1, 5, 2500), ("TTTTAGGG", 8, 100, 8, 100, 20000)]


def _scan(engine, pats, motif, reads, W, s, t, M, bitpar, monkeypatch, no_groups=False, no_split=False, **kw):
    if bitpar:
        monkeypatch.delenv("TPS_K3_BITPAR", raising=False)
    else:
        monkeypatch.setenv("TPS_K3_BITPAR", "0")
    if no_groups:
        monkeypatch.setenv("TPS_K3_NO_GROUPS", "1")
    else:
        monkeypatch.delenv("TPS_K3_NO_GROUPS", raising=False)
    if no_split:
        monkeypatch.setenv("TPS_K3_SPLIT", "0")
    else:
        monkeypatch.delenv("TPS_K3_SPLIT", raising=False)
    monkeypatch.setenv("TPS_K3_DEBUG_GS", "1")     # the bit-parallel kernel also writes its group sums out
    with engine.ScanContext(pats, len_telopattern=len(motif), min_seq_length=0, count_threshold_override=0,
                            window_size=W, slide=s, trimfirst=t, maxlengthtelo=M, max_batch_reads=1024,
                            max_batch_bases=1 << 24, **kw) as ctx:
        assert ctx.debug_info()["k3_bitpar"] == bitpar
        launches0 = ctx.kernel_launches()
        rows, _ = ctx.scan_reads(reads)
        return rows, ctx.window_sums(rows), ctx.kernel_launches() - launches0


@pytest.mark.parametrize("cfg", CFGS)
def test_window_sums_equal_oracle_and_plain_kernel(cfg, edge_records, demo_records, monkeypatch):
    from topsicle_b200 import engine
    motif, k, W, s, t, M = cfg
    pats = orc.patterns_to_search(motif, k)
    rng = np.random.default_rng(900 + CFGS.index(cfg))
    reads = [sq for _, sq in edge_records] + [sq[::-1] for _, sq in edge_records] + \
            [sq for _, sq in demo_records[:10]] + make_reads(rng, motif, 40, 30000) + \
            [motif * 4000, (motif * 4000)[::-1], "ACGT" * 6000, "A" * 25000, "N" * 3000 + motif * 500]
    rows, cws, launches = _scan(engine, pats, motif, reads, W, s, t, M, True, monkeypatch)
    rows0, cws0, launches0 = _scan(engine, pats, motif, reads, W, s, t, M, False, monkeypatch)
    rows1, cws1, _ = _scan(engine, pats, motif, reads, W, s, t, M, True, monkeypatch, no_groups=True)
    # ~100 reads on a grid of several hundred CTAs: the scans above dealt every read's tiles over four CTAs; one CTA per read:
    rows2, cws2, _ = _scan(engine, pats, motif, reads, W, s, t, M, True, monkeypatch, no_split=True)
    assert launches == 3 and launches0 == 4      # K1, K2, fused K3+K4  vs  K1, K2, K3, K4
    assert rows.tobytes() == rows0.tobytes() == rows1.tobytes() == rows2.tobytes()
    assert set(cws) == set(cws0) == set(cws1) == set(cws2)
    assert all(np.array_equal(cws[i], cws1[i]) and np.array_equal(cws[i], cws2[i]) for i in cws)
    n_cp = 0
    for i, seq in enumerate(reads):
        row = rows[i]
        if len(seq) == 0:
            assert row["status"] == engine.ST_FILTERED
            continue
        tail = engine.TAIL_NAMES[row["tail"]]
        counts = orc.window_counts(orc.oriented_region(seq, tail, t, M), pats, W, s)
        assert row["n_windows"] == counts.shape[0]
        if counts.shape[0] < 7:
            assert row["status"] == engine.ST_BADSEG and row["telo_length"] == -1
            continue
        c_w = counts.sum(axis=1)
        assert np.array_equal(cws[i], engine.group_sums(c_w)), (cfg, i)
        assert np.array_equal(cws0[i], engine.group_sums(c_w)), (cfg, i)
        b = orc.change_point_exact(c_w)
        assert (int(row["status"]), int(row["bkp"]), int(row["telo_length"])) == (engine.ST_PASS, b, t + s * b), (cfg, i)
        n_cp += 1
    assert n_cp > 40


@pytest.mark.parametrize("seed", range(24))
def test_random_geometry(seed, monkeypatch):
    """Random motif / phrase / window geometry / orientation forcing, bit-parallel kernel vs oracle."""
    from topsicle_b200 import engine
    rng = np.random.default_rng(5000 + seed)
    motif = MOTIFS[seed % len(MOTIFS)]
    k = int(rng.integers(2, min(8, len(motif)) + 1))
    pats = orc.patterns_to_search(motif, k)
    W = int(rng.choice([k + 32, 50, 64, 100, 101, 150, 333, 700]))
    W = max(W, k + 32)
    s = int(rng.choice([1, 2, 3, 5, 6, 7, 13, 40]))
    t = int(rng.choice([0, 1, 17, 100, 200]))
    M = int(rng.choice([1000, 2500, 5000, 20000, 30000]))
    force = [None, None, "forward", "reverse"][int(rng.integers(0, 4))]
    reads = make_reads(rng, motif, 60, 32000) + [motif * 3000, (motif * 3000)[::-1]]
    rows, cws, _ = _scan(engine, pats, motif, reads, W, s, t, M, True, monkeypatch, force_tail=force)
    n = 0
    for i, seq in enumerate(reads):
        row = rows[i]
        tail = engine.TAIL_NAMES[row["tail"]]
        if force:
            assert tail == force
        counts = orc.window_counts(orc.oriented_region(seq, tail, t, M), pats, W, s)
        assert row["n_windows"] == counts.shape[0]
        if counts.shape[0] < 7:
            assert row["status"] == engine.ST_BADSEG
            continue
        c_w = counts.sum(axis=1)
        assert np.array_equal(cws[i], engine.group_sums(c_w)), (seed, i)
        b = orc.change_point_exact(c_w)
        assert (int(row["bkp"]), int(row["telo_length"])) == (b, t + s * b), (seed, i)
        n += 1
    assert n > 10


def test_many_reads_many_tiles(monkeypatch):
    """A batch with thousands of TRC-pass reads of five tiles each (the per-read tile counters and the fused
    change point under contention), ends + regions protocol included: rows equal the plain kernels'."""
    from topsicle_b200 import engine, synth
    spec = synth.CONFIGS[5]
    n = 20000
    off = synth.read_lengths(spec, 0, n)
    bases = np.empty(int(off[-1]), dtype=np.uint8)
    synth.fill_reads(spec, 0, off, bases)
    pats = orc.patterns_to_search("TTAGGG", 4)
    out = []
    for bitpar in (True, False):
        if bitpar:
            monkeypatch.delenv("TPS_K3_BITPAR", raising=False)
        else:
            monkeypatch.setenv("TPS_K3_BITPAR", "0")
        with engine.ScanContext(pats, len_telopattern=6, cutoff=0.7, min_seq_length=9000, window_size=50, slide=3,
                                max_batch_reads=n, max_batch_bases=int(off[-1]) + 4096, n_slots=3) as ctx:
            bids = [ctx.submit(bases, off) for _ in range(3)]
            rows = [ctx.wait(b)[0] for b in bids]
            assert rows[0].tobytes() == rows[1].tobytes() == rows[2].tobytes()
            out.append(rows[0])
    assert out[0].tobytes() == out[1].tobytes()
    assert int((out[0]["status"] == engine.ST_PASS).sum()) > 500
