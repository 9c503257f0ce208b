"""GPU tier: the streaming pipeline (C reader -> pinned batches -> kernels -> ordered harvest) on a
synthetic FASTQ cut into many small batches, three telophrases from one upload, two workers on the
same device -- every field of every TRC-pass read equals the oracle's."""
import os

import numpy as np
import pytest

from oracle import topsicle_oracle as orc
from topsicle_b200 import engine, pipeline, synth
from topsicle_b200.patterns import patterns_to_search

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth_file(tmp_path_factory):
    spec = dict(synth.CONFIGS[2])
    spec.update(f_telo=0.06, near_frac=0.03, len_max=60000)
    bases, off, kinds = synth.generate(spec, 0, 2500)
    path = str(tmp_path_factory.mktemp("syn") / "reads.fastq")
    synth.write_fastq(path, bases, off, prefix="syn2")
    buf = bases.tobytes().decode("ascii")
    recs = [(f"syn2_{i}", buf[int(off[i]):int(off[i + 1])]) for i in range(2500)]
    return path, recs


def _cfgs(phrases, **kw):
    kw.setdefault("slide", 6)
    return [pipeline.ScanConfig(patterns=patterns_to_search("CCCTAA", k), len_telopattern=6, phrase=k, **kw)
            for k in phrases]


def _check(passes, recs, k, cutoff, W=100, s=6, t=100, M=20000, minlen=9000, counts=False):
    want = orc.scan_records(recs, "CCCTAA", k, cutoff, minlen, W, s, t, M, exact=True, want_counts=counts)
    want = [w for w in want]
    assert [p.read_id for p in passes] == [w["id"] for w in want]
    for p, w in zip(passes, want):
        assert (p.tail, p.count, p.literal, p.telo_length) == (
            w["tail"], w["count"], patterns_to_search("CCCTAA", k)[w["best"]], w["telo_length"]), p.read_id
        assert p.trc == w["trc"] and p.index == int(p.read_id.split("_")[1])
        if counts:
            assert np.array_equal(p.counts.astype(np.int64), w["counts"])
    return len(want)


def test_many_small_batches_three_phrases_two_workers(synth_file):
    path, recs = synth_file
    cfgs = _cfgs([4, 5, 6], cutoff=0.4)
    stats, per = pipeline.collect_file(path, cfgs, devices=[0, 0], max_batch_bases=6 << 20, max_batch_reads=512,
                                       depth=3)
    assert stats.n_reads == 2500 and stats.n_batches > 8
    assert stats.n_bases == sum(len(s) for _, s in recs)
    n = [_check(per[i], recs, k, 0.4) for i, k in enumerate([4, 5, 6])]
    assert n[0] > 60 and n[0] >= n[2]


def test_rawcount_and_capacity_split(synth_file):
    """Raw count tables through the pipeline; a tiny pass / rawcount capacity forces the split re-scan."""
    path, recs = synth_file
    cfgs = _cfgs([4], cutoff=0.6, want_rawcount=True, window_size=50, slide=3)
    stats, per = pipeline.collect_file(path, cfgs, devices=[0], max_batch_bases=16 << 20, max_batch_reads=1024,
                                       max_pass_reads=8, rawcount_capacity=6617 * 12 * 3)
    assert _check(per[0], recs, 4, 0.6, W=50, s=3, counts=True) > 40


def test_cli_on_synthetic_file(synth_file, tmp_path):
    """The CLI end to end on the synthetic file == oracle CSV text (CRLF rows, %.3f TRC)."""
    from topsicle_b200 import main as tmain
    path, recs = synth_file
    out = tmp_path / "out"
    if hasattr(tmain.tprint, "logfile"):
        del tmain.tprint.logfile
    os.environ["TOPSICLE_BATCH_BASES"] = str(8 << 20)
    try:
        tmain.main(["-i", path, "-o", str(out), "--pattern", "CCCTAA", "--cutoff", "0.5"])
    finally:
        del os.environ["TOPSICLE_BATCH_BASES"]
    rows = orc.scan_records(recs, "CCCTAA", 4, 0.5, 9000, 100, 6, 100, 20000, exact=True)
    assert open(out / "telolengths_all.csv", newline="").read() == orc.csv_text("reads", 4, rows)
    sub = open(out / "reads_trc_over_0.5.fastq").read().split("\n")
    assert [ln[1:] for ln in sub[0::4] if ln] == [r["id"] for r in rows]


def test_ends_first_many_batches_three_phrases_two_workers(synth_file):
    """Ends-first mode on the real kernels: head + tail batches cut by file text, three phrases from one upload,
    region batches cut by a small pass capacity, raw counts -- every field equals the oracle's."""
    path, recs = synth_file
    cfgs = _cfgs([4, 5, 6], cutoff=0.4)
    stats, per = pipeline.collect_file(path, cfgs, devices=[0, 0], max_batch_bases=6 << 20, max_batch_reads=512,
                                       depth=3, ends_first=True, ends_raw_bytes=6 << 20, max_pass_reads=16)
    assert stats.n_reads == 2500 and stats.n_batches > 8
    assert stats.n_bases == sum(len(s) for _, s in recs)
    n = [_check(per[i], recs, k, 0.4) for i, k in enumerate([4, 5, 6])]
    assert n[0] > 60 and n[0] >= n[2]
    cfgs = _cfgs([4], cutoff=0.6, want_rawcount=True, window_size=50, slide=3)
    stats, per = pipeline.collect_file(path, cfgs, devices=[0], max_batch_bases=16 << 20, max_batch_reads=1024,
                                       max_pass_reads=8, rawcount_capacity=6617 * 12 * 3, ends_first=True)
    assert _check(per[0], recs, 4, 0.6, W=50, s=3, counts=True) > 40


def test_device_span_of_overlapping_scans_across_contexts():
    """tps_elapsed_between: the device time from the first kernel of one context's scan to the end of another
    context's later scan (what bench.py times its K steps with) -- ordered, consistent with the per-scan
    timeline, and loud when a scan is not in the event ring."""
    import torch
    from topsicle_b200 import engine, synth
    from topsicle_b200.patterns import patterns_to_search
    spec = synth.CONFIGS[2]
    n = 4000
    off = synth.read_lengths(spec, 0, n)
    host = np.empty(int(off[-1]), dtype=np.uint8)
    synth.fill_reads(spec, 0, off, host)
    dev = torch.device("cuda", 0)
    db = torch.empty((host.size + 2047) // 2048 * 2048, dtype=torch.uint8, device=dev)
    db[:host.size].copy_(torch.from_numpy(host))
    do = torch.from_numpy(off.view(np.int64)).to(dev)
    rows = [torch.empty(n * 40, dtype=torch.uint8, device=dev) for _ in range(4)]
    kw = dict(cutoff=0.7, min_seq_length=9000, n_slots=3, max_batch_reads=n, max_batch_bases=host.size + 4096)
    with engine.ScanContext(patterns_to_search("CCCTAA", 4), len_telopattern=6, slide=6, **kw) as a, \
            engine.ScanContext(patterns_to_search("TTTAGGG", 5), len_telopattern=7, slide=7, **kw) as b:
        for i in range(6):          # a, b, a, b, a, b on rotating slots
            c = a if i % 2 == 0 else b
            c.scan_device(db.data_ptr(), do.data_ptr(), n, host.size, rows[i % 4].data_ptr(), slot=(i // 2) % 3)
        a.sync()
        b.sync()
        span = a.elapsed_to(2, 0, b, 0, 3)                    # first scan of a -> end of the last scan of b
        assert span > 0
        own = a.timeline(0, 2)                                # a's last scan relative to a's first
        assert 0 < own[0] <= own[3] and a.elapsed_to(2, 0, a, 0, 3) == pytest.approx(own[3], rel=1e-3, abs=1e-3)
        assert span >= a.elapsed_to(2, 0, b, 2, 3) > 0        # b's first scan ended earlier than its last
        t = [c.timings(0)["total"] for c in (a, b)]
        assert span >= max(t)
        with pytest.raises(engine.TpsError):
            a.elapsed_to(3, 0, b, 0, 3)                       # a has only three scans on record
