"""CPU tier (oracle stand-in for the GPU): one Scanner, several pattern sets, every input file scanned under its
own -- BASELINE config 5's "mixed-species batch": the reference takes one --pattern per invocation
(main.py:321), so three species are three invocations; here they are three configs of one Scanner whose
devices take batches of any file."""
import os

import pytest

from tests import fake_engine
from tests.conftest import GOLD


def _cfg(pipeline, patterns_to_search, motif, k, **kw):
    return pipeline.ScanConfig(patterns=patterns_to_search(motif, k), len_telopattern=len(motif), phrase=k,
                               slide=kw.pop("slide", len(motif)), **kw)


@pytest.mark.parametrize("ends_first", [False, True])
def test_files_scanned_under_their_own_pattern_set(monkeypatch, ends_first):
    fake_engine.install(monkeypatch)
    from topsicle_b200 import pipeline
    from topsicle_b200.patterns import patterns_to_search
    demo = os.path.join(GOLD, "demo.fastq.gz")
    cfgs = [_cfg(pipeline, patterns_to_search, "CCCTAAA", 5, slide=6), _cfg(pipeline, patterns_to_search, "AAACCCT", 5),
            _cfg(pipeline, patterns_to_search, "CCCTAA", 4, cutoff=0.4)]
    key = lambda ps: [(p.index, p.read_id, p.literal, p.tail, p.count, p.telo_length) for p in ps]  # noqa: E731
    want = []
    for c in cfgs:                                  # one invocation per pattern, as the reference would run them
        _, per = pipeline.collect_file(demo, [c], max_batch_reads=7, ends_first=ends_first)
        want.append(key(per[0]))
    assert [len(w) for w in want] == [17, 17, 27]
    got = [[] for _ in cfgs]
    jobs = [pipeline.FileJob(demo, (lambda res, k=k: got[k].extend(res.passes[0])), cfg_ids=[k]) for k in (2, 0, 1, 0)]
    with pipeline.Scanner(cfgs, devices=[0, 1], max_batch_reads=7, leaders=(0, 1, 2), ends_first=ends_first) as sc:
        stats = sc.scan_files(jobs, readers=2)
    assert all(st.n_reads == 44 for st in stats)
    assert key(got[1]) == want[1] and key(got[2]) == want[2]
    assert sorted(key(got[0])) == sorted(want[0] + want[0])       # config 0 scanned the file twice


def test_cfg_ids_must_start_with_a_leader(monkeypatch):
    fake_engine.install(monkeypatch)
    from topsicle_b200 import pipeline
    from topsicle_b200.patterns import patterns_to_search
    cfgs = [_cfg(pipeline, patterns_to_search, "CCCTAAA", 5), _cfg(pipeline, patterns_to_search, "CCCTAAA", 4)]
    with pipeline.Scanner(cfgs, devices=[0]) as sc:
        with pytest.raises(ValueError):
            sc.scan_files([pipeline.FileJob(os.path.join(GOLD, "demo.fastq.gz"), lambda r: None, cfg_ids=[1])])
        sc.scan_files([pipeline.FileJob(os.path.join(GOLD, "demo.fastq.gz"), lambda r: None, cfg_ids=[0])])
