"""GPU tier, needs >= 2 devices (skipped otherwise): the multi-device product path -- one process,
`pipeline.Scanner(devices=[0, 1])`, batches dealt to whichever GPU asks next -- must give the rows of a
single-device scan, in file order, also for files scanned under different pattern sets (BASELINE config 5) and in
ends-first mode.  Reference: files over a process pool, main.py:232-235."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def two_devices():
    from topsicle_b200 import engine
    if engine.device_count() < 2:
        pytest.skip("needs two CUDA devices")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    from topsicle_b200 import synth
    d = tmp_path_factory.mktemp("multi")
    spec = synth.CONFIGS[5]
    out = []
    for mi, motif in enumerate(spec["sub_batches"]):
        off = synth.read_lengths(spec, 0, 12000, motif)
        bases = np.empty(int(off[-1]), dtype=np.uint8)
        synth.fill_reads(spec, 0, off, bases, motif=motif)
        path = str(d / f"{motif}.fastq")
        synth.write_fastq(path, bases, off, prefix=motif)
        out.append((path, mi, motif))
    return out


@pytest.mark.parametrize("ends_first", [False, True])
def test_two_devices_equal_one_device(two_devices, files, ends_first):
    from topsicle_b200 import pipeline, synth
    from topsicle_b200.patterns import patterns_to_search
    cli = synth.CONFIGS[5]["cli"]
    cfgs = [pipeline.ScanConfig(patterns=patterns_to_search(m, len(m) - 2), len_telopattern=len(m), phrase=len(m) - 2,
                                window_size=cli["windowSize"], slide=cli["slide"]) for _, _, m in files]
    key = lambda ps: [(p.index, p.read_id, p.literal, p.tail, p.count, p.status, p.n_windows, p.telo_length) for p in ps]  # noqa: E731

    def run(devices):
        got = {p: [] for p, _, _ in files}
        with pipeline.Scanner(cfgs, devices=devices, leaders=(0, 1, 2), ends_first=ends_first, max_batch_reads=2048,
                              max_batch_bases=1 << 26) as sc:
            stats = sc.scan_files([pipeline.FileJob(p, (lambda res, p=p: got[p].extend(res.passes[0])), cfg_ids=[mi])
                                   for p, mi, _ in files])
        assert all(st.n_reads == 12000 and st.n_batches >= 3 for st in stats)
        return {p: key(g) for p, g in got.items()}

    one, two = run([0]), run([0, 1])
    assert one == two
    assert all(len(v) > 100 for v in one.values())
