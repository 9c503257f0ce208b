"""CPU tier, world_size 2 over gloo: the one-process-per-GPU path (shard by bases, scan the shard,
gather the result rows on rank 0) gives exactly the single-process rows."""
import os
import socket
import sys

import numpy as np
import pytest

from tests.conftest import GOLD, REPO
from topsicle_b200 import sharding


def test_shard_by_bases_properties():
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 100000, 1000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 3, 8, 1500):
        sh = sharding.shard_by_bases(off, world)
        assert sh[0][0] == 0 and sh[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        if world <= 8:
            loads = [int(off[hi] - off[lo]) for lo, hi in sh]
            assert max(loads) - min(loads) <= 2 * lens.max()
    assert sharding.shard_by_bases(np.zeros(1, np.uint64), 4) == [(0, 0)] * 4
    assert sharding.shard_by_count(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from oracle import topsicle_oracle as orc
    from tests.fake_engine import OracleContext
    from topsicle_b200 import engine, pipeline, sharding
    from topsicle_b200.patterns import patterns_to_search
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    recs = list(orc.read_fastx(os.path.join(GOLD, "demo.fastq.gz")))
    bases, off = engine.pack_reads([s for _, s in recs])
    lo, hi = sharding.shard_by_bases(off, world)[rank]
    cfg = pipeline.ScanConfig(patterns=patterns_to_search("CCCTAAA", 5), len_telopattern=7, phrase=5, slide=6)
    ctx = OracleContext(cfg, rank, 1 << 10, 1 << 24, 1)
    sub_off = (off[lo:hi + 1] - off[lo]).astype(np.uint64)
    rows, _ = ctx.scan(bases[int(off[lo]):int(off[hi])], sub_off)
    t = sharding.reduce_scalar(float(rank + 1), "max")
    s = sharding.reduce_scalar(float(hi - lo), "sum")
    allrows = sharding.gather_rows(rows, dst=0)
    if rank == 0:
        assert t == world and s == len(recs)
        np.save(out_path, allrows)
    else:
        assert allrows is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_equal_one_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle import topsicle_oracle as orc
    from tests.fake_engine import OracleContext
    from topsicle_b200 import engine, pipeline
    from topsicle_b200.patterns import patterns_to_search
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "rows.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    recs = list(orc.read_fastx(os.path.join(GOLD, "demo.fastq.gz")))
    bases, off = engine.pack_reads([s for _, s in recs])
    cfg = pipeline.ScanConfig(patterns=patterns_to_search("CCCTAAA", 5), len_telopattern=7, phrase=5, slide=6)
    want, _ = OracleContext(cfg, 0, 1 << 10, 1 << 24, 1).scan(bases, off)
    assert got.tobytes() == want.tobytes()
    assert int((got["status"] == engine.ST_PASS).sum()) == 17
