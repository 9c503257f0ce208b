#!/usr/bin/env python
"""GPU: time K1 (tps_pack_kernel) for each TPS_K1_UNROLL setting on one config-2 batch."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from topsicle_b200 import engine, synth
from topsicle_b200.patterns import patterns_to_search

spec = synth.CONFIGS[2]
R = 63488
off = synth.read_lengths(spec, 0, R)
n = int(off[-1])
hb = np.empty(n, np.uint8)
synth.fill_reads(spec, 0, off, hb)
dev = torch.device("cuda", 0)
pad = (n + 2047) // 2048 * 2048
db = torch.empty(pad, dtype=torch.uint8, device=dev); db[:n].copy_(torch.from_numpy(hb))
do = torch.from_numpy(off.view(np.int64)).to(dev)
rows = torch.empty(R * 40, dtype=torch.uint8, device=dev)
for var in sys.argv[1:] or ["2", "4", "8"]:
    if var.startswith("p"):
        os.environ["TPS_K1_PROBE"] = var[1:]
    else:
        os.environ.pop("TPS_K1_PROBE", None)
        os.environ["TPS_K1_UNROLL"] = var
    ctx = engine.ScanContext(patterns_to_search("CCCTAA", 4), len_telopattern=6, n_slots=1, max_batch_reads=R, max_batch_bases=n)
    for i in range(13):
        ctx.scan_device(db.data_ptr(), do.data_ptr(), R, n, rows.data_ptr())
    ctx.sync()
    t = [ctx.timings(b) for b in range(10)]
    k1 = statistics.mean(x["k1_pack"] for x in t)
    print(f"unroll {var}: K1 {k1*1e3:.1f} us  {1.25*n/k1/1e6:.0f} GB/s alg  ({1.25*n/k1/1e6/6554.6:.3f} of peak); "
          f"K2 {statistics.mean(x['k2_trc'] for x in t)*1e3:.1f} us K3K4 {statistics.mean(x['k3_windows_cp'] for x in t)*1e3:.1f} us")
    ctx.close()
