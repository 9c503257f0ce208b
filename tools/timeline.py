#!/usr/bin/env python
"""GPU: event timeline of overlapping scans (config-2 batches on N slots): when each batch's K1 / K2 / K3+K4
start and end relative to the first batch, to see what actually overlaps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from topsicle_b200 import engine, synth
from topsicle_b200.patterns import patterns_to_search

n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
spec = synth.CONFIGS[2]
R = 63488
dev = torch.device("cuda", 0)
bufs = []
for b in range(2):
    off = synth.read_lengths(spec, b * R, R)
    n = int(off[-1])
    hb = np.empty(n, np.uint8)
    synth.fill_reads(spec, b * R, off, hb)
    pad = (n + 2047) // 2048 * 2048
    db = torch.empty(pad, dtype=torch.uint8, device=dev); db[:n].copy_(torch.from_numpy(hb))
    do = torch.from_numpy(off.view(np.int64)).to(dev)
    bufs.append((db, do, n))
rows = [torch.empty(R * 40, dtype=torch.uint8, device=dev) for _ in range(n_slots)]
ctx = engine.ScanContext(patterns_to_search("CCCTAA", 4), len_telopattern=6, n_slots=max(n_slots, 1), max_batch_reads=R,
                         max_batch_bases=max(b[2] for b in bufs))
def go(k):
    for i in range(k):
        db, do, n = bufs[i % 2]
        ctx.scan_device(db.data_ptr(), do.data_ptr(), R, n, rows[i % n_slots].data_ptr(), slot=i % n_slots)
    ctx.sync()
go(6)
go(steps)
brief = os.environ.get("TL_BRIEF")
if brief:
    ts = [ctx.timeline(b, steps - 1) for b in range(steps - 1, -1, -1)]
    k1 = sorted(t[1] - t[0] for t in ts)
    print(f"slots {n_slots}: {ts[-1][3]*1e3/steps:7.1f} us/batch   K1 min {k1[0]*1e3:.1f} med {k1[len(k1)//2]*1e3:.1f} max {k1[-1]*1e3:.1f}")
    ctx.close()
    sys.exit(0)
print(f"slots {n_slots}: per batch (K1 start, K1 end, K2 end, K4 end) in us from the first K1 start")
prev = None
for back in range(steps - 1, -1, -1):
    t = ctx.timeline(back, steps - 1)
    print(f"  batch {steps-1-back:2d} slot {(steps-1-back)%n_slots}: " + "  ".join(f"{x*1e3:8.1f}" for x in t) +
          f"   K1 {1e3*(t[1]-t[0]):6.1f}  K2 {1e3*(t[2]-t[1]):6.1f} K3K4 {1e3*(t[3]-t[2]):6.1f}")
last = ctx.timeline(0, steps - 1)[3]
print(f"  {steps} batches in {last*1e3:.1f} us -> {last*1e3/steps:.1f} us / batch")
ctx.close()
