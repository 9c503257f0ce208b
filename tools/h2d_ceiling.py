#!/usr/bin/env python
"""What the box gives a plain host-to-device stream: N GPUs at once, each fed by cudaMemcpyAsync from pinned
256 MiB buffers, no kernels.  Measured two ways -- one process with one thread per GPU (how `pipeline.Scanner`
drives the devices of a box) and one process per GPU (how `bench.py --gpus N` under torchrun does) -- so that the
end-to-end numbers of the scan can be read against the ceiling of the machine rather than of one PCIe link.

  python tools/h2d_ceiling.py [--gpus 1 2 4 8] [--seconds 2] [--mib 256] [--numa 0|1]      -> JSON lines
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def pump(dev, nbytes, seconds, start_evt, out, numa_local, slots=3, write_combined=False):
    """Copy `slots` pinned buffers of nbytes round-robin to device `dev` for `seconds`; out[dev] = GB/s."""
    import torch
    from topsicle_b200 import numa
    torch.cuda.set_device(dev)
    ctx = numa.near_device(dev) if numa_local else None
    if ctx:
        ctx.__enter__()
    try:
        host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        for h in host:
            h.fill_(65)
    finally:
        if ctx:
            ctx.__exit__(None, None, None)
    devb = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}") for _ in range(slots)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(slots)]
    for s, h, d in zip(streams, host, devb):          # warm-up
        with torch.cuda.stream(s):
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(dev)
    start_evt.wait()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        for s, h, d in zip(streams, host, devb):
            with torch.cuda.stream(s):
                d.copy_(h, non_blocking=True)
        for s in streams:
            s.synchronize()
        n += slots
    dt = time.perf_counter() - t0
    out[dev] = n * nbytes / dt / 1e9


def proc_main(dev, nbytes, seconds, barrier, q, numa_local):
    out = {}
    evt = threading.Event()
    barrier.wait()
    evt.set()
    pump(dev, nbytes, seconds, evt, out, numa_local)
    q.put((dev, out[dev]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--numa", type=int, default=1)
    a = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    nbytes = a.mib << 20
    for n in [g for g in a.gpus if g <= have]:
        # one process, one thread per GPU
        out, evt = {}, threading.Event()
        ts = [threading.Thread(target=pump, args=(d, nbytes, a.seconds, evt, out, bool(a.numa))) for d in range(n)]
        for t in ts:
            t.start()
        time.sleep(1.0 + 0.3 * n)
        evt.set()
        for t in ts:
            t.join()
        print(json.dumps({"mode": "threads_in_one_process", "gpus": n, "pinned_mib": a.mib, "numa_local": bool(a.numa),
                          "gb_per_s_per_gpu": [round(out[d], 2) for d in range(n)], "gb_per_s_total": round(sum(out.values()), 1)}),
              flush=True)
        # one process per GPU
        ctx = mp.get_context("spawn")
        bar, q = ctx.Barrier(n), ctx.Queue()
        ps = [ctx.Process(target=proc_main, args=(d, nbytes, a.seconds, bar, q, bool(a.numa))) for d in range(n)]
        for p in ps:
            p.start()
        got = dict(q.get() for _ in range(n))
        for p in ps:
            p.join()
        print(json.dumps({"mode": "one_process_per_gpu", "gpus": n, "pinned_mib": a.mib, "numa_local": bool(a.numa),
                          "gb_per_s_per_gpu": [round(got[d], 2) for d in range(n)], "gb_per_s_total": round(sum(got.values()), 1)}),
              flush=True)


if __name__ == "__main__":
    main()
