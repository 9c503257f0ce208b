"""Host-side placement: allocate a device's pinned staging buffers (and run its worker thread) on the
CPUs / NUMA node the GPU hangs off.  On an 8-GPU box every GPU pulls its batches over its own PCIe
link; if the pinned buffers of all devices sit on one socket, the inter-socket link and that socket's
memory controllers become the bottleneck of the H2D stream.  NVML knows the ideal CPU set of each GPU."""
from __future__ import annotations

import contextlib
import os


def cpus_near_device(device: int):
    """CPU ids NVML reports as local to CUDA device `device` (respecting the current affinity mask),
    or None if that cannot be determined."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:                                   # CUDA ordinal -> NVML index / UUID
            tok = [t.strip() for t in vis.split(",") if t.strip()][device]
            h = pynvml.nvmlDeviceGetHandleByUUID(tok.encode()) if tok.startswith("GPU-") else \
                pynvml.nvmlDeviceGetHandleByIndex(int(tok))
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(device)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i * 64 + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return sorted(cpus) or None
    except Exception:  # noqa: BLE001 - placement is an optimisation, never a requirement
        return None


@contextlib.contextmanager
def near_device(device: int):
    """Temporarily pin the calling thread to the CPUs local to `device` (first-touch / cudaHostAlloc then
    places pages on that node).  No-op when the topology is unknown or TOPSICLE_NO_NUMA is set."""
    if os.environ.get("TOPSICLE_NO_NUMA"):
        yield None
        return
    old = os.sched_getaffinity(0)
    cpus = cpus_near_device(device)
    if cpus and set(cpus) != old:
        os.sched_setaffinity(0, cpus)
        try:
            yield cpus
        finally:
            os.sched_setaffinity(0, old)
    else:
        yield cpus
