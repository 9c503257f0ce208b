"""`synth-v1` synthetic read generator -- ctypes wrapper over csrc/tps_host.c.

Used by bench.py and the tests to build the BASELINE.json workloads (configs 2-5).  Reads
are a pure function of (seed, global read index), so shards can be generated per rank.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libtps_host.so")


class SynthCfg(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("len_kind", C.c_uint32), ("len_a", C.c_double), ("len_b", C.c_double),
        ("len_min", C.c_uint32), ("len_max", C.c_uint32), ("f_telo", C.c_double),
        ("telo_min", C.c_uint32), ("telo_max", C.c_uint32),
        ("sub_rate", C.c_double), ("ins_rate", C.c_double), ("del_rate", C.c_double),
        ("n_rate", C.c_double), ("near_frac", C.c_double), ("lower_frac", C.c_double),
        ("motif_len", C.c_uint32), ("motif", C.c_char * 32),
    ]


_hlib = None


def host_library() -> C.CDLL:
    global _hlib
    if _hlib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} not found: run __graft_entry__.build()")
        lib = C.CDLL(HOST_LIB_PATH)
        lib.tps_synth_lengths.restype = C.c_int
        lib.tps_synth_lengths.argtypes = [C.POINTER(SynthCfg), C.c_uint64, C.c_uint32, C.c_void_p]
        lib.tps_synth_fill.restype = C.c_int
        lib.tps_synth_fill.argtypes = [C.POINTER(SynthCfg), C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int]
        lib.tps_host_threads.restype = C.c_int
        _hlib = lib
    return _hlib


ONT = dict(sub_rate=0.03, ins_rate=0.01, del_rate=0.01)
HIFI = dict(sub_rate=0.001, ins_rate=0.0005, del_rate=0.0005)

# BASELINE.json configs[1..4].  `total_reads` is the size the config is quoted on.
CONFIGS = {
    2: dict(name="config2: synthetic human ONT, 1M reads, N50 30 kb, --pattern CCCTAA --minSeqLength 9000",
            seed=2002, total_reads=1_000_000, len_kind=1, len_a=math.log(30000.0) - 0.49, len_b=0.7,
            len_min=1000, len_max=250_000, f_telo=0.01, telo_min=2000, telo_max=12000, motif="CCCTAA", **ONT,
            cli=dict(pattern="CCCTAA", minSeqLength=9000)),
    3: dict(name="config3: synthetic PacBio HiFi, 5M x 15 kb, --pattern CCCTAA --telophrase 4 5 6 --cutoff 0.4 0.7",
            seed=2003, total_reads=5_000_000, len_kind=0, len_a=15000.0, len_b=0.0, len_min=15000, len_max=15000,
            f_telo=0.02, telo_min=2000, telo_max=12000, motif="CCCTAA", **HIFI,
            cli=dict(pattern="CCCTAA", telophrase=[4, 5, 6], cutoff=[0.4, 0.7], rawcountpattern=True)),
    4: dict(name="config4: synthetic ultra-long ONT 100 kb+, ~200 Gbases, --pattern CCCTAA --maxlengthtelo 20000",
            seed=2004, total_reads=1_333_000, len_kind=2, len_a=100_000.0, len_b=50_000.0, len_min=100_000,
            len_max=1_000_000, f_telo=0.005, telo_min=2000, telo_max=18000, motif="CCCTAA", **ONT,
            cli=dict(pattern="CCCTAA", maxlengthtelo=20000)),
    # three sub-batches of 1 M reads, one per species: the reference takes one --pattern per invocation
    # (main.py:321), so each sub-batch is scanned under its own pattern set
    5: dict(name="config5: mixed species (TTAGGG, TTTAGGG, AAACCCT sub-batches), LogNormal N50 20 kb, "
                 "--windowSize 50 --slide 3",
            seed=2005, total_reads=3_000_000, len_kind=1, len_a=math.log(20000.0) - 0.49, len_b=0.7,
            len_min=1000, len_max=250_000, f_telo=0.05, telo_min=2000, telo_max=12000, motif="TTAGGG", **ONT,
            sub_batches=["TTAGGG", "TTTAGGG", "AAACCCT"],
            cli=dict(pattern="TTAGGG", windowSize=50, slide=3)),
}


def make_cfg(spec: dict, motif: str | None = None) -> SynthCfg:
    c = SynthCfg()
    c.seed = spec["seed"]
    c.len_kind, c.len_a, c.len_b = spec["len_kind"], spec["len_a"], spec["len_b"]
    c.len_min, c.len_max = spec["len_min"], spec["len_max"]
    c.f_telo, c.telo_min, c.telo_max = spec["f_telo"], spec["telo_min"], spec["telo_max"]
    c.sub_rate, c.ins_rate, c.del_rate = spec["sub_rate"], spec["ins_rate"], spec["del_rate"]
    c.n_rate = spec.get("n_rate", 0.0005)
    c.near_frac = spec.get("near_frac", 0.01)
    c.lower_frac = spec.get("lower_frac", 0.01)
    m = (motif or spec["motif"]).upper().encode()
    c.motif_len = len(m)
    c.motif = m
    return c


def read_lengths(spec: dict, first_read: int, n_reads: int, motif: str | None = None) -> np.ndarray:
    """offsets[n_reads+1] (uint64) of reads first_read .. first_read+n_reads-1 laid back to back."""
    cfg = make_cfg(spec, motif)
    off = np.zeros(n_reads + 1, dtype=np.uint64)
    rc = host_library().tps_synth_lengths(C.byref(cfg), first_read, n_reads, off.ctypes.data)
    assert rc == 0
    return off


def fill_reads(spec: dict, first_read: int, offsets: np.ndarray, out: np.ndarray, threads: int = 0,
               motif: str | None = None) -> np.ndarray:
    """Generate the reads into `out` (uint8, >= offsets[-1] bytes); returns per-read class codes."""
    cfg = make_cfg(spec, motif)
    n = len(offsets) - 1
    assert out.dtype == np.uint8 and out.size >= int(offsets[-1])
    kinds = np.zeros(n, dtype=np.uint8)
    rc = host_library().tps_synth_fill(C.byref(cfg), first_read, n, offsets.ctypes.data, out.ctypes.data,
                                       kinds.ctypes.data, threads)
    assert rc == 0
    return kinds


def generate(spec: dict, first_read: int, n_reads: int, threads: int = 0, motif: str | None = None):
    off = read_lengths(spec, first_read, n_reads, motif)
    bases = np.empty(int(off[-1]), dtype=np.uint8)
    kinds = fill_reads(spec, first_read, off, bases, threads, motif)
    return bases, off, kinds


def write_fastq(path: str, bases: np.ndarray, offsets: np.ndarray, prefix: str = "syn", first_read: int = 0):
    """4-line FASTQ, header `@{prefix}_{i}`, quality 'I' * L."""
    with open(path, "wb") as fh:
        for i in range(len(offsets) - 1):
            s = bases[int(offsets[i]):int(offsets[i + 1])].tobytes()
            fh.write(b"@%s_%d\n%s\n+\n%s\n" % (prefix.encode(), first_read + i, s, b"I" * len(s)))
