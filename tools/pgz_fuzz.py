#!/usr/bin/env python
"""Differential fuzz of the parallel gzip inflater (csrc/tps_pgz.c) through its C ABI: random texts (FASTQ-like,
repeats at every distance, skewed alphabets, binary) compressed with every zlib level / strategy / memLevel, as one
or several members, inflated with random thread counts, piece sizes and read sizes -- the bytes must be the input.
  python tools/pgz_fuzz.py [--cases 300] [--seed 1]"""
import argparse
import ctypes as C
import os
import sys
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")


def lib():
    L = C.CDLL(os.path.join(REPO, "topsicle_b200", "libtps_host.so"))
    L.tps_pgz_open.restype = C.c_void_p
    L.tps_pgz_open.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
    L.tps_pgz_close.argtypes = [C.c_void_p]
    L.tps_pgz_read.restype = C.c_int64
    L.tps_pgz_read.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.tps_pgz_error.restype = C.c_char_p
    L.tps_pgz_error.argtypes = [C.c_void_p]
    L.tps_pgz_set_piece.argtypes = [C.c_void_p, C.c_uint64]
    L.tps_pgz_inflate_block.restype = C.c_int
    L.tps_pgz_inflate_block.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    return L


def make_text(rng, kind, n):
    if kind == 0:      # FASTQ-like
        out = []
        while sum(map(len, out)) < n:
            L = int(rng.integers(50, 20000))
            seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L))
            if rng.random() < 0.2:
                seq = (b"TTAGGG" * (L // 6 + 1))[:L]
            q = bytes(rng.integers(33, 33 + int(rng.integers(2, 60)), L).astype(np.uint8))
            out.append(b"@r%d\n%s\n+\n%s\n" % (len(out), seq, q))
        return b"".join(out)[:n]
    if kind == 1:      # repeats at every distance
        base = bytes(rng.integers(65, 91, 1 + int(rng.integers(1, 70000))).astype(np.uint8))
        return (base * (n // len(base) + 1))[:n]
    if kind == 2:      # skewed large alphabet: long Huffman codes
        p = 0.85 ** np.arange(200)
        return bytes((32 + rng.choice(200, n, p=p / p.sum())).astype(np.uint8))
    if kind == 3:      # binary
        return bytes(rng.integers(0, 256, n).astype(np.uint8))
    # mixture: text with islands of binary and long runs
    parts = []
    while sum(map(len, parts)) < n:
        parts.append(make_text(rng, int(rng.integers(0, 4)), int(rng.integers(1, 300000))))
        parts.append(bytes([int(rng.integers(0, 256))]) * int(rng.integers(0, 100000)))
    return b"".join(parts)[:n]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    L = lib()
    rng = np.random.default_rng(a.seed)
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]
    n_par = 0
    for case in range(a.cases):
        kind = int(rng.integers(0, 5))
        n = int(rng.choice([0, 1, 100, 70000, 1 << 20, 5 << 20, 12 << 20]))
        text = make_text(rng, kind, n) if n else b""
        members = int(rng.choice([1, 1, 1, 2, 3]))
        cuts = sorted(int(x) for x in rng.integers(0, len(text) + 1, members - 1))
        z = b""
        for lo, hi in zip([0] + cuts, cuts + [len(text)]):
            level, strat, mem = int(rng.integers(0, 10)), strategies[int(rng.integers(0, 5))], int(rng.integers(1, 10))
            co = zlib.compressobj(level, zlib.DEFLATED, 31, mem, strat)
            z += co.compress(text[lo:hi]) + co.flush()
        env = {"TPS_PGZ_FAST": str(int(rng.integers(0, 2))), "TPS_PGZ_BYTES": str(int(rng.integers(0, 2))),
               "TPS_PGZ_CLMUL": str(int(rng.integers(0, 2)))}
        os.environ.update(env)
        zbuf = np.frombuffer(z, np.uint8)
        threads = int(rng.choice([1, 2, 3, 8]))
        g = L.tps_pgz_open(zbuf.ctypes.data, len(z), threads)
        assert g, "open failed"
        L.tps_pgz_set_piece(g, int(rng.choice([1 << 16, 1 << 17, 1 << 19, 4 << 20])))
        cap = int(rng.choice([1 << 12, 1 << 16, 1 << 20, 1 << 24]))
        dst = np.empty(cap, np.uint8)
        got = []
        while True:
            r = L.tps_pgz_read(g, dst.ctypes.data, cap)
            if r < 0:
                raise SystemExit(f"case {case}: error {L.tps_pgz_error(g)} kind {kind} n {n} {env} threads {threads}")
            if r == 0:
                break
            got.append(dst[:r].tobytes())
        L.tps_pgz_close(g)
        if b"".join(got) != text:
            raise SystemExit(f"case {case}: MISMATCH kind {kind} n {n} members {members} {env} threads {threads} cap {cap}")
        n_par += threads > 1
        # the single-block entry (BGZF blocks): a raw deflate stream of <= 64 KiB of the same text, and a damaged one
        cut = text[:int(rng.integers(0, 65537))]
        co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, -15, int(rng.integers(1, 10)),
                              strategies[int(rng.integers(0, 5))])
        raw = co.compress(cut) + co.flush()
        rb = np.frombuffer(raw, np.uint8)
        out = np.empty(max(len(cut), 1), np.uint8)
        crc = C.c_uint32()
        rc = L.tps_pgz_inflate_block(rb.ctypes.data, len(raw), out.ctypes.data, len(cut), C.byref(crc))
        if rc != 0 or out[:len(cut)].tobytes() != cut or crc.value != (zlib.crc32(cut) & 0xffffffff):
            raise SystemExit(f"case {case}: single block rc {rc} kind {kind} len {len(cut)} {env}")
        if len(raw) > 40:
            bad = bytearray(raw)
            bad[len(bad) // 2] ^= 0x41
            bb = np.frombuffer(bytes(bad), np.uint8)
            rc = L.tps_pgz_inflate_block(bb.ctypes.data, len(bad), out.ctypes.data, len(cut), C.byref(crc))
            if rc == 0 and crc.value == (zlib.crc32(cut) & 0xffffffff) and out[:len(cut)].tobytes() != cut:
                raise SystemExit(f"case {case}: damaged block accepted with the right CRC but other bytes")
    print(f"{a.cases} cases identical ({n_par} with more than one thread)")


if __name__ == "__main__":
    main()
