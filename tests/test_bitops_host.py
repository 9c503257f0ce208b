"""CPU: the bit-level helpers the kernels are built from (topsicle_b200/csrc/tps_bitops.h),
compiled for the host and checked against naive Python / the oracle."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import topsicle_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "csrc", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libbitops_host.so")
    src = os.path.join(HERE, "csrc", "bitops_host.cpp")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    L = C.CDLL(so)
    for f in ("t_pack16", "t_exact_mask16", "t_exact_mask16_simd", "t_code_at", "t_linear_planes", "t_ascii_code", "t_greedy_count",
              "t_range_popcount"):
        getattr(L, f).restype = C.c_uint32
    L.t_change_point.restype = C.c_int32
    return L


CODE = {"A": 0, "C": 1, "T": 2, "G": 3}


def test_pack16_all_bytes(lib):
    rnd = random.Random(1)
    valid = set(b"ACGTacgt")
    for trial in range(4000):
        if trial < 256:
            b = bytes([trial] * 16)
        elif trial < 2000:
            b = bytes(rnd.choice(b"ACGTacgt") for _ in range(16))
        else:
            b = bytes(rnd.choice(b"ACGTacgtNnRYKM-*\n\r@+I!") if rnd.random() < 0.9 else rnd.randrange(256)
                      for _ in range(16))
        bad = C.c_uint32()
        u = lib.t_pack16(b, C.byref(bad))
        want_bad = any(x not in valid for x in b)
        assert (bad.value != 0) == want_bad, b
        mask = lib.t_exact_mask16(b)
        assert lib.t_exact_mask16_simd(b) == mask, b
        for g in range(16):
            assert ((mask >> g) & 1) == (b[g] in valid)
            if b[g] in valid:
                assert lib.t_code_at(u, g) == CODE[chr(b[g]).upper()]
                assert lib.t_ascii_code(b[g]) == CODE[chr(b[g]).upper()]
            else:
                assert lib.t_ascii_code(b[g]) == 0xFF
        y = lib.t_linear_planes(u)
        for g in range(16):
            code = lib.t_code_at(u, g)
            assert ((y >> g) & 1) == (code & 1)
            assert ((y >> (16 + g)) & 1) == (code >> 1)


def _mask_words(text, lit):
    n = len(text)
    bits = np.zeros((n + 63) // 32 + 1, dtype=np.uint32)
    for i in range(n - len(lit) + 1):
        if text[i:i + len(lit)] == lit:
            bits[i >> 5] |= np.uint32(1 << (i & 31))
    return bits


def test_greedy_and_popcount(lib):
    rnd = random.Random(2)
    for _ in range(400):
        n = rnd.randint(1, 300)
        text = "".join(rnd.choice(rnd.choice(["ACGT", "AC", "A", "CTA"])) for _ in range(n))
        for lit in ("AA", "CCC", "CTAAC", "ACA", "CCCTAA", "A", "ACAC", "TAACCCT"):
            m = _mask_words(text, lit)
            mp = m.ctypes.data_as(C.POINTER(C.c_uint32))
            k = len(lit)
            for _ in range(6):
                a = rnd.randint(0, n)
                b = rnd.randint(a, n)
                sub = text[a:b]
                to = b - k
                got = lib.t_greedy_count(mp, a, to, k) if to >= a else 0
                assert got == orc.greedy_count(sub, lit), (text, lit, a, b)
                allocc = sum(1 for i in range(a, to + 1) if text[i:i + k] == lit)
                assert lib.t_range_popcount(mp, a, to) == allocc


def test_change_point_exact(lib):
    rnd = np.random.default_rng(3)
    for trial in range(300):
        n = int(rnd.integers(1, 400))
        kind = trial % 4
        if kind == 0:
            cw = rnd.integers(12, 120, n)
        elif kind == 1:
            b = int(rnd.integers(0, n + 1))
            cw = np.concatenate([rnd.integers(80, 100, b), rnd.integers(12, 20, n - b)])
        elif kind == 2:
            cw = np.full(n, 37)
        else:
            cw = rnd.integers(12, 14, n)
        cw = cw.astype(np.uint32)
        got = lib.t_change_point(cw.ctypes.data_as(C.POINTER(C.c_uint32)), n)
        if n < 7:
            assert got == -1
            with pytest.raises(ValueError):
                orc.change_point_exact(cw)
        else:
            assert got == orc.change_point_exact(cw)


def test_change_point_exact_long_signals(lib):
    """Full-size signals (nW up to 6617 at W50/s3, c_w up to 14*24) incl. piecewise-constant ones with
    many exactly tied gains: the float64-filtered comparator must still return the exact argmax."""
    rnd = np.random.default_rng(9)
    for trial in range(12):
        n = int(rnd.integers(2500, 6700))
        if trial % 3 == 0:
            b = int(rnd.integers(100, n - 100))
            cw = np.concatenate([np.full(b, 168), np.full(n - b, 14)])           # ideal telomere step
        elif trial % 3 == 1:
            cw = np.repeat(rnd.integers(12, 336, n // 50 + 1), 50)[:n]             # plateaus: tied gains
        else:
            cw = rnd.integers(12, 336, n)
        cw = cw.astype(np.uint32)
        got = lib.t_change_point(cw.ctypes.data_as(C.POINTER(C.c_uint32)), n)
        assert got == orc.change_point_exact(cw), (trial, n)
