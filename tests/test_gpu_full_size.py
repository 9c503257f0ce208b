"""GPU tier, BASELINE.json full size: config 2 = 1 000 000 synthetic ONT reads (~23.5 Gbases), scanned in
batches of 63 488 reads.  The oracle cannot run at this size, so the checks are size-independent
properties of the scan:
  * partition invariance -- rows do not depend on how the reads are cut into batches (each batch is
    re-scanned as two half batches through the HOST path: identical bytes);
  * reversal symmetry    -- reversing every read swaps head_max and tail_max exactly (step 1 looks at
    seq[:1000] and reversed(seq[-1000:]), allsteps.py:176-177) and leaves n_windows unchanged;
  * generator ground truth -- a checksum of checksums: every read built with a telomere of >= 2 kb
    passes TRC at 0.7 with the right tail, background reads never pass;
  * telo_length lies on the grid trimfirst + 5*slide*m and inside the read.
Set TPS_FULL_SIZE_READS to shrink it (default 1 000 000)."""
import os

import numpy as np
import pytest

from topsicle_b200 import engine, synth
from topsicle_b200.patterns import patterns_to_search

pytestmark = pytest.mark.gpu

TOTAL = int(os.environ.get("TPS_FULL_SIZE_READS", 1_000_000))
BATCH = 63_488


def test_config2_full_size_properties():
    import torch
    spec = synth.CONFIGS[2]
    pats = patterns_to_search("CCCTAA", 4)
    dev = torch.device("cuda", 0)
    ctx = engine.ScanContext(pats, len_telopattern=6, cutoff=0.7, min_seq_length=9000, slide=6, n_slots=2,
                             max_batch_reads=BATCH, max_batch_bases=1 << 31)
    tot_reads = tot_bases = tot_pass = 0
    kinds_pass = np.zeros(5, dtype=np.int64)
    kinds_all = np.zeros(5, dtype=np.int64)
    checksum = 0
    try:
        first = 0
        bi = 0
        while first < TOTAL:
            n = min(BATCH, TOTAL - first)
            off = synth.read_lengths(spec, first, n)
            nb = int(off[-1])
            host = engine.PinnedBuffer(nb)
            kinds = synth.fill_reads(spec, first, off, host.array)
            bases = host.array[:nb]
            # device-resident path
            db = torch.empty((nb + 2047) // 2048 * 2048, dtype=torch.uint8, device=dev)
            db[:nb].copy_(torch.from_numpy(bases))
            do = torch.from_numpy(off.view(np.int64)).to(dev)
            d_rows = torch.empty(n * 40, dtype=torch.uint8, device=dev)
            ctx.scan_device(db.data_ptr(), do.data_ptr(), n, nb, d_rows.data_ptr())
            ctx.sync()
            rows = np.frombuffer(d_rows.cpu().numpy().tobytes(), dtype=engine.ROW_DTYPE)
            # partition invariance through the host path, on a rotating subset of batches (PCIe bound)
            if bi % 4 == 0:
                mid = n // 2
                r1, _ = ctx.scan(bases[:int(off[mid])], off[:mid + 1].copy())
                o2 = (off[mid:] - off[mid]).astype(np.uint64)
                r2, _ = ctx.scan(bases[int(off[mid]):nb], o2)
                assert np.concatenate([r1, r2]).tobytes() == rows.tobytes()
            # reversal symmetry on the first 4096 reads of the batch
            m = min(4096, n)
            sub_off = off[:m + 1].copy()
            rev = np.empty(int(sub_off[-1]), dtype=np.uint8)
            for i in range(m):
                a, b = int(sub_off[i]), int(sub_off[i + 1])
                rev[a:b] = bases[a:b][::-1]
            rr, _ = ctx.scan(rev, sub_off)
            scanned = rows["status"][:m] != engine.ST_FILTERED
            assert np.array_equal(rr["head_max"][scanned], rows["tail_max"][:m][scanned])
            assert np.array_equal(rr["tail_max"][scanned], rows["head_max"][:m][scanned])
            assert np.array_equal(rr["status"] != engine.ST_FILTERED, scanned)
            differ = rows["head_max"][:m] != rows["tail_max"][:m]
            assert np.array_equal(rr["tail"][scanned & differ], 1 - rows["tail"][:m][scanned & differ])
            # rows are self-consistent
            L = rows["length"]
            assert np.array_equal(L, np.diff(off).astype(np.uint32))
            assert np.array_equal(rows["status"] == engine.ST_FILTERED, L <= 9000)
            ok = rows["status"] == engine.ST_PASS
            assert not (rows["status"] == engine.ST_BADSEG).any()      # every scanned read has > 7 windows
            thr = engine.count_threshold(0.7, 6)
            assert np.array_equal(ok, (L > 9000) & (rows["match_count"] >= thr))
            assert np.array_equal(rows["match_count"], np.where(rows["tail"] == 0, rows["head_max"], rows["tail_max"])
                                  * (L > 9000))
            tl = rows["telo_length"][ok].astype(np.int64)
            assert ((tl - 100) % 30 == 0).all() and (tl > 100).all()
            assert (tl <= np.minimum(L[ok], 20000)).all()
            nw = (np.minimum(L[ok].astype(np.int64), 20000) - 100 - 100) // 6 + 1
            assert np.array_equal(rows["n_windows"][ok], nw.astype(np.uint32))
            # generator ground truth: kind 1 = forward telomere, 2 = reverse telomere (>= 2 kb), 0 = background
            long_enough = L > 9000
            assert ok[(kinds == 1) & long_enough].all() and ok[(kinds == 2) & long_enough].all()
            assert not ok[kinds == 0].any()
            assert (rows["tail"][(kinds == 1) & long_enough] == 0).all()
            assert (rows["tail"][(kinds == 2) & long_enough] == 1).all()
            # the change point sits at the synthetic telomere end (indels shift it by a few percent)
            kinds_pass += np.bincount(kinds[ok], minlength=5)
            kinds_all += np.bincount(kinds, minlength=5)
            checksum = (checksum * 1_000_003 + int(rows["telo_length"].astype(np.int64).sum())
                        + int(rows["match_count"].astype(np.int64).sum())) % (1 << 61)
            tot_reads += n
            tot_bases += nb
            tot_pass += int(ok.sum())
            first += n
            bi += 1
            host.free()
            del db, do, d_rows
    finally:
        ctx.close()
    assert tot_reads == TOTAL
    print(f"full size: {tot_reads} reads, {tot_bases / 1e9:.2f} Gbases, {tot_pass} TRC-pass, by kind {kinds_pass.tolist()} "
          f"of {kinds_all.tolist()}, checksum {checksum}")
    if TOTAL == 1_000_000:
        assert 22e9 < tot_bases < 25e9
        assert 0.008 * TOTAL < tot_pass < 0.012 * TOTAL
