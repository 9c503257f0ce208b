"""CPU tier: the C FASTQ/FASTA reader (csrc/tps_fastx.c) against the oracle's parser and against
Biopython's record conventions as the reference relies on them."""
import gzip
import hashlib
import os

import numpy as np
import pytest

from oracle import topsicle_oracle as orc
from tests.conftest import GOLD
from topsicle_b200 import fastx


def read_all(path, max_reads=1 << 12, max_bases=1 << 22, window=None, threads=4):
    out = []
    with fastx.FastxFile(path, threads=threads) as fx:
        if window:
            fx.set_window(window)
        bases = np.empty(max_bases, np.uint8)
        offs = np.empty(max_reads + 1, np.uint64)
        while True:
            b = fx.next_batch(bases, offs)
            if b is None:
                break
            assert b.first_read == len(out)
            for i in range(b.n_reads):
                out.append((b.read_id(i), b.sequence(i).decode(), b.title(i), b.record_text(i)))
            b.release()
    return out


@pytest.mark.parametrize("name", ["demo.fastq.gz", "edge.fastq", "edge.fasta"])
@pytest.mark.parametrize("caps", [(1 << 12, 1 << 22, None), (3, 1 << 20, 5000), (1, 1 << 20, 4096)])
def test_reader_matches_oracle_parser(name, caps):
    path = os.path.join(GOLD, name)
    got = read_all(path, *caps)
    assert [(a, b) for a, b, _, _ in got] == list(orc.read_fastx(path))


def test_subset_text_is_seqio_write_format(tmp_path):
    """record_text == what SeqIO.write emits: the reference's golden subset FASTQ is reproduced from the
    ids of the golden CSV (md5 of Topsicle_demo/result_justone/..._trc_over_0.7.fastq)."""
    path = os.path.join(GOLD, "demo.fastq.gz")
    import json
    case = json.load(open(os.path.join(GOLD, "demo_cli.json")))["cases"][0]
    ids = {ln.split(",")[3] for ln in case["csv"].split("\r\n")[1:] if ln}
    text = b"".join(r[3] for r in read_all(path) if r[0] in ids)
    assert hashlib.md5(text).hexdigest() == case["files"]["demo.fastq_trc_over_0.7.fastq"]


def test_fasta_conventions(tmp_path):
    p = tmp_path / "x.fasta"
    p.write_bytes(b">r1 desc here  \nACGT\r\nAC GT\n\nTT\n>r2\n>  r3\tz\nNNNN")
    got = read_all(str(p))
    assert [(g[0], g[1], g[2]) for g in got] == [("r1", "ACGTACGTTT", "r1 desc here"), ("r2", "", "r2"),
                                                ("r3", "NNNN", "  r3\tz")]
    assert got[0][3] == b">r1 desc here\nACGTACGTTT\n"
    long = tmp_path / "long.fa"
    seq = "ACGT" * 40
    long.write_text(">a\n" + seq + "\n")
    assert read_all(str(long))[0][3] == (">a\n" + seq[:60] + "\n" + seq[60:120] + "\n" + seq[120:] + "\n").encode()


def test_fastq_conventions_and_errors(tmp_path):
    p = tmp_path / "x.fq"
    p.write_bytes(b"@a b c\r\nACGT\r\n+a b c\r\n@@@@\r\n\n@b\nAC\n+\n!!")      # '@' quality, CRLF, no final newline
    got = read_all(str(p))
    assert [(g[0], g[1]) for g in got] == [("a", "ACGT"), ("b", "AC")]
    assert got[0][3] == b"@a b c\nACGT\n+\n@@@@\n"
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@a\nACGT\n+\n!!!\n")
    with pytest.raises(fastx.FastxError):
        read_all(str(bad))
    multi = tmp_path / "multi.fq"
    multi.write_bytes(b"@a\nACGT\nACGT\n+\n!!!!!!!!\n")
    with pytest.raises(fastx.FastxError):
        read_all(str(multi))
    other = tmp_path / "x.txt"
    other.write_text("hello\n")
    with pytest.raises(fastx.FastxError):
        fastx.FastxFile(str(other))
    assert fastx.sniff_format(str(other)) == 0
    with pytest.raises(fastx.FastxError):
        read_all(str(p), max_bases=3)                                        # read longer than the batch


def test_parallel_index_equals_sequential(tmp_path):
    """Big synthetic FASTQ (qualities full of '@' and '+'), plain and gzip: 8-thread segmented
    indexing, small windows and the sequential path all give the same records."""
    rng = np.random.default_rng(5)
    p = tmp_path / "big.fastq"
    want = []
    with open(p, "wb") as fh:
        for i in range(3000):
            L = int(rng.integers(1, 4000))
            seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L))
            qual = bytes(rng.choice(np.frombuffer(b"@+I!5", np.uint8), L))
            fh.write(b"@r%d extra\n%s\n+\n%s\n" % (i, seq, qual))
            want.append((f"r{i}", seq.decode()))
    for threads, window in [(1, None), (8, None), (8, 3 << 20), (3, 1 << 20)]:
        got = read_all(str(p), 1 << 12, 1 << 24, window, threads)
        assert [(a, b) for a, b, _, _ in got] == want, (threads, window)
    gz = tmp_path / "big.fastq.gz"
    with open(p, "rb") as src, gzip.open(gz, "wb", compresslevel=1) as dst:
        dst.write(src.read())
    got = read_all(str(gz), 500, 1 << 20, 1 << 20, 4)
    assert [(a, b) for a, b, _, _ in got] == want


def test_rawcount_csv_text_equals_pandas():
    import pandas as pd
    rng = np.random.default_rng(1)
    pats = ["AACC", "ACCC", "TTGG"]
    counts = rng.integers(1, 15, (57, 3)).astype(np.uint8)
    text = fastx.format_rawcount_csv(counts, 6, "reverse", pats)
    rows = [("reverse", w * 6, pats[p], int(counts[w, p])) for w in range(57) for p in range(3)]
    want = pd.DataFrame(rows, columns=["tail", "position", "pattern", "count"]).to_csv()
    assert text.decode() == want
    from topsicle_b200.allsteps import rawcount_frame
    assert rawcount_frame(counts, "reverse", 6, pats).to_csv() == want


# ----------------------------------------------------------------------------- span batches (one-pass reader)
def read_all_spans(path, max_reads=1 << 12, max_span=1 << 22, window=None, threads=4, two_pass=False):
    out = []
    with fastx.FastxFile(path, threads=threads) as fx:
        if window:
            fx.set_window(window)
        if two_pass:
            fx.set_two_pass()
        bases = np.full(max_span, ord("!"), np.uint8)
        starts = np.zeros(max_reads + 1, np.uint64)
        lens = np.zeros(max_reads, np.uint32)
        while True:
            b = fx.next_spans(bases, starts, lens)
            if b is None:
                break
            ends = starts[:b.n_reads] + lens[:b.n_reads]
            assert (ends[:-1] <= starts[1:b.n_reads]).all() and int(ends[-1]) <= b.span <= max_span
            assert b.n_bases == int(lens[:b.n_reads].sum())
            for i in range(b.n_reads):
                out.append((b.read_id(i), b.sequence(i).decode(), b.title(i), b.record_text(i)))
            b.release()
    return out


@pytest.mark.parametrize("name", ["demo.fastq.gz", "edge.fastq", "edge.fasta"])
@pytest.mark.parametrize("caps", [(1 << 12, 1 << 22, None), (3, 1 << 20, 5000), (1, 1 << 20, 4096)])
def test_span_reader_matches_oracle_parser(name, caps):
    path = os.path.join(GOLD, name)
    got = read_all_spans(path, *caps)
    assert [(a, b) for a, b, _, _ in got] == list(orc.read_fastx(path))
    assert got == read_all(path, *caps)


def test_one_pass_equals_two_pass_on_awkward_fastq(tmp_path):
    """CRLF, '@'/'+' in qualities, blank lines, trailing blanks (forces the validating fallback), a record
    bigger than the first window, no final newline -- plain and gzip, several thread counts."""
    rng = np.random.default_rng(17)
    p = tmp_path / "awk.fastq"
    want = []
    with open(p, "wb") as fh:
        for i in range(1500):
            L = int(rng.integers(0, 3000)) if i != 700 else 300000
            seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), L))
            qual = bytes(rng.choice(np.frombuffer(b"@+I!5", np.uint8), L))
            eol = b"\r\n" if i % 7 == 0 else b"\n"
            pad = b"      " if i % 97 == 0 else b""           # > 4 trailing blanks: two-pass path
            fh.write(b"@r%d some text%s%s%s%s+%s%s%s" % (i, eol, seq, pad, eol, eol, qual, eol if i < 1499 else b""))
            if i % 50 == 0:
                fh.write(b"\n")
            want.append((f"r{i}", seq.decode()))
    for threads, window in [(1, None), (8, None), (8, 2 << 20), (3, 1 << 20)]:
        got = read_all_spans(str(p), 1 << 12, 1 << 24, window, threads)
        assert [(a, b) for a, b, _, _ in got] == want, (threads, window)
        assert got == read_all_spans(str(p), 1 << 12, 1 << 24, window, threads, two_pass=True)
    gz = tmp_path / "awk.fastq.gz"
    with open(p, "rb") as src, gzip.open(gz, "wb", compresslevel=1) as dst:
        dst.write(src.read())
    got = read_all_spans(str(gz), 700, 1 << 20, None, 4)
    assert [(a, b) for a, b, _, _ in got] == want
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@a\nACGT\n+\n!!!!\n@b\nACGT\n+\n!!!\n")
    with pytest.raises(fastx.FastxError):
        read_all_spans(str(bad))


@pytest.mark.parametrize("end_len", [1000, 37, 50000])
@pytest.mark.parametrize("suffix", [".fastq.gz", ".fastq", ".fasta"])
def test_next_ends_head_tail_and_regions(tmp_path, end_len, suffix):
    """The ends-first reader: head + tail of every read back to back, real lengths aside, and the full
    sequence / the step-2 region still reachable through the record index (plain, gzip, multi-line FASTA)."""
    import gzip
    from oracle import topsicle_oracle as orc
    src = os.path.join(GOLD, "demo.fastq.gz")
    recs = list(orc.read_fastx(src))
    path = str(tmp_path / ("x" + suffix))
    if suffix == ".fastq.gz":
        path = src
    elif suffix == ".fastq":
        open(path, "wb").write(gzip.open(src, "rb").read())
    else:
        with open(path, "w") as fh:
            for rid, s in recs:
                fh.write(f">{rid} desc\n" + "".join(s[j:j + 70] + "\n" for j in range(0, len(s), 70)))
    bases = np.empty(1 << 22, np.uint8)
    starts, lens, tl = np.empty(64, np.uint64), np.empty(64, np.uint32), np.empty(64, np.uint32)
    got = []
    with fastx.FastxFile(path, threads=3) as fx:
        while True:
            b = fx.next_ends(bases, starts, lens, tl, end_len, raw_cap=250_000)
            if b is None:
                break
            assert b.n_bases == int(tl[:b.n_reads].sum())
            for i in range(b.n_reads):
                a = int(starts[i])
                got.append((b.read_id(i), int(tl[i]), bases[a:a + int(lens[i])].tobytes(), b.sequence(i),
                            b.region(i, 0, 20000), b.region(i, 1, 777)))
            b.release()
    assert len(got) == len(recs)
    for (rid, s), g in zip(recs, got):
        s = s.encode()
        assert g[0] == rid and g[1] == len(s)
        assert g[2] == (s if len(s) <= 2 * end_len else s[:end_len] + s[-end_len:])
        assert g[3] == s and g[4] == s[:20000] and g[5] == s[-777:]


def _write_bgzf(path, data, block=0xff00):
    """bgzip's format: gzip members of <= 64 KiB with the 'BC' extra field (compressed block size - 1),
    closed by the empty end-of-file block."""
    import struct
    import zlib
    with open(path, "wb") as fh:
        for a in list(range(0, len(data), block)) + [None]:
            chunk = b"" if a is None else data[a:a + block]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            body = co.compress(chunk) + co.flush()
            bsize = 12 + 6 + len(body) + 8
            fh.write(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) +
                     b"BC" + struct.pack("<HH", 2, bsize - 1) + body +
                     struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))


@pytest.mark.parametrize("threads", [1, 4, -4])
def test_bgzf_blocks_inflated_in_parallel(tmp_path, threads, monkeypatch):
    """A BGZF (.gz) file gives the same records as the plain-gzip file, through every reader entry point;
    Python's gzip (what the reference uses) reads it too; a corrupt block is an error, not a short file."""
    if threads < 0:                                     # the blocks through zlib instead of the repo's byte decoder
        monkeypatch.setenv("TPS_FX_BGZF_ZLIB", "1")
        threads = -threads
    src = os.path.join(GOLD, "demo.fastq.gz")
    text = gzip.open(src, "rb").read() * 3              # ~5 MB, 80 blocks
    path = str(tmp_path / "demo3.fastq.gz")
    _write_bgzf(path, text)
    assert gzip.open(path, "rb").read() == text
    want = [(rid, s) for _ in range(3) for rid, s in orc.read_fastx(src)]
    bases = np.empty(1 << 23, np.uint8)
    offsets = np.empty(4096, np.uint64)
    starts, lens, tl = np.empty(4096, np.uint64), np.empty(4096, np.uint32), np.empty(4096, np.uint32)
    for mode in ("next", "spans", "ends"):
        got = []
        with fastx.FastxFile(path, threads=threads) as fx:
            assert fx.format_name == "fastq"
            fx.set_window(700_000)
            while True:
                if mode == "next":
                    b = fx.next_batch(bases, offsets, max_bases=400_000)
                elif mode == "spans":
                    b = fx.next_spans(bases, starts, lens, max_span=300_000)
                else:
                    b = fx.next_ends(bases, starts, lens, tl, 1000, raw_cap=500_000)
                if b is None:
                    break
                for i in range(b.n_reads):
                    got.append((b.read_id(i), b.sequence(i).decode()))
                b.release()
        assert got == want, mode
    raw = bytearray(open(path, "rb").read())
    raw[len(raw) // 2] ^= 0x55
    bad = str(tmp_path / "bad.fastq.gz")
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(fastx.FastxError):
        with fastx.FastxFile(bad, threads=threads) as fx:
            while fx.next_batch(bases, offsets) is not None:
                pass


def test_record_text_fast_path_equals_general_form(tmp_path):
    """record_text slices the raw text when the record already has SeqIO.write's shape, and rebuilds it when the
    '+' line repeats the title, the title has trailing blanks or the lines end in CRLF."""
    recs = [("r1 desc", "ACGTNNacgt", "IIIIIIIIII"), ("r2", "TTAGGG" * 5, "#" * 30), ("r3  ", "AC", "!!")]
    variants = {
        "plain": "".join(f"@{t}\n{s}\n+\n{q}\n" for t, s, q in recs),
        "plus_title": "".join(f"@{t}\n{s}\n+{t}\n{q}\n" for t, s, q in recs),
        "crlf": "".join(f"@{t}\r\n{s}\r\n+\r\n{q}\r\n" for t, s, q in recs),
        "no_final_newline": "".join(f"@{t}\n{s}\n+\n{q}\n" for t, s, q in recs)[:-1],
    }
    want = ["@{}\n{}\n+\n{}\n".format(t.rstrip(), s, q).encode() for t, s, q in recs]
    for name, text in variants.items():
        path = str(tmp_path / f"{name}.fastq")
        open(path, "w", newline="").write(text)
        bases, offsets = np.empty(4096, np.uint8), np.empty(16, np.uint64)
        with fastx.FastxFile(path, threads=1) as fx:
            b = fx.next_batch(bases, offsets)
            got = [b.record_text(i) for i in range(b.n_reads)]
            b.release()
        assert got == want, name


def test_next_ends_on_awkward_fastq(tmp_path):
    """The ends-first reader on the same awkward input (CRLF, blank lines, trailing blanks, a record larger than
    the batch's text window, no final newline): ends, real lengths, full sequences and regions, for several
    thread counts, window sizes and tiny read / base capacities; plain and gzip."""
    rng = np.random.default_rng(23)
    p = tmp_path / "awk.fastq"
    want = []
    with open(p, "wb") as fh:
        for i in range(1200):
            L = int(rng.integers(0, 3000)) if i != 500 else 300000
            seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), L))
            qual = bytes(rng.choice(np.frombuffer(b"@+I!5", np.uint8), L))
            eol = b"\r\n" if i % 7 == 0 else b"\n"
            pad = b"      " if i % 97 == 0 else b""
            fh.write(b"@r%d some text%s%s%s%s+%s%s%s" % (i, eol, seq, pad, eol, eol, qual, eol if i < 1199 else b""))
            if i % 50 == 0:
                fh.write(b"\n")
            want.append((f"r{i}", seq))
    gz = tmp_path / "awk.fastq.gz"
    with open(p, "rb") as src, gzip.open(gz, "wb", compresslevel=1) as dst:
        dst.write(src.read())
    H = 400
    for path, threads, raw_cap, max_reads, max_bases in [(p, 1, 1 << 30, 4096, 1 << 22), (p, 8, 200_000, 4096, 1 << 22),
                                                        (p, 3, 50_000, 7, 1 << 22), (p, 4, 1 << 20, 4096, 3000),
                                                        (gz, 4, 300_000, 100, 1 << 22)]:
        bases = np.empty(1 << 22, np.uint8)
        starts, lens, tl = np.empty(4096, np.uint64), np.empty(4096, np.uint32), np.empty(4096, np.uint32)
        got = []
        with fastx.FastxFile(str(path), threads=threads) as fx:
            while True:
                b = fx.next_ends(bases, starts, lens, tl, H, raw_cap=raw_cap, max_reads=max_reads, max_bases=max_bases)
                if b is None:
                    break
                assert b.n_reads <= max_reads and b.span <= max_bases
                for i in range(b.n_reads):
                    a = int(starts[i])
                    got.append((b.read_id(i), int(tl[i]), bases[a:a + int(lens[i])].tobytes(), b.sequence(i),
                                b.region(i, 0, 1000), b.region(i, 1, 1000)))
                b.release()
        assert len(got) == len(want), (threads, raw_cap)
        for (rid, s), g in zip(want, got):
            assert g[0] == rid and g[1] == len(s), rid
            assert g[2] == (s if len(s) <= 2 * H else s[:H] + s[-H:]), rid
            assert g[3] == s and g[4] == s[:1000] and g[5] == s[-1000:], rid


def test_records_text_equals_record_text(tmp_path):
    """The batched SeqIO.write text (one C call) equals the per-record form on FASTQ (plain / CRLF / '+title' /
    trailing blanks), multi-line FASTA and gzip, through whole-read and ends batches."""
    rng = np.random.default_rng(5)
    fq, fa = tmp_path / "a.fastq", tmp_path / "a.fasta"
    with open(fq, "wb") as f1, open(fa, "wb") as f2:
        for i in range(300):
            L = int(rng.integers(0, 400))
            seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), L))
            qual = bytes(rng.choice(np.frombuffer(b"@+I!5", np.uint8), L))
            eol = b"\r\n" if i % 5 == 0 else b"\n"
            plus = b"+r%d" % i if i % 4 == 0 else b"+"
            pad = b"   " if i % 9 == 0 else b""
            f1.write(b"@r%d some text%s%s%s%s%s%s%s%s" % (i, pad, eol, seq, eol, plus, eol, qual, eol))
            f2.write(b">r%d t%s\n" % (i, pad) + b"".join(seq[j:j + 70] + b"\n" for j in range(0, L, 70)))
    gz = tmp_path / "a.fastq.gz"
    with open(fq, "rb") as src, gzip.open(gz, "wb") as dst:
        dst.write(src.read())
    bases, offsets = np.empty(1 << 20, np.uint8), np.empty(4096, np.uint64)
    starts, lens, tl = np.empty(4096, np.uint64), np.empty(4096, np.uint32), np.empty(4096, np.uint32)
    for path in (fq, fa, gz):
        for mode in ("next", "ends"):
            with fastx.FastxFile(str(path), threads=2) as fx:
                b = fx.next_batch(bases, offsets) if mode == "next" else fx.next_ends(bases, starts, lens, tl, 50)
                want = [b.record_text(i) for i in range(b.n_reads)]
                got = [bytes(t) for t in b.records_text(range(b.n_reads))]
                some = [bytes(t) for t in b.records_text([7, 3, 299])]
                ids = b.read_ids([7, 3, 299])
                b.release()
            assert b.n_reads == 300 and got == want, (path, mode)
            assert some == [want[7], want[3], want[299]] and ids == ["r7", "r3", "r299"]


# ------------------------------------------------------------------ plain gzip: parallel inflate (csrc/tps_pgz.c)
def _synthetic_fastq(n_reads=1000, seed=5):
    """FASTQ text with realistic (poorly compressible) quality lines."""
    from topsicle_b200 import synth
    bases, off, _ = synth.generate(synth.CONFIGS[2], 0, n_reads)
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_reads):
        s = bases[int(off[i]):int(off[i + 1])].tobytes()
        out.append(b"@r%d some description\n%s\n+\n%s\n" % (i, s, rng.integers(35, 74, len(s)).astype(np.uint8).tobytes()))
    return b"".join(out)


def _read_all(path, threads=8, **env):
    from topsicle_b200 import fastx
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        with fastx.FastxFile(path, threads=threads) as fx:
            bases = np.empty(1 << 26, dtype=np.uint8)
            offsets = np.empty((1 << 15) + 1, dtype=np.uint64)
            recs = []
            while True:
                b = fx.next_batch(bases, offsets)
                if b is None:
                    break
                recs += [(b.read_id(i), b.sequence(i), b.quality(i)) for i in range(b.n_reads)]
                b.release()
            return recs, fx.inflate_stats()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("level", [1, 6, 9])
def test_plain_gzip_parallel_inflate_equals_zlib(tmp_path, level):
    """A `.fastq.gz` written by gzip (one deflate stream) is inflated by all parser threads -- block starts guessed,
    unknown windows decoded symbolically, pieces chained, CRC-32 checked -- and gives exactly the records zlib
    gives (TPS_FX_NO_PGZ=1) and the plain file gives.  Reference: gzip.open in unzip_file, allsteps.py:142-146."""
    import gzip
    text = _synthetic_fastq()
    plain, gz = tmp_path / "r.fastq", tmp_path / "r.fastq.gz"
    plain.write_bytes(text)
    gz.write_bytes(gzip.compress(text, compresslevel=level))
    want, st0 = _read_all(str(plain))
    assert st0["text_bytes"] == 0
    got, st = _read_all(str(gz), TPS_PGZ_PIECE=1 << 18)
    assert got == want and len(got) == 1000
    assert st["text_bytes"] == len(text) and st["members"] == 1 and st["chain_breaks"] == 0
    assert st["parallel_text_bytes"] > 0.7 * len(text) and st["segments"] > 3 * st["stretches"]
    zl, st_z = _read_all(str(gz), TPS_FX_NO_PGZ=1)
    assert zl == want and st_z["text_bytes"] == 0
    one, st1 = _read_all(str(gz), threads=1)
    assert one == want and st1["parallel_text_bytes"] == 0
    # the plain decoder (one symbol per table lookup) and zlib's crc32 in place of the folded one
    slow, st_s = _read_all(str(gz), TPS_PGZ_PIECE=1 << 18, TPS_PGZ_FAST=0, TPS_PGZ_CLMUL=0)
    assert slow == want and st_s["text_bytes"] == len(text)
    # pieces that stay symbolic to their end (no change-over to byte output), and large pieces that all change over
    sym, st_y = _read_all(str(gz), TPS_PGZ_PIECE=1 << 18, TPS_PGZ_BYTES=0)
    assert sym == want and st_y["text_bytes"] == len(text)
    big, st_b = _read_all(str(gz), TPS_PGZ_PIECE=1 << 21)
    assert big == want and st_b["text_bytes"] == len(text) and st_b["chain_breaks"] == 0


def test_plain_gzip_long_codes_and_fixed_blocks(tmp_path):
    """Streams that leave the decoder's direct tables: a skewed 94-letter quality alphabet (literal codes longer
    than 12 bits, the general path inside the fast loop), long repeats (length 258, distance codes with 13 extra
    bits), repeats at distances below 8, and a file small enough for fixed-Huffman blocks."""
    import gzip
    rng = np.random.default_rng(41)
    p_q = 0.82 ** np.arange(94)
    p_q /= p_q.sum()
    recs = []
    for i in range(1500):
        L = int(rng.integers(200, 9000))
        kind = i % 5
        if kind == 0:
            seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L))
        elif kind == 1:
            seq = (b"TTAGGG" * (L // 6 + 1))[:L]
        elif kind == 2:
            seq = (b"A" * L)
        elif kind == 3:
            seq = (b"AC" * (L // 2 + 1))[:L]
        else:
            seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgtn", np.uint8), L))
        qual = bytes((33 + rng.choice(94, L, p=p_q)).astype(np.uint8)).replace(b"@", b"A")
        recs.append(b"@q%d\n%s\n+\n%s\n" % (i, seq, qual))
    text = b"".join(recs)
    plain = tmp_path / "r.fastq"
    plain.write_bytes(text)
    want, _ = _read_all(str(plain))
    for level in (1, 6, 9):
        gz = tmp_path / f"r{level}.fastq.gz"
        gz.write_bytes(gzip.compress(text, compresslevel=level))
        got, st = _read_all(str(gz), TPS_PGZ_PIECE=1 << 17)
        assert got == want and st["text_bytes"] == len(text) and st["chain_breaks"] == 0, level
        assert st["parallel_text_bytes"] > 0.5 * len(text)
        slow, _ = _read_all(str(gz), TPS_PGZ_PIECE=1 << 17, TPS_PGZ_FAST=0)
        assert slow == want
    tiny = b"@t1\nACGTACGTAC\n+\nIIIIIIIIII\n@t2\nTTAGGGTTAGGG\n+\nIIIIIIIIIIII\n"
    tz = tmp_path / "tiny.fastq.gz"
    tz.write_bytes(gzip.compress(tiny, 6))
    got, _ = _read_all(str(tz))
    assert [(r[0], r[1]) for r in got] == [("t1", b"ACGTACGTAC"), ("t2", b"TTAGGGTTAGGG")]


def test_plain_gzip_members_stored_blocks_and_damage(tmp_path):
    """Several members (`cat a.gz b.gz`), an empty member, stored blocks (level 0), binary data in front of the text
    (no block start is found in it: that stretch is decoded by one thread), a flipped byte (CRC-32 mismatch) and a
    truncated file (both reported, never silently accepted)."""
    import gzip
    from topsicle_b200 import fastx
    text = _synthetic_fastq(900, seed=9)
    cut = [text.rfind(b"\n@r", 0, len(text) // 3) + 1, text.rfind(b"\n@r", 0, 2 * len(text) // 3) + 1]
    multi = (gzip.compress(text[:cut[0]], 6) + gzip.compress(b"") + gzip.compress(text[cut[0]:cut[1]], 0)
             + gzip.compress(text[cut[1]:], 9))
    p = tmp_path / "multi.fastq.gz"
    p.write_bytes(multi)
    plain = tmp_path / "plain.fastq"
    plain.write_bytes(text)
    want, _ = _read_all(str(plain))
    got, st = _read_all(str(p), TPS_PGZ_PIECE=1 << 17)
    assert got == want and st["members"] == 4 and st["text_bytes"] == len(text)
    z = bytearray(gzip.compress(text, 6))
    z[len(z) // 2] ^= 0x5A
    bad = tmp_path / "bad.fastq.gz"
    bad.write_bytes(bytes(z))
    with pytest.raises(fastx.FastxError):
        _read_all(str(bad), TPS_PGZ_PIECE=1 << 17)
    trunc = tmp_path / "trunc.fastq.gz"
    trunc.write_bytes(gzip.compress(text, 6)[:len(z) // 2])
    with pytest.raises(fastx.FastxError):
        _read_all(str(trunc), TPS_PGZ_PIECE=1 << 17)
