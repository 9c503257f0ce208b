"""TEST INFRASTRUCTURE ONLY -- `Bio.SeqIO` stand-in (parse/write, 4-line FASTQ and FASTA).

Surface used by the reference:
  * `SeqIO.parse(handle, "fastq"|"fasta")`   allsteps.py:145, main.py:84
  * `SeqIO.write(record, handle, fmt)`       main.py:86
  * record `.id` (title up to first whitespace), `.seq` (sliceable, `.upper()`,
    `len()`, `str()`), `len(record)`         allsteps.py:175-177, 258-271
Formatting of `write` follows Biopython: FASTQ `@{description}\n{seq}\n+\n{qual}\n`;
FASTA `>{description}\n` followed by the sequence wrapped at 60 columns.
"""


class Seq(str):
    """A `str` whose slices / upper() keep the type (enough for the reference)."""

    def __getitem__(self, key):
        out = str.__getitem__(self, key)
        return Seq(out) if isinstance(key, slice) else out

    def upper(self):
        return Seq(str.upper(self))


class SeqRecord:
    __slots__ = ("id", "name", "description", "seq", "qual")

    def __init__(self, title, seq, qual=None):
        self.description = title
        self.id = title.split(None, 1)[0] if title.split() else ""
        self.name = self.id
        self.seq = Seq(seq)
        self.qual = qual

    def __len__(self):
        return len(self.seq)


def _parse_fastq(handle):
    while True:
        title = handle.readline()
        if not title:
            return
        if title in ("\n", "\r\n"):
            continue
        if not title.startswith("@"):
            raise ValueError("Records in Fastq files should start with '@' character")
        seq = handle.readline().rstrip()
        plus = handle.readline()
        if not plus.startswith("+"):
            raise ValueError("multi-line FASTQ is not supported by the shim")
        qual = handle.readline().rstrip()
        if len(qual) != len(seq):
            raise ValueError("Lengths of sequence and quality values differs")
        yield SeqRecord(title[1:].rstrip(), seq, qual)


def _parse_fasta(handle):
    title, chunks = None, []
    for line in handle:
        if line.startswith(">"):
            if title is not None:
                yield SeqRecord(title, "".join(chunks))
            title, chunks = line[1:].rstrip(), []
        elif title is not None:
            chunks.append(line.strip())
    if title is not None:
        yield SeqRecord(title, "".join(chunks))


def parse(handle, fmt):
    if fmt == "fastq":
        return _parse_fastq(handle)
    if fmt == "fasta":
        return _parse_fasta(handle)
    raise ValueError(f"Unknown format '{fmt}'")


def write(records, handle, fmt):
    if isinstance(records, SeqRecord):
        records = [records]
    n = 0
    for r in records:
        if fmt == "fastq":
            if r.qual is None:
                raise ValueError("No suitable quality scores found")
            handle.write(f"@{r.description}\n{r.seq}\n+\n{r.qual}\n")
        elif fmt == "fasta":
            handle.write(f">{r.description}\n")
            s = str(r.seq)
            for i in range(0, len(s), 60):
                handle.write(s[i:i + 60] + "\n")
        else:
            raise ValueError(f"Unknown format '{fmt}'")
        n += 1
    return n
