"""TEST INFRASTRUCTURE ONLY -- importable placeholder.

`allsteps.py:16` imports `FastqGeneralIterator` but never calls it.
"""


def FastqGeneralIterator(handle):
    from . import _parse_fastq
    for rec in _parse_fastq(handle):
        yield rec.description, str(rec.seq), rec.qual
