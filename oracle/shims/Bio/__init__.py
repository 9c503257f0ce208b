"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the `Bio` (biopython) package.

biopython is not installed in the build container and there is no network.  The
unmodified reference (`/root/reference/Topsicle/allsteps.py:14-16`) does
`import Bio; from Bio import SeqIO; from Bio.SeqIO.QualityIO import
FastqGeneralIterator`.  This shim provides exactly the surface the reference
touches so that `oracle/make_golden.py` can run the reference's own code and
emit golden vectors.  It carries no Topsicle arithmetic.
"""
__version__ = "0.0-shim"
