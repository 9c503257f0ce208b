"""CPU tier: differential fuzz of the parallel gzip inflater through its C ABI (tools/pgz_fuzz.py): random texts under
every zlib level / strategy / memLevel, one or several members, random thread counts, piece and read sizes, every
decoder variant -- the inflated bytes must equal the input.  (Reference: gzip.open in unzip_file, allsteps.py:142-146.)"""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inflater_fuzz_small():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "pgz_fuzz.py"), "--cases", "40", "--seed", "7"],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "40 cases identical" in r.stdout
