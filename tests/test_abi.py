"""CPU tier: the C-ABI library loads, exports every function `include/topsicle_b200.h` declares,
agrees with the ctypes mirror on struct sizes, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import REPO
from topsicle_b200 import engine, fastx


def declared_functions(header="topsicle_b200.h"):
    text = open(os.path.join(REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(tps_\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("tps_create", "tps_destroy", "tps_submit", "tps_wait", "tps_batch_info", "tps_scan_device",
                 "tps_sync", "tps_get_timings", "tps_get_timeline", "tps_elapsed_between", "tps_last_error", "tps_alloc_pinned", "tps_free_pinned",
                 "tps_abi_version", "tps_build_info", "tps_kernel_launches", "tps_debug_copy", "tps_device_count",
                 "tps_follow_scan", "tps_submit_spans", "tps_submit_ends", "tps_submit_regions", "tps_submit_shared", "tps_scan_device_slot"):
        assert must in names, must


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    for name in declared_functions():
        assert getattr(lib, name) is not None, name
    assert lib.tps_abi_version() == 1
    assert b"sm_100a" in lib.tps_build_info()


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof as gcc sees the header == the ctypes / numpy mirrors."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "topsicle_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(tps_params), sizeof(tps_row),'
                   'offsetof(tps_params, flags), offsetof(tps_params, max_batch_bases),'
                   'offsetof(tps_row, telo_length), offsetof(tps_row, rawcount_offset));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    P = engine.TpsParams
    want = [C.sizeof(P), engine.ROW_DTYPE.itemsize, P.flags.offset, P.max_batch_bases.offset,
            engine.ROW_DTYPE.fields["telo_length"][1], engine.ROW_DTYPE.fields["rawcount_offset"][1]]
    assert got == want
    assert fastx.REC_DTYPE.itemsize == 48


def test_no_cpu_fallback():
    """Without a CUDA device the product fails loudly; with one, a bad struct_size is rejected."""
    import torch
    lib = engine.load_library()
    p = engine.TpsParams()
    h = C.c_void_p()
    if not torch.cuda.is_available():
        with pytest.raises(engine.TpsError) as e:
            engine.ScanContext(["CCCTA"])
        assert e.value.code == -6 and "no CPU fallback" in str(e.value)
    p.struct_size = 1
    assert lib.tps_create(C.byref(h), 0, C.byref(p)) == -1
    assert b"ABI mismatch" in lib.tps_last_error(None)


def test_host_library_exports_every_declared_symbol():
    """include/topsicle_host.h is the ABI of libtps_host.so (reader, formatters, workload generator)."""
    lib = fastx.host_library()
    names = declared_functions("topsicle_host.h")
    for must in ("tps_fastx_open", "tps_fastx_next", "tps_fastx_next_spans", "tps_fastx_next_ends", "tps_fastx_release",
                 "tps_fastx_close", "tps_fastx_find_id", "tps_fastx_join_ids", "tps_fastx_gather_regions",
                 "tps_fastx_records_text", "tps_format_rawcount", "tps_synth_fill", "tps_synth_lengths",
                 "tps_host_threads", "tps_pgz_open", "tps_pgz_read", "tps_fastx_inflate_stats"):
        assert must in names, must
    for name in names:
        assert getattr(lib, name) is not None, name
    # and nothing is exported that the header does not declare
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", fastx.HOST_LIB_PATH], text=True)
    exported = sorted(ln.split()[-1] for ln in out.splitlines() if " T tps_" in ln)
    assert exported == names


def test_host_struct_layout_matches_header(tmp_path):
    import subprocess
    from topsicle_b200 import synth
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "topsicle_host.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(tps_fastx_rec), offsetof(tps_fastx_rec, seq_len),'
                   'sizeof(tps_synth_cfg), offsetof(tps_synth_cfg, f_telo), offsetof(tps_synth_cfg, motif));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    S = synth.SynthCfg
    assert got == [fastx.REC_DTYPE.itemsize, fastx.REC_DTYPE.fields["seq_len"][1], C.sizeof(S), S.f_telo.offset,
                   S.motif.offset]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "topsicle_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
