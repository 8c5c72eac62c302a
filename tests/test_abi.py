"""CPU: the C-ABI library loads and exports every symbol include/vgl_b200.h declares; the ctypes mirror agrees with
the header (names and struct sizes); compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vgl_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(vglb_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_exported(vgl):
    L = vgl.lib()
    declared = _declared_functions()
    assert len(declared) >= 40
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, f"libvgl_b200.so does not export {missing}"


def test_ctypes_mirror_covers_header(vgl):
    declared = set(_declared_functions())
    mirrored = set(vgl._SIGNATURES)
    assert declared == mirrored, (sorted(declared - mirrored), sorted(mirrored - declared))


def test_struct_sizes_match_header(vgl):
    prog = r"""
    #include <stdio.h>
    #include "vgl_b200.h"
    int main(void) {
        printf("%zu %zu %zu %zu\n", sizeof(vglb_graph_info), sizeof(vglb_stats), sizeof(vglb_bfs_opts),
               sizeof(vglb_frontier_info));
        return 0;
    }
    """
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(vgl.GraphInfo), C.sizeof(vgl.Stats), C.sizeof(vgl.BfsOpts), C.sizeof(vgl.FrontierInfo)]


def test_no_cpu_fallback(vgl):
    if vgl.lib().vglb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(vgl.VglbError) as e:
        vgl.Context(0)
    assert "no CPU fallback" in str(e.value) and "[vglb error 2]" in str(e.value)


def test_host_generator_matches_oracle(vgl, oracle):
    for kind, scale, ef in [(0, 9, 4), (1, 8, 16), (2, 10, 2)]:
        s0, d0 = oracle.generate_edges(kind, scale, ef)
        s1, d1 = vgl.generate_edges_host(kind, scale, ef)
        assert np.array_equal(s0, s1) and np.array_equal(d0, d1)
        assert s0.min() >= 0 and s0.max() < (1 << scale)


def test_argument_errors_are_reported(vgl):
    L = vgl.lib()
    assert L.vglb_generate_edges_host(7, 4, 16, 1, 57, 19, 19, None, None) == 1  # VGLB_EINVAL
    assert b"NULL" in L.vglb_last_error()
    out = C.c_void_p()
    assert L.vglb_init(-1, C.byref(out)) == 2  # VGLB_ENODEVICE, whether or not a GPU exists
