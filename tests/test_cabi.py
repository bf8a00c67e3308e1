"""The drop-in boundary: libb2az.so must load and export every symbol include/b2az.h declares
(no compute calls here: this runs without a GPU), and must refuse to work without CUDA."""
import ctypes as C
import os
import re

import pytest

import b2az
import parity_harness as ph
from conftest import has_cuda


def _declared_symbols():
    text = open(os.path.join(ph.ROOT, "include", "b2az.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2az_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    syms = _declared_symbols()
    assert "b2az_create" in syms and "b2az_step" in syms and len(syms) >= 12
    L = C.CDLL(b2az.DEFAULT_LIB)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/b2az.h but not exported by libb2az.so"


def test_struct_layouts_match_header():
    # sizes computed by the C compiler for the header's structs must equal the ctypes mirrors
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "b2az.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(b2az_params), sizeof(b2az_stats), sizeof(b2az_tafl_selfplay_params), sizeof(b2az_perm_stats), sizeof(b2az_forest_params), sizeof(b2az_tafl_selfplay_slot));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ph.ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")], text=True).split()
    assert int(out[0]) == C.sizeof(b2az.Params) and int(out[1]) == C.sizeof(b2az.Stats)
    assert int(out[2]) == C.sizeof(b2az.TaflSelfplayParams) and int(out[3]) == C.sizeof(b2az.PermStats)
    assert int(out[4]) == C.sizeof(b2az.ForestParams) and int(out[5]) == b2az.SLOT_DTYPE.itemsize


@pytest.mark.skipif(has_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    lib = b2az.load()
    with pytest.raises(b2az.B2azError) as ei:
        b2az.Engine(b2az.default_params(lib), lib=lib)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
