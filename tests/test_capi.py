"""The C-ABI library loads and exports every symbol include/pgb.h declares.  CPU only
(no compute calls)."""

import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pgb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from pygeon_b200 import _lib

    syms = header_symbols()
    assert len(syms) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in pgb.h but not exported by libpgb.so"
    assert set(_lib.EXPORTS) == set(syms), "ctypes table and header disagree"


def test_version_and_error_string():
    from pygeon_b200 import _lib

    assert _lib.lib.pgb_version() >= 100
    assert isinstance(_lib.lib.pgb_last_error(), bytes)


def test_local_sizes():
    from pygeon_b200 import _lib

    L = _lib.lib.pgb_local_size
    K = _lib.KIND
    assert [L(K["rt0_mass"], 3), L(K["n0_mass"], 3), L(K["bdm1_mass"], 3), L(K["l1_stiff"], 2)] == [4, 6, 12, 3]
    assert [L(K["n0_curlcurl"], 2), L(K["p0_mass"], 3), L(K["bdm1_divdiv"], 2)] == [3, 1, 6]


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pygeon_b200 as pg

    sd = pg.structured_grid(2, 2)
    sd.compute_geometry()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pg.RT0("k").assemble_mass_matrix(sd)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pygeon_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("oracle for", ""), fn


def test_ctypes_argument_counts_match_header():
    """Every prototype in pgb.h has as many parameters as its ctypes signature (a mismatch would
    corrupt the call silently: ctypes cannot check a C ABI)."""
    from pygeon_b200 import _lib

    text = open(os.path.join(ROOT, "include", "pgb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(pgb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) == len(_lib._SIGS)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib._SIGS[name][1]), f"{name}: header has {n} parameters, ctypes table {len(_lib._SIGS[name][1])}"
