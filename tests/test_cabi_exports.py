"""The C-ABI library loads without a GPU and exports every symbol include/mdagpu.h declares.

No compute entry point is called here (that is what the `-m gpu` tests do); the few
calls made are the ones that never touch a device.
"""
import ctypes as C
import os
import re

import pytest

from mdanalysis_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mdagpu.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"MDAGPU_API[^;(]*?\b(mdagpu_\w+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    assert len(names) >= 25
    for must in ("mdagpu_distance_array", "mdagpu_self_distance_array", "mdagpu_calc_bonds", "mdagpu_calc_angles",
                 "mdagpu_calc_dihedrals", "mdagpu_coord_transform", "mdagpu_wrap", "mdagpu_grid_create",
                 "mdagpu_grid_search", "mdagpu_grid_self_search", "mdagpu_pairs_export", "mdagpu_capped_bruteforce",
                 "mdagpu_qcp", "mdagpu_qcp_trajectory", "mdagpu_rdf_accumulate", "mdagpu_contacts_self"):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_cabi.LIB_PATH):
        pytest.fail(f"{_cabi.LIB_PATH} missing: run __graft_entry__.build()")
    lib = C.CDLL(_cabi.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"libmdagpu.so lacks {missing}"


def test_ctypes_table_matches_header():
    """_cabi.SIGNATURES binds exactly the declared entry points, with the declared arity."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decl = {}
    for m in re.finditer(r"MDAGPU_API[^;(]*?\b(mdagpu_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    assert set(decl) == set(_cabi.SIGNATURES)
    for name, (_, argtypes) in _cabi.SIGNATURES.items():
        assert len(argtypes) == decl[name], name


def test_load_and_deviceless_calls():
    lib = _cabi.load()
    assert lib.mdagpu_abi_version() == _cabi.ABI_VERSION
    assert isinstance(_cabi.launch_count(), int)
    assert lib.mdagpu_last_error() is not None


def test_invalid_arguments_are_reported_without_a_device():
    """Argument validation happens before any CUDA call: ValueError like the reference front end."""
    lib = _cabi.load()
    rc = lib.mdagpu_rdf_accumulate(None, 10, None, 10, 1, None, None, 0, 0, 0, 5.0, 25.0, 1, None, 10, 0.0, 5.0,
                                   0, 0, None, _cabi.HOST, None)
    assert rc == -2
    with pytest.raises(ValueError):
        _cabi.check(rc)
    # geometry-only entry point (host arithmetic): FastNS._prepare_box on the reference's test box
    import numpy as np
    box6 = np.array([10, 10, 10, 90, 90, 90], dtype=np.float32)
    h9 = np.diag([10.0, 10.0, 10.0]).astype(np.float32).reshape(9)
    nc = (C.c_int * 3)()
    cs = (C.c_double * 9)()
    per = (C.c_int * 3)()
    tri = C.c_int()
    mc = C.c_float()
    rc = lib.mdagpu_grid_geometry(2.0, _cabi.ptr(box6), _cabi.ptr(h9), 1, nc, cs, per, C.byref(tri), C.byref(mc))
    assert rc == 0
    assert list(nc) == [4, 4, 4] and list(per) == [1, 1, 1] and tri.value == 0     # floor(10 / 2.001) = 4
    assert mc.value == pytest.approx(5.0)


def test_no_cpu_fallback_in_product():
    """Nothing under mdanalysis_b200/ imports the oracle."""
    pkg = os.path.join(ROOT, "mdanalysis_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"
                assert "liboracle" not in src and "libmda_ref" not in src, f"{fn} links the oracle"
