"""Shared checks against tests/golden/reference_vectors_r01h.npz (DistanceMatrix from the real reference)."""
import os

import numpy as np

_G = None
CASES = [("dm/ca", "ca", False, False, -4.0), ("dm/all", "all", False, False, -4.0),
         ("dm/ca_weighted", "ca", True, False, -4.0), ("dm/ca_cutoff", "ca", False, False, 1.5),
         ("dmfit/ca", "ca", False, True, -4.0), ("dmfit/ca_weighted", "ca", True, True, -4.0)]


def golden():
    global _G
    if _G is None:
        _G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_r01h.npz")))
    return _G


def inputs(sel, weighted):
    g = golden()
    return g["xyz"], (None if sel == "all" else g["ca_indices"]), (g["weights"] if weighted else None)


def check(matrix, key, superposition):
    want = golden()[key]
    assert matrix.shape == want.shape and matrix.dtype == np.float64
    assert np.array_equal(matrix, matrix.T)
    # 1e-5 relative is north_star's bar for RMSD; the diagonal of the superposed flavour is
    # round-off noise of the quartic (~1e-7..1e-6 A in the reference as well): absolute bar there
    np.testing.assert_allclose(matrix, want, rtol=1e-9 if not superposition else 1e-7,
                               atol=0.0 if not superposition else 5e-6)
    if not superposition:
        assert np.all(np.diag(matrix) == 0.0)
