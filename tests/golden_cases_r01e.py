"""Checks against tests/golden/reference_vectors_r01e.npz (minimize_vectors, InterRDF_s,
Contacts; outputs of the unmodified reference, tests/golden/make_golden_r01e.py).  Shared by
the CPU oracle tests (-m "not gpu") and the CUDA tests (-m gpu)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BOXES = {
    "none": None,
    "ortho": np.array([30, 25, 40, 90, 90, 90], dtype=np.float32),
    "tric": np.array([40, 35, 50, 70, 80, 100], dtype=np.float32),
    "tric60": np.array([30, 30, 30, 60, 60, 90], dtype=np.float32),
}
_cache = {}


def golden():
    if "z" not in _cache:
        _cache["z"] = np.load(os.path.join(HERE, "golden", "reference_vectors_r01e.npz"))
    return _cache["z"]


def check_minimize_vectors(fn, bname, dtname):
    z = golden()
    k = f"minvec/{bname}/{dtname}/"
    got = fn(z[k + "v"], BOXES[bname])
    want = z[k + "out"]
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want)                     # same arithmetic, bit for bit


def rdfs_inputs(bname):
    z = golden()
    k = f"rdfs/{bname}/"
    sites = [[z[k + "sites0a"], z[k + "sites0b"]], [z[k + "sites1a"], z[k + "sites1b"]]]
    return z[k + "pos"], z[k + "dims"], sites


def check_rdfs(result, bname, norm):
    """result: dict with count / rdf / cdf lists and edges."""
    z = golden()
    k = f"rdfs/{bname}/{norm}/"
    assert np.array_equal(result["edges"], z[k + "edges"])
    for s in range(2):
        assert np.array_equal(result["count"][s], z[k + f"count{s}"])          # integer counts: exact
        assert np.allclose(result["rdf"][s], z[k + f"rdf{s}"], rtol=1e-12, atol=0)
        assert np.allclose(result["cdf"][s], z[k + f"cdf{s}"], rtol=1e-12, atol=0)
        assert result["count"][s].sum() > 0


CONTACT_CASES = [("hard_cut", None), ("soft_cut", None), ("radius_cut", None),
                 ("soft_cut", {"beta": 3.0, "lambda_constant": 1.5})]


def contacts_inputs(bname):
    z = golden()
    k = f"contacts/{bname}/"
    dims = z[k + "dims"] if bname != "none" else None
    return z[k + "pos"], dims, z[k + "idxA"], z[k + "idxB"]


def check_contacts(ts, bname, method, kw, rtol):
    z = golden()
    want = z[f"contacts/{bname}/{method}{'_kw' if kw else ''}"]
    assert ts.shape == want.shape
    assert np.array_equal(ts[:, 0], want[:, 0])
    if method in ("hard_cut", "radius_cut"):
        assert np.array_equal(ts[:, 1], want[:, 1])      # counts of exact float64 comparisons
    else:
        assert np.allclose(ts[:, 1], want[:, 1], rtol=rtol, atol=0)
    assert 0 < want[:, 1].mean() <= 1
