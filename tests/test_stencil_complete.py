"""not gpu: the host-side criterion that lets the triclinic InterRDF path leave the reference's grid.

``mdagpu_grid_stencil_complete`` (csrc/rdf.cu ``tric_stencil_complete``) claims that FastNS on its own
grid returns EVERY pair within the cutoff.  Checked here against the oracle's restatement of FastNS
(lib/nsgrid.pyx, which reproduces the reference's misses) and the brute-force minimum-image
distances on random triclinic cells: wherever the library says "complete" the two must agree; the
survey's probe cell ([40, 35, 50, 70, 80, 100]) must be reported incomplete."""
import ctypes as C

import numpy as np


def _complete(lib, _cabi, oracle, box, cutoff, W=1):
    h9 = np.ascontiguousarray(oracle.triclinic_vectors(box), dtype=np.float32).reshape(9)
    out = C.c_int(-1)
    rc = lib.mdagpu_grid_stencil_complete(float(cutoff), _cabi.ptr(np.ascontiguousarray(box, dtype=np.float32)),
                                          _cabi.ptr(h9), W, C.byref(out))
    assert rc == 0
    return bool(out.value)


def test_stencil_completeness_criterion(oracle):
    from mdanalysis_b200 import _cabi
    lib = _cabi.load()
    rng = np.random.default_rng(5)
    said_complete = missing_in_incomplete = 0
    for trial in range(60):
        L, ang = rng.uniform(30, 60, 3), rng.uniform(55, 125, 3)
        if trial % 4 == 0:
            ang[2] = 90.0
        if trial % 8 == 1:
            ang[:] = [60, 60, 90]
        box = np.array([*L, *ang], dtype=np.float32)
        tv = oracle.triclinic_vectors(box).astype(np.float64)
        if not np.all(np.isfinite(tv)) or tv[2, 2] <= 5:
            continue
        cutoff = float(rng.uniform(3.0, 6.5))
        xyz = (rng.random((1800, 3)) @ tv).astype(np.float32)
        try:
            grid = oracle.FastNS(cutoff, xyz, box)
        except Exception:
            continue                                     # cutoff too large for this cell
        if not all(grid.info()[2]):
            continue
        found = len(grid.self_search().get_pairs())
        brute = int((oracle.self_distance_array(xyz, box) <= cutoff).sum())
        assert found <= brute
        if _complete(lib, _cabi, oracle, box, cutoff):
            said_complete += 1
            assert found == brute, (box, cutoff, found, brute)
        elif found < brute:
            missing_in_incomplete += 1
    assert said_complete >= 10 and missing_in_incomplete >= 10   # both sides of the criterion were exercised
    assert not _complete(lib, _cabi, oracle, np.array([40, 35, 50, 70, 80, 100], dtype=np.float32), 6.0)
    assert _complete(lib, _cabi, oracle, np.array([112.25, 112.25, 112.25, 60, 60, 90], dtype=np.float32), 15.0)
    assert _complete(lib, _cabi, oracle, np.array([100, 100, 100, 90, 90, 90], dtype=np.float32), 15.0)
