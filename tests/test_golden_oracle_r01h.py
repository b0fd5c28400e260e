"""-m "not gpu": the CPU oracle's DistanceMatrix against the golden matrices of the unmodified reference."""
import pytest

import golden_cases_r01h as gc


@pytest.mark.parametrize("key,sel,weighted,superposition,cutoff", gc.CASES)
def test_distance_matrix(oracle, key, sel, weighted, superposition, cutoff):
    xyz, idx, w = gc.inputs(sel, weighted)
    gc.check(oracle.distance_matrix(xyz, select=idx, weights=w, superposition=superposition, cutoff=cutoff),
             key, superposition)
