"""-m gpu: all-pairs RMSD matrices (csrc/pairrmsd.cu, analysis/diffusionmap.py DistanceMatrix) through
the C ABI against the golden matrices of the real reference and against the oracle."""
import functools

import numpy as np
import pytest

import golden_cases_r01h as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key,sel,weighted,superposition,cutoff", gc.CASES)
@pytest.mark.parametrize("device", [False, True])
def test_distance_matrix_golden(key, sel, weighted, superposition, cutoff, device):
    import torch
    from mdanalysis_b200.analysis.diffusionmap import DistanceMatrix, rmsd_superposition
    from mdanalysis_b200.analysis.rms import rmsd
    xyz, idx, w = gc.inputs(sel, weighted)
    pos = torch.as_tensor(xyz).cuda() if device else xyz
    dm = DistanceMatrix(pos, select="all" if idx is None else idx, weights=w, cutoff=cutoff,
                        metric=rmsd_superposition if superposition else rmsd).run()
    gc.check(dm.results.dist_matrix, key, superposition)


@pytest.mark.parametrize("F,n", [(1, 5), (2, 1), (33, 17), (70, 100), (129, 48)])
@pytest.mark.parametrize("superposition", [False, True])
def test_pairwise_rmsd_vs_oracle(oracle, F, n, superposition):
    """Shapes around the 32-frame tile and the 16-atom chunk; partial(rmsd, superposition=True) as metric."""
    from mdanalysis_b200.analysis.diffusionmap import DistanceMatrix
    from mdanalysis_b200.analysis.rms import rmsd
    rng = np.random.default_rng(F * 100 + n)
    base = rng.normal(0, 8, (n, 3))
    traj = (base[None] + rng.normal(0, 1.5, (F, n, 3)) + rng.normal(0, 3, (F, 1, 3))).astype(np.float32)
    w = rng.uniform(1, 16, n) if n % 2 else None
    metric = functools.partial(rmsd, superposition=True) if superposition else rmsd
    got = DistanceMatrix(traj, metric=metric, weights=w).run().results.dist_matrix
    want = oracle.distance_matrix(traj, weights=w, superposition=superposition)
    np.testing.assert_allclose(got, want, rtol=1e-7 if superposition else 1e-10, atol=5e-6 if superposition else 0.0)
    assert np.array_equal(got, got.T)


def test_distance_matrix_front_end():
    from mdanalysis_b200.analysis.diffusionmap import DistanceMatrix
    traj = np.random.default_rng(1).normal(0, 5, (20, 30, 3)).astype(np.float32)
    full = DistanceMatrix(traj).run().results.dist_matrix
    part = DistanceMatrix(traj).run(start=2, stop=18, step=3)
    idx = np.arange(20)[2:18:3]
    assert part.n_frames == len(idx)
    assert np.array_equal(part.results.dist_matrix, full[np.ix_(idx, idx)])
    sub = DistanceMatrix(traj, select=np.arange(0, 30, 2)).run().results.dist_matrix
    assert np.array_equal(sub, DistanceMatrix(np.ascontiguousarray(traj[:, ::2])).run().results.dist_matrix)
    with pytest.raises(NotImplementedError):
        DistanceMatrix(traj, metric=lambda a, b, weights=None: 0.0)
    with pytest.raises(ValueError):
        DistanceMatrix(traj, weights=np.ones(7)).run()
    with pytest.warns(DeprecationWarning):
        assert DistanceMatrix(traj).run().dist_matrix.shape == (20, 20)
