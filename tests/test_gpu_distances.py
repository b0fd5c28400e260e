"""-m gpu: lib.distances kernels through the C ABI vs the oracle (bit-exact where
stated, else the 1e-5 relative bar of BASELINE.json's north_star)."""
import numpy as np
import pytest

from conftest import BOXES, canon_pairs

pytestmark = pytest.mark.gpu


def cloud(rng, n, scale, spread=1.0, offset=0.0):
    return ((rng.random((n, 3)) - offset) * scale * spread).astype(np.float32)


@pytest.fixture(scope="module")
def gd():
    from mdanalysis_b200.lib import distances
    return distances


@pytest.mark.parametrize("bname", list(BOXES))
@pytest.mark.parametrize("n,m", [(1, 1), (7, 3), (300, 513), (1025, 77)])
@pytest.mark.parametrize("spread,off", [(1.0, 0.0), (5.0, 0.5)])
def test_distance_array_bit_exact(oracle, gd, bname, n, m, spread, off):
    rng = np.random.default_rng(n * 1000 + m)
    a, b = cloud(rng, n, 40, spread, off), cloud(rng, m, 40, spread, off)
    a0, b0 = a.copy(), b.copy()
    got = gd.distance_array(a, b, box=BOXES[bname])
    want = oracle.distance_array(a, b, BOXES[bname])
    assert got.dtype == np.float64 and got.shape == (n, m)
    assert np.array_equal(got, want)
    assert np.array_equal(a, a0) and np.array_equal(b, b0)      # inputs untouched


@pytest.mark.parametrize("bname", list(BOXES))
@pytest.mark.parametrize("n", [2, 3, 33, 600, 1031])
def test_self_distance_array_bit_exact(oracle, gd, bname, n):
    rng = np.random.default_rng(n)
    a = cloud(rng, n, 40, 3.0, 0.5)
    got = gd.self_distance_array(a, box=BOXES[bname])
    assert got.shape == (n * (n - 1) // 2,)
    assert np.array_equal(got, oracle.self_distance_array(a, BOXES[bname]))


@pytest.mark.parametrize("bname", list(BOXES))
@pytest.mark.parametrize("n,m", [(33, 1024), (65, 1025), (31, 2500), (40, 511)])
def test_distance_array_tile_edges(oracle, gd, bname, n, m):
    """Shapes around the 32 x 1024 (triclinic: 32 x 512) tile, coincident atoms (zero
    distances take the general sqrt path) -- csrc/distances.cu da_row_values."""
    rng = np.random.default_rng(n + m)
    a, b = cloud(rng, n, 40, 2.0, 0.5), cloud(rng, m, 40, 2.0, 0.5)
    b[:n:2] = a[::2]
    assert np.array_equal(gd.distance_array(a, b, box=BOXES[bname]), oracle.distance_array(a, b, BOXES[bname]))
    assert np.array_equal(gd.self_distance_array(b, box=BOXES[bname]), oracle.self_distance_array(b, BOXES[bname]))


@pytest.mark.parametrize("bname", ["ortho", "tric", "tric60"])
def test_distance_array_lattice_ties(oracle, gd, bname):
    """Points on a lattice commensurate with the cell: many pairs have two or more images at
    exactly the same distance (the reference keeps the first of its loop) and separations of
    exactly half a box edge (round() half away from zero)."""
    from mdanalysis_b200.lib.distances import _box_args
    box = BOXES[bname]
    h = np.asarray(_box_args(box)[1], dtype=np.float64).reshape(-1)
    h = np.diag(h) if h.size == 3 else h.reshape(3, 3)
    g = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3) / 4.0
    pts = (g @ h).astype(np.float32)
    got = gd.distance_array(pts, pts, box=box)
    assert np.array_equal(got, oracle.distance_array(pts, pts, box))
    assert np.array_equal(gd.self_distance_array(pts, box=box), oracle.self_distance_array(pts, box))


def test_distance_array_extreme_values(oracle, gd):
    """Overflowing and tiny separations: outside the range of the in-line square root."""
    a = np.array([[0, 0, 0], [3e38, 0, 0], [1e-30, 0, 0], [1e-45, 1e-45, 0]], dtype=np.float32)
    b = np.array([[0, 0, 0], [-3e38, 0, 0], [-1e-30, 0, 0], [2e19, 2e19, 2e19], [1, 2, 3]], dtype=np.float32)
    with np.errstate(all="ignore"):
        assert np.array_equal(gd.distance_array(a, b), oracle.distance_array(a, b, None), equal_nan=True)
        box = BOXES["ortho"]
        assert np.array_equal(gd.distance_array(a, b, box=box), oracle.distance_array(a, b, box), equal_nan=True)


def test_distance_array_result_argument(oracle, gd):
    rng = np.random.default_rng(1)
    a, b = cloud(rng, 50, 30), cloud(rng, 60, 30)
    res = np.zeros((50, 60))
    out = gd.distance_array(a, b, result=res)
    assert out is res and np.array_equal(res, oracle.distance_array(a, b))
    big = np.zeros((50, 120))
    view = big[:, ::2]                                            # non-contiguous result
    gd.distance_array(a, b, result=view)
    assert np.array_equal(big[:, ::2], oracle.distance_array(a, b)) and not big[:, 1::2].any()
    with pytest.raises(ValueError):
        gd.distance_array(a, b, result=np.zeros((1, 2)))
    with pytest.raises(TypeError):
        gd.distance_array(a, b, result=np.zeros((50, 60), dtype=np.float32))


def test_front_end_behaviour(gd):
    a = np.zeros((0, 3), dtype=np.float32)
    b = np.ones((4, 3), dtype=np.float32)
    assert gd.distance_array(a, b).shape == (0, 4)
    assert gd.self_distance_array(b[:1]).shape == (0,)
    assert gd.calc_bonds(a, a).shape == (0,)
    assert gd.distance_array(np.zeros(3), np.ones(3)).shape == (1, 1)          # single coordinates
    assert np.isscalar(gd.calc_bonds(np.zeros(3), np.ones(3)).item())
    assert gd.calc_bonds(np.zeros(3), np.ones(3)) == pytest.approx(np.sqrt(3.0))
    with pytest.raises(ValueError):
        gd.distance_array(b, b, backend="nonsense")
    assert gd.distance_array(b, b, backend="CUDA").shape == (4, 4)             # case-insensitive
    with pytest.raises(ValueError):
        gd.calc_bonds(b, b[:2])
    with pytest.raises(ValueError):
        gd.distance_array(np.zeros((3, 2)), b)
    with pytest.raises(TypeError):
        gd.distance_array([[0, 0, 0]], b)
    with pytest.raises(ValueError):
        gd.distance_array(b, b, box=np.array([1, 1, 1]))
    f64 = np.arange(12, dtype=np.float64).reshape(4, 3)                         # dtype conversion
    assert np.array_equal(gd.distance_array(f64, b), gd.distance_array(f64.astype(np.float32), b))


def test_known_answers(gd):
    """Constants pinned by the reference's own tests (testsuite/MDAnalysisTests/lib/test_distances.py)."""
    # :290-301 + :773-779 triclinic PBC distance
    box = np.array([10., 10., 10., 90., 90., 90.], dtype=np.float32)
    ref = np.array([[0., 0., 0.], [1., 1., 2.]], dtype=np.float32)
    conf = np.array([[1., 1., 1.], [4., 4., 4.], [9., 9., 9.]], dtype=np.float32)
    d = gd.distance_array(ref, conf)
    assert d[0, 0] == pytest.approx(np.sqrt(3)) and d[1, 0] == pytest.approx(1.0)
    d = gd.distance_array(ref, conf, box=box)
    assert d[0, 2] == pytest.approx(np.sqrt(3))                                 # wraps to (-1,-1,-1)
    # :894-947 bonds on the TestCythonFunctions fixture
    a = np.array([[0., 0., 0.], [0., 0., 0.], [0., 11., 0.], [1., 1., 1.]], dtype=np.float32)
    b = np.array([[0., 0., 0.], [1., 1., 1.], [0., 0., 0.], [29., -21., 99.]], dtype=np.float32)
    dists = gd.calc_bonds(a, b)
    assert dists == pytest.approx([0.0, 1.7320508075688772, 11.0, 104.26888318])
    dists = gd.calc_bonds(a, b, box=box)
    assert dists == pytest.approx([0.0, 1.7320508075688772, 1.0, 3.46410072])
    from mdanalysis_b200.lib import mdamath
    tbox = mdamath.triclinic_box([10., 0., 0.], [1., 10., 0.], [1., 0., 10.])     # :887-891
    dists = gd.calc_bonds(a, b, box=tbox)
    # reference tolerance: assert_almost_equal(..., prec=5)  (test_distances.py:878, :969-975)
    assert dists == pytest.approx([0.0, 1.7320508, 1.4142136, 2.82842712], abs=1.5e-5)


@pytest.mark.parametrize("bname", list(BOXES))
def test_bonds_angles_dihedrals(oracle, gd, bname):
    rng = np.random.default_rng(99)
    a, b, c, d = (cloud(rng, 2000, 40, 3.0, 0.5) for _ in range(4))
    box = BOXES[bname]
    assert np.array_equal(gd.calc_bonds(a, b, box=box), oracle.calc_bonds(a, b, box))
    # atan2 is not correctly rounded on either side: 1e-12 absolute
    assert np.allclose(gd.calc_angles(a, b, c, box=box), oracle.calc_angles(a, b, c, box), rtol=0, atol=1e-12)
    assert np.allclose(gd.calc_dihedrals(a, b, c, d, box=box), oracle.calc_dihedrals(a, b, c, d, box),
                       rtol=0, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("bname", ["ortho", "tric", "tric60"])
def test_apply_pbc_and_transforms(oracle, gd, bname):
    rng = np.random.default_rng(5)
    xyz = cloud(rng, 5000, 40, 30.0, 0.5)
    box = BOXES[bname]
    assert np.array_equal(gd.apply_PBC(xyz, box), oracle.apply_PBC(xyz, box))
    assert np.array_equal(gd.transform_RtoS(xyz, box), oracle.transform_RtoS(xyz, box))
    assert np.array_equal(gd.transform_StoR(xyz, box), oracle.transform_StoR(xyz, box))
    assert gd.apply_PBC(xyz[0], box).shape == (3,)


def test_tensor_twins(oracle, gd):
    import torch
    rng = np.random.default_rng(8)
    a, b = cloud(rng, 400, 40, 2.0, 0.3), cloud(rng, 300, 40, 2.0, 0.3)
    ta, tb = torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda()
    for box in BOXES.values():
        got = gd.distance_array(ta, tb, box=box)
        assert got.is_cuda and got.dtype == torch.float64
        assert np.array_equal(got.cpu().numpy(), oracle.distance_array(a, b, box))
        assert np.array_equal(gd.self_distance_array(ta, box=box).cpu().numpy(), oracle.self_distance_array(a, box))
        assert np.array_equal(gd.calc_bonds(ta[:300], tb, box=box).cpu().numpy(), oracle.calc_bonds(a[:300], b, box))
    box = BOXES["tric"]
    assert np.array_equal(gd.apply_PBC(ta, box).cpu().numpy(), oracle.apply_PBC(a, box))
    assert np.array_equal(ta.cpu().numpy(), a)       # input tensor untouched


def test_batched_frames_c_abi(oracle):
    """nframes > 1 entry point used by the benchmark (config C1 batched)."""
    import ctypes as C
    from mdanalysis_b200 import _cabi
    rng = np.random.default_rng(4)
    F, n, m = 3, 130, 70
    ref, conf = cloud(rng, F * n, 30).reshape(F, n, 3), cloud(rng, F * m, 30).reshape(F, m, 3)
    box = np.array([30, 30, 30], dtype=np.float32)
    out = np.zeros((F, n, m))
    _cabi.check(_cabi.load().mdagpu_distance_array(_cabi.ptr(ref), n, _cabi.ptr(conf), m, 1, _cabi.ptr(box),
                                                   _cabi.ptr(out), F, 0, None))
    so = np.zeros((F, n * (n - 1) // 2))
    _cabi.check(_cabi.load().mdagpu_self_distance_array(_cabi.ptr(ref), n, 1, _cabi.ptr(box), _cabi.ptr(so), F, 0, None))
    b6 = np.array([30, 30, 30, 90, 90, 90], dtype=np.float32)
    for f in range(F):
        assert np.array_equal(out[f], oracle.distance_array(ref[f], conf[f], b6))
        assert np.array_equal(so[f], oracle.self_distance_array(ref[f], b6))


@pytest.mark.gpu
@pytest.mark.parametrize("bname", ["none", "ortho"])
def test_self_distance_array_100k_atoms_materialised(oracle, bname):
    """Config C5 at full size: 100 000 atoms -> 4 999 950 000 distances (40 GB, one buffer on the
    device).  Checked through 200 000 sampled entries against the oracle's pair arithmetic (the
    packed index k = i(2n-i-1)/2 + (j-i-1), calc_distances.h:457-458, is part of what is tested)."""
    import torch
    from mdanalysis_b200 import _cabi
    n = 100_000
    free, _ = torch.cuda.mem_get_info()
    if free < 48 * 2 ** 30:
        pytest.skip("needs 40 GB of free device memory")
    rng = np.random.default_rng(55)
    L = 215.44
    x = (rng.random((n, 3)) * L * 1.05).astype(np.float32)
    box = None if bname == "none" else np.array([L, L, L, 90, 90, 90], dtype=np.float32)
    xd = torch.as_tensor(x[None]).cuda()
    out = torch.empty(n * (n - 1) // 2, dtype=torch.float64, device="cuda")
    mode = 0 if box is None else 1
    b3 = None if box is None else np.ascontiguousarray(box[:3])
    _cabi.check(_cabi.load().mdagpu_self_distance_array(_cabi.tensor_ptr(xd), n, mode, _cabi.ptr(b3),
                                                        _cabi.tensor_ptr(out), 1, 1, _cabi.current_stream(xd)))
    i = rng.integers(0, n - 1, 200_000)
    j = i + 1 + (rng.integers(0, n, 200_000) % (n - 1 - i))
    k = i * (2 * n - i - 1) // 2 + (j - i - 1)
    got = out[torch.as_tensor(k, device="cuda")].cpu().numpy()
    want = oracle.calc_bonds(np.ascontiguousarray(x[j]), np.ascontiguousarray(x[i]), box)
    assert np.array_equal(got, want)
    # first and last entries of the packed triangle
    ends = out[[0, -1]].cpu().numpy()
    assert np.array_equal(ends, oracle.calc_bonds(x[[1, n - 1]], x[[0, n - 2]], box))
    del out
    torch.cuda.empty_cache()
