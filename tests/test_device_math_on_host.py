"""The arithmetic the kernels execute (csrc/geom.cuh), compiled for the CPU and
compared BIT FOR BIT with the oracle.  Runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import BOXES

F32P = C.POINTER(C.c_float)
F64P = C.POINTER(C.c_double)


def p(a, t):
    return a.ctypes.data_as(t)


def cloud(rng, n, scale, spread=1.0, offset=0.0):
    return ((rng.random((n, 3)) - offset) * scale * spread).astype(np.float32)


def mode_box(orc, box):
    if box is None:
        return 0, np.zeros(9, dtype=np.float32)
    return orc.check_box(box)


@pytest.mark.parametrize("bname", ["ortho", "tric", "tric60"])
@pytest.mark.parametrize("spread,off", [(1.0, 0.0), (3.0, 0.5), (40.0, 0.5)])
def test_wrap_bit_exact(oracle, hostcheck, bname, spread, off):
    rng = np.random.default_rng(7)
    xyz = cloud(rng, 4000, 40, spread, off)
    edge = np.array([[0, 0, 0], [30, 25, 40], [-1e-7, 25, -0.0], [29.999998, 1e-9, 40.000004],
                     [-30, -25, -40], [60, 50, 80]], dtype=np.float32)
    xyz = np.concatenate([xyz, edge])
    want = oracle.apply_PBC(xyz, BOXES[bname])
    mode, b = oracle.check_box(BOXES[bname])
    got = xyz.copy()
    hostcheck.hc_wrap(p(got, F32P), C.c_uint64(len(got)), mode, p(b, F32P))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("bname", ["none", "ortho", "tric", "tric60"])
@pytest.mark.parametrize("spread,off", [(1.0, 0.0), (3.0, 0.5), (25.0, 0.5)])
def test_pair_d2_matches_distance_array(oracle, hostcheck, bname, spread, off):
    """pair_d2_exact<lib.distances flavour> == distance_array**2 pre-sqrt, via bonds."""
    rng = np.random.default_rng(11)
    a = cloud(rng, 5000, 40, spread, off)
    b = cloud(rng, 5000, 40, spread, off)
    box = BOXES[bname]
    want = oracle.calc_bonds(a, b, box)          # |a - b| with the reference's arithmetic
    mode, bb = mode_box(oracle, box)
    aw, bw = a, b
    if mode == 2:                                # triclinic kernels see wrapped copies
        aw, bw = oracle.apply_PBC(a, box), oracle.apply_PBC(b, box)
    mic = {0: 0, 1: 1, 2: 3}[mode]
    d2 = np.zeros(len(a))
    hostcheck.hc_pair_d2(mic, p(aw, F32P), p(bw, F32P), C.c_uint64(len(a)), mode, p(bb, F32P), p(d2, F64P))
    assert np.array_equal(np.sqrt(d2), want)


def test_pair_d2_triclinic_lattice_ties(oracle, hostcheck):
    """Integer lattice data produces exact image ties: the FP32 image scan must
    fall back to the exhaustive loop and still match."""
    rng = np.random.default_rng(5)
    box = np.array([10, 10, 10, 60, 60, 90], dtype=np.float32)
    a = rng.integers(0, 10, (3000, 3)).astype(np.float32)
    b = rng.integers(0, 10, (3000, 3)).astype(np.float32) * 0.5
    want = oracle.calc_bonds(a, b, box)
    mode, bb = oracle.check_box(box)
    aw, bw = oracle.apply_PBC(a, box), oracle.apply_PBC(b, box)
    d2 = np.zeros(len(a))
    hostcheck.hc_pair_d2(3, p(aw, F32P), p(bw, F32P), C.c_uint64(len(a)), mode, p(bb, F32P), p(d2, F64P))
    assert np.array_equal(np.sqrt(d2), want)


def test_pair_d2_ortho_lazy_matches_fastns(oracle, hostcheck):
    rng = np.random.default_rng(3)
    box = BOXES["ortho"]
    xyz = cloud(rng, 800, 40, 2.0, 0.3)
    g = oracle.FastNS(7.0, xyz, box)
    res = g.self_search()
    pairs, d2 = res.get_pairs(), res.get_pair_distances2()
    w = oracle.apply_PBC(xyz, box)
    mode, bb = oracle.check_box(box)
    a = np.ascontiguousarray(w[pairs[:, 0]])
    b = np.ascontiguousarray(w[pairs[:, 1]])
    got = np.zeros(len(a))
    hostcheck.hc_pair_d2(2, p(a, F32P), p(b, F32P), C.c_uint64(len(a)), mode, p(bb, F32P), p(got, F64P))
    assert len(pairs) > 1000
    assert np.array_equal(got, d2)


@pytest.mark.parametrize("bname", ["none", "ortho", "tric"])
@pytest.mark.parametrize("kind", [2, 3, 4])
def test_tuples_bit_exact(oracle, hostcheck, bname, kind):
    rng = np.random.default_rng(13)
    arrs = [cloud(rng, 3000, 40, 3.0, 0.5) for _ in range(4)]
    box = BOXES[bname]
    fn = {2: oracle.calc_bonds, 3: oracle.calc_angles, 4: oracle.calc_dihedrals}[kind]
    want = fn(*arrs[:kind], box)
    mode, bb = mode_box(oracle, box)
    got = np.zeros(len(arrs[0]))
    hostcheck.hc_tuples(kind, *[p(a, F32P) for a in arrs], C.c_uint64(len(got)), mode, p(bb, F32P), p(got, F64P))
    # sqrt is exact; atan2 on the device may differ from libm by an ulp, on the host it is libm
    assert np.array_equal(got, want, equal_nan=True)


def test_degenerate_angles_and_dihedrals(oracle, hostcheck):
    z = np.zeros((4, 3), dtype=np.float32)
    line = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]], dtype=np.float32)
    got = np.zeros(4)
    bb = np.zeros(9, dtype=np.float32)
    hostcheck.hc_tuples(3, p(z, F32P), p(z, F32P), p(z, F32P), p(z, F32P), C.c_uint64(4), 0, p(bb, F32P), p(got, F64P))
    assert np.array_equal(got, oracle.calc_angles(z, z, z))          # coincident points -> 0
    a, b, c, d = (np.ascontiguousarray(np.tile(line[k], (4, 1))) for k in range(4))
    hostcheck.hc_tuples(4, p(a, F32P), p(b, F32P), p(c, F32P), p(d, F32P), C.c_uint64(4), 0, p(bb, F32P), p(got, F64P))
    assert np.isnan(got).all() and np.isnan(oracle.calc_dihedrals(a, b, c, d)).all()


@pytest.mark.parametrize("bname,cutoff", [("ortho", 4.0), ("tric", 6.0), ("tric60", 2.5)])
def test_cell_assignment(oracle, hostcheck, bname, cutoff):
    rng = np.random.default_rng(17)
    box = BOXES[bname]
    xyz = oracle.apply_PBC(cloud(rng, 5000, 60, 1.0, 0.2), box)
    nc, cs, per, tri, mc = oracle.grid_geometry(cutoff, box)
    out = np.zeros(len(xyz), dtype=np.int64)
    nc32 = nc.astype(np.int32)
    hostcheck.hc_cells(p(xyz, F32P), C.c_uint64(len(xyz)), nc32.ctypes.data_as(C.POINTER(C.c_int)),
                       p(cs, F64P), out.ctypes.data_as(C.POINTER(C.c_int64)))
    # reference formula in NumPy float64
    x, y, z = (xyz[:, k].astype(np.float64) for k in range(3))
    cz = (z / cs[8]).astype(np.int64) % nc[2]
    cy = ((y - z * cs[7]) / cs[4]).astype(np.int64) % nc[1]
    cx = ((x - y * cs[3] - z * cs[6]) / cs[0]).astype(np.int64) % nc[0]
    assert np.array_equal(out, cx + cy * nc[0] + cz * nc[0] * nc[1])
    assert out.min() >= 0 and out.max() < nc.prod()


def test_qcp_solve_matches_oracle(oracle, hostcheck):
    rng = np.random.default_rng(19)
    for n in (3, 40, 2000):
        x = rng.normal(0, 10, (n, 3))
        y = x @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 0.7, (n, 3))
        x -= x.mean(0)
        y -= y.mean(0)
        rot = np.zeros(9)
        want = oracle.CalcRMSDRotationalMatrix(x, y, n, rot, None)
        A = (y.T @ x).reshape(9).copy()           # A[a,b] = sum conf_a ref_b
        E0 = 0.5 * ((x * x).sum() + (y * y).sum())
        got_rot = np.zeros(9)
        got = hostcheck.hc_qcp_solve(p(got_rot, F64P), p(A, F64P), C.c_double(E0), C.c_double(n))
        assert got == pytest.approx(want, rel=1e-10)
        assert np.allclose(got_rot, rot, atol=1e-10)
        assert np.allclose(y @ got_rot.reshape(3, 3), x, atol=5 * want + 1e-9) or n == 3


def test_histogram_bin_semantics(hostcheck):
    rng = np.random.default_rng(23)
    for nbins, rng_ in ((75, (0.0, 15.0)), (150, (0.0, 15.0)), (33, (1.5, 9.25))):
        _, edges = np.histogram([-1], bins=nbins, range=rng_)
        d = np.concatenate([rng.uniform(rng_[0], rng_[1], 20000), edges,
                            np.nextafter(edges, np.inf)[:-1], np.nextafter(edges, -np.inf)[1:]])
        want, _ = np.histogram(d, bins=nbins, range=rng_)
        got = np.zeros(nbins, dtype=np.int64)
        for v in d:
            got[hostcheck.hc_histogram_bin(C.c_double(v), p(edges, F64P), nbins)] += 1
        assert np.array_equal(got, want)


def _brute27(d32, h):
    """calc_distances.h:72-114 in NumPy: float32 difference widened, 27 images in loop order
    (x, y, z), strict <, shifts h*i evaluated in float32."""
    d = d32.astype(np.float64)
    best = np.full(len(d), float(np.finfo(np.float32).max))
    for a in (-1, 0, 1):
        xa = d[:, 0] + np.float64(np.float32(h[0] * np.float32(a)))
        for b in (-1, 0, 1):
            xab = xa + np.float64(np.float32(h[3] * np.float32(b)))
            yb = d[:, 1] + np.float64(np.float32(h[4] * np.float32(b)))
            for c in (-1, 0, 1):
                x = xab + np.float64(np.float32(h[6] * np.float32(c)))
                y = yb + np.float64(np.float32(h[7] * np.float32(c)))
                z = d[:, 2] + np.float64(np.float32(h[8] * np.float32(c)))
                r2 = x * x + y * y + z * z
                best = np.where(r2 < best, r2, best)
    return best


@pytest.mark.parametrize("dims", [[10, 10, 10, 50, 60, 70], [20, 8, 30, 100, 120, 95], [5, 40, 12, 89, 91, 120],
                                  [30, 30, 30, 60, 60, 90], [50, 50, 50, 70.53571, 109.48542, 70.518196],
                                  [12, 12, 12, 35, 40, 45], [9, 9, 9, 90, 90, 60]])
@pytest.mark.parametrize("spread", [1.0, 2.5])
def test_pair_d2_triclinic_eight_image_shortcut(oracle, hostcheck, dims, spread):
    """The 8-image shortcut of mic_triclinic_fast8 (accepted only below the smallest
    perpendicular width of the cell) against the exhaustive 27-image loop, on skewed cells,
    for differences inside the cell (the distance_array case) and well outside it."""
    rng = np.random.default_rng(int(sum(dims)))
    box = np.array(dims, dtype=np.float32)
    mode, bb = oracle.check_box(box)
    h = bb.reshape(9)
    tv = h.reshape(3, 3).astype(np.float64)
    n = 40000
    a = ((rng.random((n, 3)) * spread - (spread - 1) / 2) @ tv).astype(np.float32)
    b = ((rng.random((n, 3)) * spread - (spread - 1) / 2) @ tv).astype(np.float32)
    if spread == 1.0:
        a, b = oracle.apply_PBC(a, box), oracle.apply_PBC(b, box)
    d2 = np.zeros(n)
    hostcheck.hc_pair_d2(3, p(a, F32P), p(b, F32P), C.c_uint64(n), mode, p(bb, F32P), p(d2, F64P))
    assert np.array_equal(d2, _brute27(a - b, h))
