"""-m gpu: FastNS / capped distances through the C ABI vs the oracle: pair SETS and
distances bit-exact (order is not part of the contract, as in the reference's tests)."""
import numpy as np
import pytest

from conftest import BOXES, canon_pairs

pytestmark = pytest.mark.gpu


def same_pairs(got, want, self_pairs=False):
    p1, d1 = canon_pairs(*got, self_pairs)
    p2, d2 = canon_pairs(*want, self_pairs)
    assert p1.shape == p2.shape, (p1.shape, p2.shape)
    assert np.array_equal(p1, p2)
    assert np.array_equal(d1, d2)
    return len(p1)


@pytest.mark.parametrize("bname,cutoff,n", [("ortho", 4.0, 3000), ("ortho", 7.3, 1500), ("tric", 6.0, 3000),
                                            ("tric60", 5.0, 2500), ("tric60", 2.0, 2500), ("tric", 0.5, 500)])
def test_fastns_search_and_self_search(oracle, bname, cutoff, n):
    from mdanalysis_b200.lib.nsgrid import FastNS
    rng = np.random.default_rng(int(cutoff * 10) + n)
    box = BOXES[bname]
    tv = oracle.triclinic_vectors(box).astype(np.float64)
    xyz = ((rng.random((n, 3)) * 3 - 1) @ tv).astype(np.float32)     # up to a cell outside
    q = ((rng.random((700, 3)) * 3 - 1) @ tv).astype(np.float32)
    g, o = FastNS(cutoff, xyz, box), oracle.FastNS(cutoff, xyz, box)
    nc, cs, per, tri, mc = g.grid_info()
    onc, ocs, oper, otri = o.info()
    assert np.array_equal(nc, onc) and np.array_equal(cs, ocs) and np.array_equal(per, oper) and tri == otri
    r, ro = g.self_search(), o.self_search()
    k = same_pairs((r.get_pairs(), r.get_pair_distances()), (ro.get_pairs(), ro.get_pair_distances()), True)
    assert k > 0
    r, ro = g.search(q), o.search(q)
    same_pairs((r.get_pairs(), r.get_pair_distances()), (ro.get_pairs(), ro.get_pair_distances()))
    assert r.get_pairs().dtype == np.intp and r.get_pair_distances().dtype == np.float64


@pytest.mark.parametrize("bname", ["ortho", "tric60"])
def test_fastns_clustered_density(oracle, bname):
    """Strongly non-uniform density: the row kernel sizes its tile from the MEAN cell occupancy
    (csrc/nsgrid.cu choose_traversal), so the cells of a dense blob do not fit and take the
    row-by-row staging with its run-time buffer size (csrc/rowpairs.cuh); sparse cells next to
    them take the plane path.  Same pair sets and distances as the oracle either way."""
    from mdanalysis_b200.lib.nsgrid import FastNS
    rng = np.random.default_rng(77)
    box = BOXES[bname]
    tv = oracle.triclinic_vectors(box).astype(np.float64)
    centre = np.array([0.5, 0.5, 0.5]) @ tv
    xyz = np.concatenate([(rng.random((6000, 3)) @ tv), centre + rng.normal(0.0, 3.0, (2500, 3)),
                          (np.array([0.02, 0.97, 0.5]) @ tv) + rng.normal(0.0, 1.5, (600, 3))]).astype(np.float32)
    q = np.concatenate([(rng.random((500, 3)) @ tv), centre + rng.normal(0.0, 4.0, (300, 3))]).astype(np.float32)
    cutoff = 4.0
    g, o = FastNS(cutoff, xyz, box), oracle.FastNS(cutoff, xyz, box)
    r, ro = g.self_search(), o.self_search()
    k = same_pairs((r.get_pairs(), r.get_pair_distances()), (ro.get_pairs(), ro.get_pair_distances()), True)
    assert k > 500_000                                               # the blobs dominate
    r, ro = g.search(q), o.search(q)
    same_pairs((r.get_pairs(), r.get_pair_distances()), (ro.get_pairs(), ro.get_pair_distances()))


def test_fastns_errors_and_edges(oracle):
    from mdanalysis_b200.lib.nsgrid import FastNS
    xyz = np.random.default_rng(0).random((100, 3)).astype(np.float32) * 10
    box = np.array([10, 10, 10, 90, 90, 90], dtype=np.float32)
    with pytest.raises(ValueError):
        FastNS(6.0, xyz, box)                       # cutoff > half box
    with pytest.raises(ValueError):
        FastNS(-1.0, xyz, box)
    with pytest.raises(ValueError):
        FastNS(1.0, xyz, np.array([0, 10, 10, 90, 90, 90], dtype=np.float32))
    with pytest.raises(ValueError):
        FastNS(1.0, xyz[:, :2], box)
    with pytest.raises(ValueError):
        FastNS(1.0, xyz, box[:3])
    g = FastNS(1.0, xyz[:1], box)
    assert g.self_search().get_pairs().shape == (0, 2)
    assert g.search(np.zeros((0, 3), dtype=np.float32)).get_pairs().shape == (0, 2)
    # reference regression #2132: point on the upper box edge
    pts = np.array([[0.1, 0.1, 0.1], [9.999999, 9.999999, 9.999999]], dtype=np.float32)
    r, ro = FastNS(1.0, pts, box).self_search(), oracle.FastNS(1.0, pts, box).self_search()
    assert len(r.get_pairs()) == len(ro.get_pairs()) == 1
    assert np.array_equal(r.get_pair_distances(), ro.get_pair_distances())


def test_nsgrid_known_answers():
    """Known answers of testsuite/MDAnalysisTests/lib/test_nsgrid.py:164-239."""
    from mdanalysis_b200.lib.nsgrid import FastNS
    np.random.seed(90003)
    points = (np.random.uniform(low=0, high=1.0, size=(100, 3)) * 10.).astype(np.float32)
    query = np.array([1., 1., 1.], dtype=np.float32).reshape((1, 3))

    def pseudo(coords_all):
        pb = np.zeros(6, dtype=np.float32)
        lmax, lmin = coords_all.max(axis=0), coords_all.min(axis=0)
        pb[:3] = 1.1 * (lmax - lmin)
        pb[3:] = 90.
        return pb, lmin

    # search, cutoff 2.0 (:164-191)
    pb, lmin = pseudo(np.concatenate([points, query]))
    res = FastNS(2.0, points - lmin, box=pb, pbc=False).search(query - lmin)
    assert sorted(res.get_pairs()[:, 1].tolist()) == [3, 13, 24]
    for box, want in ((np.array([10., 10., 10., 90., 90., 90.]), [3, 13, 24, 39, 67]),
                      (np.array([10., 10., 10., 60., 75., 90.]), [3, 13, 24, 39, 60, 79])):
        res = FastNS(2.0, points, box).search(query)
        assert sorted(res.get_pairs()[:, 1].tolist()) == want
    # self search, cutoff 1.0 (:194-222)
    pb, lmin = pseudo(points)
    assert len(FastNS(1.0, points - lmin, box=pb, pbc=False).self_search().get_pairs()) == 21
    assert len(FastNS(1.0, points, np.array([10., 10., 10., 90., 90., 90.])).self_search().get_pairs()) == 26
    assert len(FastNS(1.0, points, np.array([10., 10., 10., 60., 75., 90.])).self_search().get_pairs()) == 33
    # issue #2132 (:224-239)
    ref = np.array([[55.783722, 44.190044, -54.16671]], dtype=np.float32)
    box = np.array([53.785854, 43.951054, 57.17597, 90., 90., 90.], dtype=np.float32)
    results = FastNS(3.0, np.ones((1, 3), dtype=np.float32), box).search(ref)
    assert np.array_equal(results.get_pairs(), np.zeros((1, 2), dtype=np.int64))
    assert results.get_pair_distances() == pytest.approx([2.3689647], rel=1e-6)


@pytest.mark.parametrize("bname", list(BOXES))
@pytest.mark.parametrize("method", [None, "bruteforce", "nsgrid", "pkdtree"])
@pytest.mark.parametrize("mn", [None, 1.5])
def test_capped_distance(oracle, bname, method, mn):
    from mdanalysis_b200.lib import distances as gd
    rng = np.random.default_rng(31)
    a = (rng.random((400, 3)) * 30).astype(np.float32)
    b = (rng.random((300, 3)) * 30).astype(np.float32)
    box = BOXES[bname]
    omethod = "bruteforce" if method == "pkdtree" else method
    n = same_pairs(gd.capped_distance(a, b, 5.0, min_cutoff=mn, box=box, method=method),
                   oracle.capped_distance(a, b, 5.0, min_cutoff=mn, box=box, method=omethod))
    assert n > 100
    got = gd.self_capped_distance(a, 5.0, min_cutoff=mn, box=box, method=method)
    same_pairs(got, oracle.self_capped_distance(a, 5.0, min_cutoff=mn, box=box, method=omethod), True)
    if method in ("bruteforce", "pkdtree"):
        assert (got[0][:, 0] < got[0][:, 1]).all()       # brute force orients i < j
    p = gd.capped_distance(a, b, 5.0, box=box, method=method, return_distances=False)
    assert isinstance(p, np.ndarray) and p.shape[1] == 2


@pytest.mark.parametrize("bname", ["none", "ortho"])
def test_bruteforce_on_a_grid_for_large_inputs(oracle, bname, monkeypatch):
    """method='bruteforce' / 'pkdtree' with >= 8192 atoms: same float64 arithmetic and pair orientation,
    but binned on a grid of cells >= max_cutoff (csrc/nsgrid.cu mdagpu_capped_bruteforce) instead of
    all pairs.  Input NOT wrapped into the box (atoms up to three boxes away), against the oracle's
    all-pairs loop and against the library's own all-pairs path."""
    from mdanalysis_b200.lib import distances as gd
    rng = np.random.default_rng(77)
    box = None if bname == "none" else np.array([60, 55, 70, 90, 90, 90], dtype=np.float32)
    a = (rng.random((9000, 3)) * 60).astype(np.float32)
    b = (rng.random((8500, 3)) * 60).astype(np.float32)
    if box is not None:
        a += (rng.integers(-3, 4, (9000, 3)) * box[:3]).astype(np.float32)
        b[::7] -= box[:3] * 2
    for mn in (None, 1.0):
        got = gd.capped_distance(a, b, 4.0, min_cutoff=mn, box=box, method="pkdtree")
        n = same_pairs(got, oracle.capped_distance(a, b, 4.0, min_cutoff=mn, box=box, method="bruteforce"))
        assert n > 10000
        sgot = gd.self_capped_distance(a, 4.0, min_cutoff=mn, box=box, method="bruteforce")
        same_pairs(sgot, oracle.self_capped_distance(a, 4.0, min_cutoff=mn, box=box, method="bruteforce"), True)
        assert (sgot[0][:, 0] < sgot[0][:, 1]).all()
    gridded = gd.capped_distance(a, b, 4.0, box=box, method="bruteforce")
    monkeypatch.setenv("MDAGPU_NO_BRUTE_GRID", "1")
    same_pairs(gridded, gd.capped_distance(a, b, 4.0, box=box, method="bruteforce"))


def test_capped_distance_edges(oracle):
    from mdanalysis_b200.lib import distances as gd
    e = np.zeros((0, 3), dtype=np.float32)
    x = np.ones((5, 3), dtype=np.float32)
    p, d = gd.capped_distance(e, x, 3.0)
    assert p.shape == (0, 2) and d.shape == (0,) and p.dtype == np.intp
    p, d = gd.self_capped_distance(x[:1], 3.0)
    assert p.shape == (0, 2)
    p, d = gd.capped_distance(x[0], x, 3.0)                 # single coordinate
    assert len(p) == 5 and (d == 0).all()
    # all atoms coincide: every pair at distance 0, including min_cutoff exclusion
    p, d = gd.self_capped_distance(x, 3.0, min_cutoff=0.0)
    assert len(p) == 0
    p, d = gd.self_capped_distance(x, 3.0)
    assert len(p) == 10
    # dense single cell -> overflow of the first allocation guess must be handled
    rng = np.random.default_rng(2)
    y = (rng.random((3000, 3)) * 2).astype(np.float32)
    got = gd.self_capped_distance(y, 1.9, method="bruteforce")
    want = oracle.self_capped_distance(y, 1.9, method="bruteforce")
    same_pairs(got, want, True)


def test_contact_matrix(oracle):
    from mdanalysis_b200.analysis.distances import contact_matrix
    rng = np.random.default_rng(6)
    x = (rng.random((500, 3)) * 40).astype(np.float32)
    for box in (None, BOXES["ortho"]):
        adj = contact_matrix(x, cutoff=6.0, box=box)
        d = oracle.distance_array(x, x, box)
        pairs = oracle.capped_distance(x, x, 6.0, box=box, return_distances=False)
        want = np.zeros((500, 500), dtype=bool)
        want[pairs[:, 0], pairs[:, 1]] = True
        assert np.array_equal(adj, want) and adj.diagonal().all()


def _fastns_distance(oracle, a, b, box):
    """Distance of FastNS.calc_distsq (nsgrid.pyx:483-497) for pairs of atoms: both wrapped into the
    cell first, float32 difference, then the lazy orthorhombic image or the 27-image search."""
    if np.all(box[3:] == 90.0):
        aw, bw = oracle.apply_PBC(a, box), oracle.apply_PBC(b, box)
        d = (aw - bw).astype(np.float64)
        Lh = box[:3].astype(np.float64)
        half = (0.5 * box[:3].astype(np.float64)).astype(np.float32).astype(np.float64)
        d = np.where(d > half, d - Lh, np.where(d <= -half, d + Lh, d))
        return np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    return oracle.calc_bonds(a, b, box)                    # wraps, then minimum_image_triclinic


@pytest.mark.parametrize("angles", [(60, 60, 90), (90, 90, 90)])
def test_fastns_one_million_atoms(oracle, angles):
    """Config C3 at full size: 1 M atoms, 8 A, triclinic [L, L, L, 60, 60, 90] (and its
    orthorhombic twin), ~107 M pairs.  Size-independent properties instead of a full oracle run:
    (1) every unordered pair is reported once;
    (2) a sample of 100 000 reported pairs carries exactly the reference's distance, <= cutoff;
    (3) completeness: for 16 random atoms the oracle's search against ALL atoms finds exactly the
        neighbours the GPU result lists for them;
    (4) search(first 100 000 atoms) returns the ordered pairs of those atoms plus their self pairs."""
    import torch
    from mdanalysis_b200.lib.nsgrid import FastNS
    nn, cutoff = 1_000_000, 8.0
    vol_factor = 0.70710678 if angles[0] == 60 else 1.0
    L = (nn / 0.1 / vol_factor) ** (1.0 / 3.0)
    box = np.array([L, L, L, *angles], dtype=np.float32)
    rng = np.random.default_rng(2026)
    tv = oracle.triclinic_vectors(box).astype(np.float64)
    xyz = (rng.random((nn, 3)) @ tv).astype(np.float32)
    grid = FastNS(cutoff, torch.as_tensor(xyz).cuda(), box)
    res = grid.self_search()
    pairs, dist = res.get_pairs(), res.get_pair_distances()
    del res
    npairs = len(pairs)
    assert 1.06e8 < npairs < 1.09e8                       # n * rho * 4/3 pi r^3 / 2 = 1.072e8
    lo, hi = np.minimum(pairs[:, 0], pairs[:, 1]), np.maximum(pairs[:, 0], pairs[:, 1])
    assert (lo < hi).all()
    skey = np.sort(lo.astype(np.int64) * nn + hi)
    assert (np.diff(skey) > 0).all()                                              # (1)
    s = rng.integers(0, npairs, 100_000)
    want = _fastns_distance(oracle, np.ascontiguousarray(xyz[pairs[s, 0]]), np.ascontiguousarray(xyz[pairs[s, 1]]), box)
    assert np.array_equal(dist[s], want) and (dist[s] <= cutoff).all()           # (2)
    probe = rng.integers(0, nn, 16)
    op, _ = oracle.capped_distance(np.ascontiguousarray(xyz[probe]), xyz, cutoff, box=box, method="nsgrid")
    for t, atom in enumerate(probe):                                              # (3)
        nb = op[op[:, 0] == t][:, 1]
        nb = nb[nb != atom]
        k = np.minimum(nb, atom).astype(np.int64) * nn + np.maximum(nb, atom)
        pos = np.searchsorted(skey, k)
        assert (pos < npairs).all() and np.array_equal(skey[pos], k)
        assert int((lo == atom).sum() + (hi == atom).sum()) == len(nb)
    n_search = len(grid.search(torch.as_tensor(xyz[:100_000]).cuda()))
    assert n_search == int((lo < 100_000).sum() + (hi < 100_000).sum()) + 100_000    # (4)
