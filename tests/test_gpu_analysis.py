"""-m gpu: QCP / RMSD / AlignTraj / InterRDF / contacts through the C ABI vs the oracle."""
import numpy as np
import pytest

from conftest import BOXES

pytestmark = pytest.mark.gpu
RTOL = 1e-5        # BASELINE.json north_star: distances and RMSD within 1e-5 relative


def make_traj(rng, F, N, noise=0.7):
    ref = rng.normal(0, 12, (N, 3))
    frames = []
    for _ in range(F):
        Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        frames.append(ref @ Q + rng.normal(0, 5, 3) + rng.normal(0, noise, (N, 3)))
    return ref.astype(np.float32), np.asarray(frames, dtype=np.float32)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [3, 50, 1000, 20011])
def test_calc_rmsd_rotational_matrix(oracle, dt, n):
    from mdanalysis_b200.lib import qcprot
    rng = np.random.default_rng(n)
    x = rng.normal(0, 10, (n, 3))
    y = x @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 1, (n, 3))
    x, y = (x - x.mean(0)).astype(dt), (y - y.mean(0)).astype(dt)
    w = rng.uniform(1, 16, n).astype(dt)
    tol = 1e-10 if dt == np.float64 else 2e-3       # the reference's own float32 tolerance
    for ww in (None, w):
        r1, r2 = np.zeros(9, dtype=dt), np.zeros(9, dtype=dt)
        v1 = qcprot.CalcRMSDRotationalMatrix(x, y, n, r1, ww)
        v2 = oracle.CalcRMSDRotationalMatrix(x, y, n, r2, ww)
        assert v1 == pytest.approx(v2, rel=tol)
        assert np.allclose(r1, r2, atol=max(tol, 1e-9))
        assert qcprot.CalcRMSDRotationalMatrix(x, y, n, None, ww) == pytest.approx(v2, rel=tol)


def test_qcprot_known_answers():
    """testsuite/MDAnalysisTests/lib/test_qcprot.py:35-46 (test_dummy)."""
    from mdanalysis_b200.lib import qcprot
    for dt, err in ((np.float64, 1e-6), (np.float32, 1e-4)):
        a = np.array([[1.0, 1.0, 2.0]], dtype=dt)
        b = np.array([[1.0, 1.0, 1.0]], dtype=dt)
        rot = np.zeros(9, dtype=dt)
        r = qcprot.CalcRMSDRotationalMatrix(a, b, 1, rot, None)
        assert r == pytest.approx(0.7174389197644676, abs=err)
        assert rot.dtype == dt
        assert np.allclose(rot.reshape(3, 3), np.eye(3))


def test_inner_product_and_fast_calc_twins(oracle):
    """The two public halves of qcprot (InnerProduct :179, FastCalcRMSDAndRotation :326) compose to
    CalcRMSDRotationalMatrix: test_qcprot.py:49-67 (test_regression) drives them this way."""
    from mdanalysis_b200.lib import qcprot
    rng = np.random.default_rng(12)
    n = 777
    x = rng.normal(0, 10, (n, 3))
    y = x @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 1, (n, 3))
    x, y = x - x.mean(0), y - y.mean(0)
    w = rng.uniform(1, 16, n)
    for ww in (None, w):
        A = np.zeros(9)
        E0 = qcprot.InnerProduct(A, y, x, n, ww)                    # CalcRMSDRotationalMatrix: InnerProduct(A, conf, ref, ...)
        wcol = np.ones(n) if ww is None else ww
        assert np.allclose(A.reshape(3, 3), (wcol[:, None] * y).T @ x, rtol=1e-12, atol=1e-9)
        assert E0 == pytest.approx(0.5 * ((wcol[:, None] * y * y).sum() + (wcol[:, None] * x * x).sum()), rel=1e-12)
        rot = np.zeros(9)
        rmsd = qcprot.FastCalcRMSDAndRotation(rot, A, E0, n)
        rot2 = np.zeros(9)
        want = oracle.CalcRMSDRotationalMatrix(x, y, n, rot2, ww)
        assert rmsd == pytest.approx(want, rel=1e-10) and np.allclose(rot, rot2, atol=1e-9)
        assert qcprot.FastCalcRMSDAndRotation(None, A, E0, n) == pytest.approx(want, rel=1e-10)
    # first N atoms only, float32 input (the reference's fused type)
    A32 = np.zeros(9)
    e = qcprot.InnerProduct(A32, y.astype(np.float32), x.astype(np.float32), 100, None)
    y3, x3 = y[:100].astype(np.float32).astype(np.float64), x[:100].astype(np.float32).astype(np.float64)
    assert np.allclose(A32.reshape(3, 3), y3.T @ x3, rtol=1e-10) and e == pytest.approx(0.5 * ((y3 * y3).sum() + (x3 * x3).sum()))


def test_rmsd_function(oracle):
    from mdanalysis_b200.analysis.rms import rmsd
    rng = np.random.default_rng(2)
    a, b = rng.normal(0, 10, (100, 3)), rng.normal(0, 10, (100, 3))
    w = rng.uniform(1, 3, 100)
    assert rmsd(a, b, weights=w, superposition=True) == pytest.approx(oracle.rmsd(a, b, weights=w, superposition=True), rel=1e-12)
    assert rmsd(a, b, center=True) == oracle.rmsd(a, b, center=True)
    with pytest.raises(ValueError):
        rmsd(a, b[:5])


@pytest.mark.parametrize("N,F", [(301, 9), (10000, 6), (20000, 3)])
@pytest.mark.parametrize("weighted", [False, True])
def test_rmsd_trajectory(oracle, N, F, weighted):
    from mdanalysis_b200.analysis.rms import RMSD
    from mdanalysis_b200 import FrameBlock
    rng = np.random.default_rng(N + F)
    ref, traj = make_traj(rng, F, N)
    masses = rng.uniform(1, 32, N)
    blk = FrameBlock(traj, masses=masses)
    R = RMSD(blk, FrameBlock(ref, masses=masses), weights="mass" if weighted else None).run()
    want = oracle.rmsd_trajectory(traj, ref, weights=masses if weighted else None)
    assert R.results.rmsd.shape == (F, 3)
    assert np.array_equal(R.results.rmsd[:, 0], np.arange(F))
    assert np.allclose(R.results.rmsd[:, 2], want, rtol=RTOL, atol=0)
    assert np.allclose(R.results.rmsd[:, 2], want, rtol=1e-9, atol=0)      # what we actually reach
    # frame slicing as AnalysisBase.run(start, stop, step)
    R2 = RMSD(blk, FrameBlock(ref, masses=masses)).run(start=1, step=2)
    assert np.array_equal(R2.results.rmsd[:, 0], np.arange(F)[1::2])
    # device-resident input gives the same numbers
    R3 = RMSD(blk.to_device(), FrameBlock(ref, masses=masses), weights="mass" if weighted else None).run()
    assert np.array_equal(R3.results.rmsd, R.results.rmsd)


def test_rmsd_selection_and_groups(oracle):
    from mdanalysis_b200.analysis.rms import RMSD
    rng = np.random.default_rng(77)
    ref, traj = make_traj(rng, 5, 400)
    sel = np.sort(rng.choice(400, 150, replace=False))
    grp = np.sort(rng.choice(400, 60, replace=False))
    R = RMSD(traj, ref, select=sel, groupselections=[grp]).run()
    want = oracle.rmsd_trajectory(traj[:, sel], ref[sel])
    assert np.allclose(R.results.rmsd[:, 2], want, rtol=1e-9)
    # group RMSD after superposition on `sel` (rms.py:697-721), restated with the oracle pieces
    for f in range(5):
        com = traj[f, sel].astype(np.float64).mean(0)
        ref_com = ref[sel].astype(np.float64).mean(0)
        rot = np.zeros(9)
        oracle.CalcRMSDRotationalMatrix((ref[sel] - ref_com).astype(np.float64),
                                        traj[f, sel].astype(np.float64) - com, len(sel), rot, None)
        x = traj[f].copy()
        x -= com
        x[:] = np.dot(x, rot.reshape(3, 3))
        x += ref_com
        g = np.sqrt(((x[grp].astype(np.float64) - ref[grp].astype(np.float64)) ** 2).sum() / len(grp))
        assert R.results.rmsd[f, 3] == pytest.approx(g, rel=1e-6)


@pytest.mark.parametrize("weighted", [False, True])
def test_align_trajectory(oracle, weighted):
    from mdanalysis_b200.analysis.align import AlignTraj, rotation_matrix
    from mdanalysis_b200 import FrameBlock
    rng = np.random.default_rng(21)
    ref, traj = make_traj(rng, 7, 1500)
    traj[3] = ref                                   # exercises the np.allclose early exit
    masses = rng.uniform(1, 32, 1500)
    sel = np.arange(0, 1500, 3)
    blk = FrameBlock(traj.copy(), masses=masses)
    A = AlignTraj(blk, FrameBlock(ref, masses=masses), select=sel, weights="mass" if weighted else None).run()
    want_xyz, want_rmsd = oracle.align_trajectory(traj, ref[sel], select=sel,
                                                  weights=masses[sel] if weighted else None)
    # frame 3 is the reference itself: its RMSD is rounding noise (~5e-7) in the reference
    assert np.allclose(A.results.rmsd, want_rmsd, rtol=1e-9, atol=2e-6)
    if not weighted:
        assert A.results.rmsd[3] == 0.0
    # the three float32 roundings are replicated: at most 1 ulp(float32) away
    diff = np.abs(blk.positions - want_xyz)
    assert diff.max() <= 4 * np.spacing(np.abs(want_xyz).max().astype(np.float32))
    assert (diff > 0).mean() < 1e-3
    R, v = rotation_matrix(traj[0, sel] - traj[0, sel].mean(0), ref[sel] - ref[sel].mean(0))
    Ro, vo = oracle.rotation_matrix(traj[0, sel] - traj[0, sel].mean(0), ref[sel] - ref[sel].mean(0))
    assert np.allclose(R, Ro, atol=1e-10) and v == pytest.approx(vo, rel=1e-10)


@pytest.mark.parametrize("bname,rmax,nb", [("ortho", 6.0, 75), ("ortho", 12.0, 150), ("tric60", 6.0, 75),
                                            ("tric60", 12.0, 150), ("tric", 9.0, 50)])
@pytest.mark.parametrize("same", [True, False])
def test_interrdf_counts_exact(oracle, bname, rmax, nb, same):
    from mdanalysis_b200.analysis.rdf import InterRDF
    from mdanalysis_b200 import FrameBlock
    rng = np.random.default_rng(int(rmax) + nb)
    box = BOXES[bname]
    tv = oracle.triclinic_vectors(box).astype(np.float64)
    F, N = 3, 700
    pos = np.stack([(rng.random((N, 3)) * 1.4 - 0.2) @ tv for _ in range(F)]).astype(np.float32)
    if same:
        g1 = g2 = FrameBlock(pos, box)
        a = b = pos
    else:
        a, b = np.ascontiguousarray(pos[:, :400]), np.ascontiguousarray(pos[:, 250:])
        g1, g2 = FrameBlock(a, box), FrameBlock(b, box)
    try:
        want = oracle.interrdf(a, b, box, nbins=nb, rrange=(0.0, rmax))
    except ValueError:
        with pytest.raises(ValueError):
            InterRDF(g1, g2, nbins=nb, range=(0.0, rmax)).run()
        return
    got = InterRDF(g1, g2, nbins=nb, range=(0.0, rmax)).run()
    assert np.array_equal(got.results.edges, want["edges"])
    assert np.array_equal(got.results.bins, want["bins"])
    assert np.array_equal(got.results.count, want["count"])        # bit-exact integer counts
    assert np.allclose(got.results.rdf, want["rdf"], rtol=1e-12)
    if same:
        assert got.results.count[0] >= F * N                        # self pairs land in bin 0


def test_interrdf_large_nsgrid_path(oracle):
    """n*m >= 1e8 forces the reference's nsgrid method (lazy minimum image)."""
    from mdanalysis_b200.analysis.rdf import InterRDF
    from mdanalysis_b200 import FrameBlock
    rng = np.random.default_rng(1)
    N, L = 10001, 66.0
    box = np.array([L, L, L, 90, 90, 90], dtype=np.float32)
    pos = (rng.random((2, N, 3)) * L * 1.2 - 0.1 * L).astype(np.float32)
    blk = FrameBlock(pos, box)
    got = InterRDF(blk, blk, nbins=150, range=(0.0, 15.0)).run()
    want = oracle.interrdf(pos, pos, box, nbins=150, rrange=(0.0, 15.0))
    assert np.array_equal(got.results.count, want["count"])
    got2 = InterRDF(blk.to_device(), None, nbins=150, range=(0.0, 15.0), exclusion_block=(1, 1)).run()
    want2 = oracle.interrdf(pos, pos, box, nbins=150, rrange=(0.0, 15.0), exclusion_block=(1, 1))
    assert np.array_equal(got2.results.count, want2["count"])
    assert np.allclose(got2.results.rdf, want2["rdf"], rtol=1e-12)
    assert got2.results.count[0] == want["count"][0] - 2 * N


def test_interrdf_known_answers():
    """testsuite/MDAnalysisTests/analysis/test_rdf.py:39-136: two hand-placed waters."""
    from mdanalysis_b200.analysis.rdf import InterRDF
    from mdanalysis_b200 import FrameBlock
    x, y = [1] * 3 + [2] * 3, [2, 1, 3] * 2
    pos = np.array([[[a, b, 0.0] for a, b in zip(x, y)]], dtype=np.float32)
    box = np.array([100, 100, 100, 90, 90, 90], dtype=np.float32)
    s1 = FrameBlock(pos[:, [0, 3]], box)             # name OW
    s2 = FrameBlock(pos[:, [1, 2, 4, 5]], box)       # name HW1 HW2
    rdf = InterRDF(s1, s2).run()
    assert rdf.results.count.sum() == 8                                   # :69-74
    assert len(rdf.results.count[rdf.results.count == 4]) == 2            # :77-81
    rdf.run()                                                             # :84-89 double run
    assert len(rdf.results.count[rdf.results.count == 4]) == 2
    assert InterRDF(s1, s2, exclusion_block=(1, 2)).run().results.count.sum() == 4   # :92-96
    assert len(InterRDF(s1, s2, nbins=412).run().results.bins) == 412     # :52-57
    r = InterRDF(s1, s2, range=(1.0, 13.0)).run()                         # :60-66
    assert r.results.edges[0] == 1.0 and r.results.edges[-1] == 13.0
    for norm, value in (("density", 1.956823), ("rdf", 244602.88385), ("none", 4)):   # :129-136
        assert max(InterRDF(s1, s2, norm=norm).run().results.rdf) == pytest.approx(value, rel=1e-6)
    with pytest.raises(ValueError, match="invalid norm"):
        InterRDF(s1, s2, norm="foo")
    with pytest.raises(ValueError, match="The exclude_same argument to InterRDF must be"):
        InterRDF(s2, s2, exclude_same="unsupported")


def test_self_contacts(oracle):
    from mdanalysis_b200.analysis.contacts import self_contacts, contact_matrix
    rng = np.random.default_rng(15)
    F, n = 3, 2500
    pos = (rng.random((F, n, 3)) * 60 * 1.5 - 10).astype(np.float32)
    for box in (None, np.array([60, 60, 60, 90, 90, 90], dtype=np.float32), BOXES["tric60"]):
        counts, deg = self_contacts(pos, 8.0, box=box, return_degree=True)
        for f in range(F):
            d = oracle.self_distance_array(pos[f], box)
            assert counts[f] == (d <= 8.0).sum()
            iu = np.triu_indices(n, 1)
            hit = d <= 8.0
            want_deg = np.bincount(iu[0][hit], minlength=n) + np.bincount(iu[1][hit], minlength=n)
            assert np.array_equal(deg[f], want_deg)
    import torch
    c2 = self_contacts(torch.as_tensor(pos).cuda(), 8.0)
    assert np.array_equal(c2.cpu().numpy().astype(np.uint64), self_contacts(pos, 8.0))
    assert contact_matrix(np.array([1.0, 8.0, 8.1]), 8.0).tolist() == [True, True, False]


@pytest.mark.parametrize("n,F", [(2048, 5), (4100, 9), (10000, 3), (2048, 700), (2304, 2500)])
@pytest.mark.parametrize("weighted", [False, True])
def test_qcp_trajectory_streamed_kernel(oracle, n, F, weighted):
    """natoms % 4 == 0 and >= 2048 fitted atoms: the TMA-streamed persistent kernel (qcp.cu,
    qcp_stream_kernel), incl. ragged last chunk, partial batches of frames (one batch in one CTA; 44 CTAs
    with 15 or 16 frames; 148 CTAs with a full batch and a batch of one frame), in-place alignment."""
    import torch
    from mdanalysis_b200.lib import qcprot
    rng = np.random.default_rng(n + F)
    ref, traj = make_traj(rng, F, n)
    traj[-1] = ref                                       # np.allclose early exit of rotation_matrix (unweighted)
    w = rng.uniform(1, 32, n) if weighted else None
    wn = None if w is None else w / w.mean()
    want = oracle.rmsd_trajectory(traj, ref, weights=w)
    for dev in (False, True):
        t = torch.as_tensor(traj).cuda() if dev else traj
        vals, rot, _ = qcprot.qcp_trajectory(t, ref, weights=wn, want_rotation=True)
        vals = vals.cpu().numpy() if dev else vals
        assert np.allclose(vals[:-1], want[:-1], rtol=1e-9, atol=0)
        assert abs(vals[-1] - want[-1]) < 1e-5             # identical structures: cancellation noise only
    moved_want, rm_want = oracle.align_trajectory(traj, ref, weights=w)
    t = traj.copy()
    vals, _, moved = qcprot.qcp_trajectory(t, ref, weights=wn, align=True, convention=1,
                                           allclose_shortcut=not weighted, out=t)
    assert moved is t
    assert np.allclose(vals[:-1], rm_want[:-1], rtol=1e-9, atol=0)
    assert np.abs(t - moved_want).max() <= 8e-6
    if not weighted:
        assert vals[-1] == 0.0 and rm_want[-1] == 0.0      # the shortcut fired on both sides


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ortho_w2", "ortho_w2_ab", "ortho_w2_ab_excl", "ortho_w1", "ortho_w1_ab",
                                  "ortho_perframe", "tric_w1", "tric_w1_ab", "tric_perframe", "tric_flat",
                                  "tric_w2", "tric_w2_ab", "tric_flat_dense", "tric_perframe_dense", "tric_npt_dense"])
def test_interrdf_row_traversal(oracle, case):
    """The row-staged traversal (csrc/rowtile.cuh): every grid flavour it takes, against
    the oracle's FastNS + np.histogram, bit-exact counts.  The dense triclinic cases have
    enough atoms per cell for the library's own finer triclinic grid (W = 2, csrc/rdf.cu
    grid_geometry_triclinic_free), which is only taken when the reference's 27-cell stencil is
    complete for the cell (tric_stencil_complete): the 60/60/90 cell qualifies, the "flat" cells
    (b_x c_y != 0: the reference misses pairs across the z boundary) must stay on its grid."""
    from mdanalysis_b200.analysis.rdf import InterRDF
    from mdanalysis_b200 import FrameBlock
    import zlib
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    F = 3
    kw = {}
    if case.startswith("ortho"):
        if case.startswith("ortho_w1"):
            N, rmax, nb = 4000, 9.0, 90
            dims = np.tile(np.array([40, 37, 52, 90, 90, 90], dtype=np.float32), (F, 1))
        else:
            N, rmax, nb = 9000, 7.0, 140
            dims = np.tile(np.array([47, 52, 61, 90, 90, 90], dtype=np.float32), (F, 1))
        if case == "ortho_perframe":
            dims[1, :3] *= 1.13
            dims[2, :3] *= 0.91
    else:
        N, rmax, nb = 7000, 6.0, 75
        if case.startswith("tric_w2") or case.endswith("_dense"):
            N = 24000
        dims = np.tile(np.array([60, 60, 60, 60, 60, 90], dtype=np.float32), (F, 1))
        if case.startswith("tric_flat"):
            dims[:] = np.array([70, 64, 58, 75, 100, 85], dtype=np.float32)
        if case == "tric_npt_dense":                       # a 60/60/90 cell that breathes: per-frame cells, all of them
            N = 24000                                      # with a complete reference stencil (8 cells of >= 7.35 A)
            dims[1, :3] *= 1.03
            dims[2, :3] *= 1.038
        if case.startswith("tric_perframe"):
            dims[1] = np.array([64, 58, 66, 70, 80, 95], dtype=np.float32)
            dims[2, :3] *= 1.07
    pos = np.empty((F, N, 3), dtype=np.float32)
    for f in range(F):
        tv = oracle.triclinic_vectors(dims[f]).astype(np.float64)
        pos[f] = ((rng.random((N, 3)) * 1.3 - 0.15) @ tv).astype(np.float32)
    if case.endswith("_ab") or case.endswith("_ab_excl"):
        a, b = np.ascontiguousarray(pos[:, : N // 2]), np.ascontiguousarray(pos[:, N // 3:])
        g1, g2 = FrameBlock(a, dims), FrameBlock(b, dims)
        if case.endswith("excl"):
            kw["exclusion_block"] = (3, 2)
    else:
        a = b = pos
        g1 = g2 = FrameBlock(pos, dims)
    want = oracle.interrdf(a, b, dims, nbins=nb, rrange=(0.0, rmax), **kw)
    got = InterRDF(g1, g2, nbins=nb, range=(0.0, rmax), **kw).run()
    assert np.array_equal(got.results.count, want["count"])
    assert np.allclose(got.results.rdf, want["rdf"], rtol=1e-12)
    # device-resident input takes the same path
    g1d = g1.to_device()
    got_d = InterRDF(g1d, g1d if g2 is g1 else g2.to_device(), nbins=nb, range=(0.0, rmax), **kw).run()
    assert np.array_equal(got_d.results.count, want["count"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["dense_groups", "dense_long_rows", "lattice_on_edges", "small_cutoff_range"])
def test_interrdf_quad_kernel_against_row_kernel(oracle, case, monkeypatch):
    """The packed kernel (csrc/rowquad.cuh) against the per-pair row kernel it replaces
    (itself pinned to the oracle above) on inputs the oracle is too slow for: stencil rows
    that need several staging groups, rows longer than the tile, atoms sitting exactly on
    histogram edges (every pair goes through the float64 queue), and a cutoff that is not
    the top edge (the packed kernel must decline)."""
    from mdanalysis_b200.analysis.rdf import InterRDF
    from mdanalysis_b200 import FrameBlock
    import zlib
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    kw = dict(nbins=60, range=(0.0, 3.0))
    if case == "dense_groups":
        N, box = 60000, np.array([30, 31, 29, 90, 90, 90], dtype=np.float32)
        kw = dict(nbins=80, range=(0.0, 4.0))
        pos = (rng.random((2, N, 3)) * box[:3] * 1.1).astype(np.float32)
    elif case == "dense_long_rows":
        N, box = 800000, np.array([20, 22, 24, 90, 90, 90], dtype=np.float32)
        pos = (rng.random((1, N, 3)) * box[:3]).astype(np.float32)
    elif case == "lattice_on_edges":
        # simple cubic lattice with spacing = 2 bin widths: most distances are multiples of edges
        n1, a = 24, 0.5
        g = np.arange(n1, dtype=np.float32) * a
        pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(1, -1, 3).astype(np.float32)
        box = np.array([n1 * a] * 3 + [90, 90, 90], dtype=np.float32)
        kw = dict(nbins=16, range=(0.0, 2.0))
    else:
        N, box = 9000, np.array([47, 52, 61, 90, 90, 90], dtype=np.float32)
        pos = (rng.random((2, N, 3)) * box[:3]).astype(np.float32)
        kw = dict(nbins=140, range=(0.5, 7.0))
    blk = FrameBlock(pos, box).to_device()
    got = InterRDF(blk, None, **kw).run().results.count
    monkeypatch.setenv("MDAGPU_RDF_NO_QUAD", "1")
    want = InterRDF(blk, None, **kw).run().results.count
    monkeypatch.delenv("MDAGPU_RDF_NO_QUAD")
    assert np.array_equal(got, want)
    assert got.sum() > 0
    if case == "lattice_on_edges":
        ref = oracle.interrdf(pos, pos, box, nbins=kw["nbins"], rrange=kw["range"])
        assert np.array_equal(got, ref["count"])
    # host-resident frames go through the upload ring in pieces
    got_h = InterRDF(FrameBlock(pos, box), None, **kw).run().results.count
    assert np.array_equal(got_h, want)


@pytest.mark.gpu
def test_rmsd_aligntraj_full_size_c4(oracle):
    """Config C4 at full size: 50 000 frames of a 10 000-atom structure (6 GB, generated on the
    device): RMSD (plain and mass-weighted) and AlignTraj over all frames in one call each; 40
    sampled frames are checked against the oracle (float64 QCP: rtol 1e-9 observed ~1e-13;
    aligned coordinates within 1 ulp-ish of float32, 8e-6 A at these magnitudes)."""
    import torch
    from mdanalysis_b200.lib import qcprot
    F, N = 50_000, 10_000
    free, _ = torch.cuda.mem_get_info()
    if free < 16 * 2 ** 30:
        pytest.skip("needs 16 GB of free device memory")
    g = torch.Generator(device="cuda")
    g.manual_seed(2026_4)
    ref = torch.randn((N, 3), generator=g, device="cuda") * 20.0
    quat = torch.randn((F, 4), generator=g, device="cuda")
    quat = quat / quat.norm(dim=1, keepdim=True)
    w, x, y, z = quat.unbind(1)
    Q = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1).reshape(F, 3, 3)
    trans = torch.randn((F, 1, 3), generator=g, device="cuda") * 5.0
    traj = torch.empty((F, N, 3), dtype=torch.float32, device="cuda")
    for f0 in range(0, F, 5000):                          # chunks: bounded temporaries
        sl = slice(f0, f0 + 5000)
        traj[sl] = torch.matmul(ref[None], Q[sl]) + trans[sl] + torch.randn((5000, N, 3), generator=g, device="cuda")
    refh = ref.cpu().numpy().astype(np.float32)
    masses = np.random.default_rng(4).uniform(1, 32, N)
    pick = np.sort(np.random.default_rng(5).choice(F, 40, replace=False))
    sample = traj[torch.as_tensor(pick, device="cuda")].cpu().numpy()
    for wts in (None, masses):
        wn = None if wts is None else wts / wts.mean()
        vals, _, _ = qcprot.qcp_trajectory(traj, refh, weights=wn, convention=0)
        vals = vals.cpu().numpy()
        assert vals.shape == (F,) and np.isfinite(vals).all() and 1.5 < vals.mean() < 2.0   # noise sigma 1 per axis
        want = oracle.rmsd_trajectory(sample, refh, weights=wts)
        assert np.allclose(vals[pick], want, rtol=1e-9, atol=0)
    vals, _, moved = qcprot.qcp_trajectory(traj, refh, align=True, convention=1, out=traj)
    assert moved is traj
    aligned_want, rm_want = oracle.align_trajectory(sample, refh)
    assert np.allclose(vals.cpu().numpy()[pick], rm_want, rtol=1e-9, atol=0)
    got = traj[torch.as_tensor(pick, device="cuda")].cpu().numpy()
    assert np.abs(got - aligned_want).max() <= 8e-6
    # after the superposition every frame sits on the reference up to the noise
    dev = (traj[::997] - ref[None]).pow(2).sum(-1).mean(-1).sqrt()
    assert float(dev.max()) < 2.0


@pytest.mark.gpu
def test_interrdf_full_size_c2(oracle):
    """Config C2 at full size (33 334 sites, 1000 frames, 150 bins to 15 A): additivity over frame
    blocks (counts of the whole run == sum of the counts of its halves, through different chunking
    and the host upload ring), the self-pair count, and the first two frames against the oracle."""
    import torch
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rdf import InterRDF
    F, N, L = 1000, 33334, 100.0
    box = np.array([L, L, L, 90, 90, 90], dtype=np.float32)
    g = torch.Generator(device="cuda")
    g.manual_seed(20261018)
    pos = (torch.rand((F, N, 3), generator=g, device="cuda") * L).contiguous()
    blk = FrameBlock(pos, box)
    whole = InterRDF(blk, None, nbins=150, range=(0.0, 15.0)).run().results
    a = InterRDF(blk, None, nbins=150, range=(0.0, 15.0)).run(stop=431).results.count
    b = InterRDF(FrameBlock(pos[431:].cpu().numpy(), box), None, nbins=150, range=(0.0, 15.0)).run().results.count
    assert np.array_equal(whole.count, a + b)
    assert whole.count[0] >= F * N and whole.count.sum() > 1.5e10      # 15.7 M ordered pairs per frame
    assert abs(whole.rdf[100:].mean() - 1.0) < 2e-3                    # ideal gas: g(r) = 1
    two = InterRDF(blk, None, nbins=150, range=(0.0, 15.0)).run(stop=2).results.count
    h = pos[:2].cpu().numpy()
    assert np.array_equal(two, oracle.interrdf(h, h, box, nbins=150, rrange=(0.0, 15.0))["count"])


@pytest.mark.gpu
@pytest.mark.parametrize("bname", ["ortho", "tric60"])
@pytest.mark.parametrize("same", [True, False])
def test_interrdf_exclude_same(oracle, bname, same):
    """exclude_same (rdf.py:280-285): pairs whose atoms carry the same residue / segment / chain
    id are dropped; ids may be any dtype (chain ids are strings)."""
    from mdanalysis_b200.analysis.rdf import InterRDF
    from mdanalysis_b200 import FrameBlock
    rng = np.random.default_rng(17)
    box = BOXES[bname]
    tv = oracle.triclinic_vectors(box).astype(np.float64)
    F, N = 3, 1500
    pos = np.stack([(rng.random((N, 3)) * 1.2 - 0.1) @ tv for _ in range(F)]).astype(np.float32)
    resid = np.repeat(np.arange(N // 3), 3)
    if same:
        a = b = pos
        g1 = g2 = FrameBlock(pos, box)
        ids = (resid, resid)
    else:
        a, b = np.ascontiguousarray(pos[:, :900]), np.ascontiguousarray(pos[:, 600:])
        g1, g2 = FrameBlock(a, box), FrameBlock(b, box)
        ids = (np.array([f"c{r % 7}" for r in resid[:900]]), np.array([f"c{r % 7}" for r in resid[600:]]))
    want = oracle.interrdf(a, b, box, nbins=60, rrange=(0.0, 6.0), exclude_same_ids=ids)
    got = InterRDF(g1, g2, nbins=60, range=(0.0, 6.0), exclude_same=ids).run()
    assert np.array_equal(got.results.count, want["count"])
    assert np.allclose(got.results.rdf, want["rdf"], rtol=1e-12)
    none = InterRDF(g1, g2, nbins=60, range=(0.0, 6.0)).run().results.count
    assert got.results.count.sum() < none.sum()
    got_d = InterRDF(g1.to_device(), g1.to_device() if same else g2.to_device(), nbins=60, range=(0.0, 6.0),
                     exclude_same=ids).run()
    assert np.array_equal(got_d.results.count, want["count"])
    with pytest.raises(ValueError, match="cannot be used with"):
        InterRDF(g1, g2, exclude_same=ids, exclusion_block=(3, 3))
    with pytest.raises(TypeError, match="needs AtomGroups"):
        InterRDF(g1, g2, exclude_same="residue")
