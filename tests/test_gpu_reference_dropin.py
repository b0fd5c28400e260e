"""-m gpu: the CUDA backend plugged into the REAL, unmodified reference.

baseline/_ref holds MDAnalysis 2.8.0-dev0 built out of tree from /root/reference (git-ignored,
shipped to the GPU box by gpurun; recipe in DESIGN.md).  When it is there, these tests call the
reference's own front ends with ``backend='cuda'`` (mdanalysis_b200.register.install) next to
``backend='serial'``, and run the reference's analysis classes on a MemoryReader universe next
to the GPU classes fed from the same universe.  Skipped when baseline/_ref is absent.
"""
import os
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def mda():
    if not os.path.isdir(os.path.join(REF, "MDAnalysis")):
        pytest.skip("baseline/_ref (the built reference) is not present")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import MDAnalysis
    import mdanalysis_b200.register as reg
    reg.install(patch_capped=True)
    return MDAnalysis


BOXES = {
    "none": None,
    "ortho": np.array([30, 25, 40, 90, 90, 90], dtype=np.float32),
    "tric": np.array([40, 35, 50, 70, 80, 100], dtype=np.float32),
}


@pytest.mark.parametrize("bname", list(BOXES))
def test_lib_distances_backend_cuda(mda, bname):
    import MDAnalysis.lib.distances as D
    assert "cuda" in D._distances
    rng = np.random.default_rng(5)
    box = BOXES[bname]
    a, b, c, d = [((rng.random((n, 3)) * 5 - 2) * 35).astype(np.float32) for n in (301, 257, 301, 301)]
    for args, fn in (((a, b), D.distance_array), ((a,), D.self_distance_array), ((a, c), D.calc_bonds)):
        assert np.array_equal(fn(*args, box=box, backend="cuda"), fn(*args, box=box, backend="serial"))
    for args, fn in (((a, c, d), D.calc_angles), ((a, c, d, a[::-1].copy()), D.calc_dihedrals)):
        got, want = fn(*args, box=box, backend="cuda"), fn(*args, box=box, backend="serial")
        assert np.allclose(got, want, rtol=1e-5, atol=1e-12, equal_nan=True)      # atan2: <= 1e-5 relative
    if box is not None:
        for fn in (D.apply_PBC, D.transform_RtoS, D.transform_StoR):
            assert np.array_equal(fn(a, box, backend="cuda"), fn(a, box, backend="serial"))
    # results written into a caller-owned (here: non-contiguous) array, as _check_result_array allows
    res = np.zeros((301, 514), dtype=np.float64)[:, ::2]
    D.distance_array(a, b, box=box, result=res, backend="cuda")
    assert np.array_equal(res, D.distance_array(a, b, box=box, backend="serial"))
    with pytest.raises(ValueError):
        D.distance_array(a, b, box=box, backend="nope")


@pytest.mark.parametrize("bname", list(BOXES))
@pytest.mark.parametrize("method", ["bruteforce", "nsgrid", "pkdtree"])
def test_capped_distance_backend_cuda(mda, bname, method):
    import MDAnalysis.lib.distances as D
    rng = np.random.default_rng(9)
    box = BOXES[bname]
    a, b = (rng.random((400, 3)) * 30).astype(np.float32), (rng.random((300, 3)) * 30).astype(np.float32)

    def canon(p, d, self_pairs=False):
        p = np.asarray(p).reshape(-1, 2)
        if self_pairs:
            p = np.sort(p, axis=1)
        o = np.lexsort((p[:, 1], p[:, 0]))
        return p[o], np.asarray(d)[o]
    for mn in (None, 1.5):
        pw, dw = canon(*D.capped_distance(a, b, 5.0, min_cutoff=mn, box=box, method=method))
        pg, dg = canon(*D.capped_distance(a, b, 5.0, min_cutoff=mn, box=box, method=method, backend="cuda"))
        assert np.array_equal(pg, pw) and np.allclose(dg, dw, rtol=1e-12)
        pw, dw = canon(*D.self_capped_distance(a, 5.0, min_cutoff=mn, box=box, method=method), True)
        pg, dg = canon(*D.self_capped_distance(a, 5.0, min_cutoff=mn, box=box, method=method, backend="cuda"), True)
        assert np.array_equal(pg, pw) and np.allclose(dg, dw, rtol=1e-12)


def _universe(mda, xyz, dims, masses=None):
    u = mda.Universe.empty(xyz.shape[1], trajectory=True)
    if masses is not None:
        u.add_TopologyAttr("masses", masses)
    u.load_new(xyz.copy(), order="fac", **({} if dims is None else {"dimensions": dims}))
    return u


def test_interrdf_class_against_reference(mda):
    from MDAnalysis.analysis.rdf import InterRDF as RefRDF, InterRDF_s as RefRDFs
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rdf import InterRDF, InterRDF_s
    rng = np.random.default_rng(21)
    F, N = 3, 2500
    box = np.array([32, 34, 30, 90, 90, 90], dtype=np.float32)
    pos = (rng.random((F, N, 3)) * box[:3] * 1.1).astype(np.float32)
    u = _universe(mda, pos, np.tile(box, (F, 1)))
    for kw in ({}, {"exclusion_block": (1, 1)}, {"norm": "density"}):
        want = RefRDF(u.atoms, u.atoms, nbins=90, range=(0.0, 9.0), **kw).run()
        blk = FrameBlock.from_atomgroup(u.atoms)                # bulk ingest: timeseries(order='fac')
        got = InterRDF(blk, blk, nbins=90, range=(0.0, 9.0), **kw).run()
        assert np.array_equal(got.results.count, want.results.count)
        assert np.array_equal(got.results.edges, want.results.edges)
        assert np.allclose(got.results.rdf, want.results.rdf, rtol=1e-12)
    sites = [[np.array([0, 1]), np.arange(50, 400)]]
    want = RefRDFs(u, [[u.atoms[a], u.atoms[b]] for a, b in sites], nbins=30, range=(0.0, 6.0)).run()
    got = InterRDF_s(FrameBlock.from_atomgroup(u.atoms), sites, nbins=30, range=(0.0, 6.0)).run()
    assert np.array_equal(got.results.count[0], want.results.count[0])
    assert np.allclose(got.results.rdf[0], want.results.rdf[0], rtol=1e-12)


def test_rmsd_aligntraj_contacts_classes_against_reference(mda):
    from MDAnalysis.analysis import rms as rrms, align as ralign, contacts as rcontacts
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rms import RMSD
    from mdanalysis_b200.analysis.align import AlignTraj
    from mdanalysis_b200.analysis.contacts import Contacts
    rng = np.random.default_rng(33)
    F, N = 12, 900
    ref_xyz = rng.normal(0, 12, (N, 3))
    traj = np.stack([(ref_xyz @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 4, 3)
                      + rng.normal(0, 0.7, (N, 3))) for _ in range(F)]).astype(np.float32)
    masses = rng.uniform(1, 32, N)
    r = _universe(mda, ref_xyz.astype(np.float32)[None], None, masses)
    for w in (None, "mass"):
        u = _universe(mda, traj, None, masses)
        want = rrms.RMSD(u, r, select="all", weights=w).run().results.rmsd
        got = RMSD(FrameBlock.from_atomgroup(u.atoms), FrameBlock.from_atomgroup(r.atoms), weights=w).run().results.rmsd
        assert np.allclose(got, want, rtol=1e-5, atol=0)        # north_star tolerance; observed ~1e-12
        assert np.allclose(got[:, 2], want[:, 2], rtol=1e-9)
        u2 = _universe(mda, traj, None, masses)
        al = ralign.AlignTraj(u2, r, select="all", weights=w, in_memory=True).run()
        blk = FrameBlock.from_atomgroup(_universe(mda, traj, None, masses).atoms)
        g = AlignTraj(blk, FrameBlock.from_atomgroup(r.atoms), weights=w).run()
        assert np.allclose(g.results.rmsd, al.results.rmsd, rtol=1e-9)
        moved = u2.trajectory.timeseries(order="fac")
        assert np.abs(np.asarray(blk.positions) - moved).max() <= 1.5e-4      # test_align.py:450 tolerance
    box = np.array([60, 60, 60, 90, 90, 90], dtype=np.float32)
    u = _universe(mda, traj + 30, np.tile(box, (F, 1)), masses)
    gA, gB = u.atoms[:300], u.atoms[500:800]
    u.trajectory[0]
    for method in ("hard_cut", "soft_cut", "radius_cut"):
        want = rcontacts.Contacts(u, select=(gA, gB), refgroup=(gA, gB), method=method, radius=8.0).run()
        u.trajectory[0]
        blk = FrameBlock.from_atomgroup(u.atoms)
        refA = FrameBlock(blk.positions[:1, :300], blk.dimensions[:1])
        refB = FrameBlock(blk.positions[:1, 500:800], blk.dimensions[:1])
        got = Contacts(blk, select=(np.arange(300), np.arange(500, 800)), refgroup=(refA, refB), method=method,
                       radius=8.0).run()
        assert np.allclose(got.results.timeseries, want.results.timeseries, rtol=1e-12)


def test_run_on_gpu_fills_reference_objects(mda):
    """mdanalysis_b200.run_on_gpu(obj) in place of obj.run(): the reference objects, built from
    AtomGroups as usual, end up with the results the reference computes itself."""
    from MDAnalysis.analysis.rdf import InterRDF, InterRDF_s
    from MDAnalysis.analysis import rms as rrms, align as ralign, contacts as rcontacts
    from mdanalysis_b200 import run_on_gpu
    rng = np.random.default_rng(77)
    F, N = 8, 1200
    box = np.array([36, 33, 30, 90, 90, 90], dtype=np.float32)
    pos = (rng.random((F, N, 3)) * box[:3]).astype(np.float32)
    masses = rng.uniform(1, 32, N)
    dims = np.tile(box, (F, 1))
    u = _universe(mda, pos, dims, masses)
    # InterRDF, with a frame stride and an exclusion block
    for kw, run_kw in (({}, {}), ({"exclusion_block": (4, 4)}, {"step": 2}), ({"norm": "none"}, {"start": 1, "stop": 6})):
        want = InterRDF(u.atoms[:600], u.atoms[400:], nbins=60, range=(0.0, 9.0), **kw).run(**run_kw)
        got = run_on_gpu(InterRDF(u.atoms[:600], u.atoms[400:], nbins=60, range=(0.0, 9.0), **kw), **run_kw)
        assert np.array_equal(got.results.count, want.results.count)
        assert np.allclose(got.results.rdf, want.results.rdf, rtol=1e-12)
        assert np.array_equal(got.frames, want.frames) and got.n_frames == want.n_frames
    ags = [[u.atoms[[1, 2]], u.atoms[100:300]]]
    want = InterRDF_s(u, ags, nbins=25, range=(0.0, 7.0)).run()
    got = run_on_gpu(InterRDF_s(u, ags, nbins=25, range=(0.0, 7.0)))
    assert np.array_equal(got.results.count[0], want.results.count[0])
    assert np.allclose(got.results.rdf[0], want.results.rdf[0], rtol=1e-12)
    assert np.allclose(got.get_cdf()[0], want.get_cdf()[0], rtol=1e-12)
    # RMSD with a sub-selection, group selections and mass weights
    ref_xyz = rng.normal(0, 10, (N, 3))
    traj = np.stack([(ref_xyz @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 3, 3)
                      + rng.normal(0, 0.6, (N, 3))) for _ in range(F)]).astype(np.float32)
    r = _universe(mda, ref_xyz.astype(np.float32)[None], None, masses)
    kw = dict(select="index 0:799", groupselections=["index 800:999", "index 1000:1199"], weights="mass")
    want = rrms.RMSD(_universe(mda, traj, None, masses), r, **kw).run(step=2)
    got = run_on_gpu(rrms.RMSD(_universe(mda, traj, None, masses), r, **kw), step=2)
    assert np.allclose(got.results.rmsd, want.results.rmsd, rtol=1e-6, atol=0)
    # AlignTraj rewrites the in-memory trajectory
    u_ref, u_gpu = _universe(mda, traj, None, masses), _universe(mda, traj, None, masses)
    want = ralign.AlignTraj(u_ref, r, select="index 0:599", in_memory=True).run()
    got = run_on_gpu(ralign.AlignTraj(u_gpu, r, select="index 0:599", in_memory=True))
    assert np.allclose(got.results.rmsd, want.results.rmsd, rtol=1e-9)
    assert np.abs(u_gpu.trajectory.timeseries(order="fac") - u_ref.trajectory.timeseries(order="fac")).max() <= 1.5e-4
    u_gpu.trajectory[3]
    u_ref.trajectory[3]
    assert np.abs(u_gpu.atoms.positions - u_ref.atoms.positions).max() <= 1.5e-4
    # Contacts with the soft cut-off
    uc = _universe(mda, traj + 25, np.tile(np.array([55, 55, 55, 90, 90, 90], dtype=np.float32), (F, 1)), masses)
    gA, gB = uc.atoms[:300], uc.atoms[700:1000]
    for method in ("hard_cut", "soft_cut", "radius_cut"):
        uc.trajectory[0]
        want = rcontacts.Contacts(uc, select=(gA, gB), refgroup=(gA, gB), method=method, radius=9.0).run()
        uc.trajectory[0]
        got = run_on_gpu(rcontacts.Contacts(uc, select=(gA, gB), refgroup=(gA, gB), method=method, radius=9.0))
        assert np.allclose(got.results.timeseries, want.results.timeseries, rtol=1e-12)
    with pytest.raises(TypeError, match="no GPU path"):
        run_on_gpu(object())


def test_cuda_backend_executor_for_analysisbase_run(mda):
    """Plug point 4: RMSD.run(backend=CudaBackend, unsupported_backend=True) -- the reference's own
    run() splits the frames, our executor computes each part on the GPU, the reference's
    aggregator (ndarray_vstack) merges them."""
    from MDAnalysis.analysis import rms as rrms
    from MDAnalysis.analysis.rdf import InterRDF
    from mdanalysis_b200.adapters import cuda_backend
    rng = np.random.default_rng(4)
    F, N = 10, 500
    ref_xyz = rng.normal(0, 10, (N, 3))
    traj = np.stack([(ref_xyz + rng.normal(0, 0.8, (N, 3))) for _ in range(F)]).astype(np.float32)
    masses = rng.uniform(1, 32, N)
    r = _universe(mda, ref_xyz.astype(np.float32)[None], None, masses)
    want = rrms.RMSD(_universe(mda, traj, None, masses), r, select="all").run()
    for parts in (1, 3):
        got = rrms.RMSD(_universe(mda, traj, None, masses), r, select="all").run(
            backend=cuda_backend(n_workers=parts), n_parts=parts, unsupported_backend=True)
        assert np.allclose(got.results.rmsd, want.results.rmsd, rtol=1e-9, atol=0)
        assert np.array_equal(got.frames, want.frames)
    u = _universe(mda, traj, np.tile(np.array([80, 80, 80, 90, 90, 90], dtype=np.float32), (F, 1)), masses)
    with pytest.raises(ValueError):                  # the reference itself refuses: not parallelizable
        InterRDF(u.atoms, u.atoms).run(backend=cuda_backend(), unsupported_backend=True)


def test_config_c1_adk_all_frames(mda):
    """Config C1 as BASELINE.json states it: distance_array + self_distance_array on the
    MDAnalysisTests adk PSF/DCD (3341 atoms, 98 frames), read by the reference itself; every frame,
    no box / orthorhombic / triclinic, backend='cuda' against the reference's OpenMP build
    (bit-identical to its serial build; used here only because it is 10x faster on the host)."""
    import MDAnalysis.lib.distances as D
    psf = os.path.join(REF, "MDAnalysisTests", "data", "adk.psf")
    dcd = os.path.join(REF, "MDAnalysisTests", "data", "adk_dims.dcd")
    if not (os.path.exists(psf) and os.path.exists(dcd)):
        pytest.skip("adk PSF/DCD not present in baseline/_ref")
    u = mda.Universe(psf, dcd)
    assert u.atoms.n_atoms == 3341 and len(u.trajectory) == 98
    boxes = [None, np.array([50, 50, 50, 90, 90, 90], dtype=np.float32),
             np.array([50, 50, 50, 70.53571, 109.48542, 70.518196], dtype=np.float32)]
    u.trajectory[0]
    x0 = u.atoms.positions.copy()
    cpu = "openmp" if "openmp" in D._distances else "serial"
    for ts in u.trajectory:
        x = u.atoms.positions
        for box in boxes:
            d = D.distance_array(x0, x, box=box, backend="cuda")
            assert np.array_equal(d, D.distance_array(x0, x, box=box, backend=cpu))
            if ts.frame % 14 == 0:
                assert np.array_equal(D.self_distance_array(x, box=box, backend="cuda"),
                                      D.self_distance_array(x, box=box, backend=cpu))
    # the constants the reference's own test pins (lib/test_distances.py:439-454): frame 0 vs frame 10,
    # to the 5 decimals that test asks for
    u.trajectory[10]
    d = D.distance_array(x0, u.atoms.positions, backend="cuda")
    assert d.min() == pytest.approx(0.11981228170520701, abs=1e-5)
    assert d.max() == pytest.approx(53.572192429459619, abs=1e-5)


def test_interrdf_exclude_same_against_reference(mda):
    """exclude_same='residue' on a universe with residues (three atoms each), reference vs
    run_on_gpu (which reads the attribute the reference constructor resolved)."""
    from MDAnalysis.analysis.rdf import InterRDF
    from mdanalysis_b200 import run_on_gpu
    rng = np.random.default_rng(91)
    F, nres = 4, 500
    N = 3 * nres
    box = np.array([28, 30, 29, 90, 90, 90], dtype=np.float32)
    pos = (rng.random((F, N, 3)) * box[:3]).astype(np.float32)
    u = mda.Universe.empty(N, n_residues=nres, atom_resindex=np.repeat(np.arange(nres), 3), trajectory=True)
    u.load_new(pos.copy(), order="fac", dimensions=np.tile(box, (F, 1)))
    ow, hw = u.atoms[::3], u.atoms[np.arange(N) % 3 != 0]
    for g1, g2 in ((ow, hw), (u.atoms, u.atoms)):
        want = InterRDF(g1, g2, nbins=50, range=(0.0, 7.0), exclude_same="residue").run()
        got = run_on_gpu(InterRDF(g1, g2, nbins=50, range=(0.0, 7.0), exclude_same="residue"))
        assert np.array_equal(got.results.count, want.results.count)
        assert np.allclose(got.results.rdf, want.results.rdf, rtol=1e-12)


def test_average_structure_against_reference(mda):
    from MDAnalysis.analysis import align as ralign
    from mdanalysis_b200 import FrameBlock, run_on_gpu
    from mdanalysis_b200.analysis.align import AverageStructure
    rng = np.random.default_rng(12)
    F, N = 15, 700
    ref_xyz = rng.normal(0, 9, (N, 3))
    traj = np.stack([(ref_xyz @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 3, 3)
                      + rng.normal(0, 0.5, (N, 3))) for _ in range(F)]).astype(np.float32)
    masses = rng.uniform(1, 32, N)

    def full(xyz):          # AverageStructure merges the selection into a new universe: needs a full topology
        u = mda.Universe.empty(N, n_residues=N, atom_resindex=np.arange(N), n_segments=1,
                               residue_segindex=np.zeros(N, dtype=int), trajectory=True)
        u.add_TopologyAttr("masses", masses)
        for attr, val in (("names", ["CA"] * N), ("types", ["C"] * N), ("resnames", ["ALA"] * N),
                          ("resids", np.arange(1, N + 1)), ("segids", ["A"])):
            u.add_TopologyAttr(attr, val)
        u.load_new(xyz.copy(), order="fac")
        return u
    r = full(ref_xyz.astype(np.float32)[None])
    for sel, w in (("all", None), ("index 0:399", "mass")):
        want = ralign.AverageStructure(full(traj), r, select=sel, weights=w).run()
        got = run_on_gpu(ralign.AverageStructure(full(traj), r, select=sel, weights=w))
        assert np.allclose(got.results.positions, want.results.positions, rtol=0, atol=2e-5)
        assert got.results.rmsd == pytest.approx(want.results.rmsd, rel=1e-9)
        assert np.allclose(got.results.universe.atoms.positions, want.results.universe.atoms.positions, atol=2e-5)
    idx = np.arange(400)
    g = AverageStructure(FrameBlock(traj.copy(), masses=masses), FrameBlock(ref_xyz.astype(np.float32)[None], masses=masses),
                         select=idx, weights="mass").run()
    assert np.allclose(g.results.positions, want.results.positions, rtol=0, atol=2e-5)


def test_distance_matrix_against_reference(mda):
    """diffusionmap.DistanceMatrix of the real reference (a Python loop over frame pairs) next to the
    GPU class and to run_on_gpu on the reference object; plain and superposed metric, weights, frames=."""
    import functools

    from MDAnalysis.analysis import diffusionmap as rdm
    from MDAnalysis.analysis.rms import rmsd as ref_rmsd
    from mdanalysis_b200 import run_on_gpu
    from mdanalysis_b200.analysis.diffusionmap import DistanceMatrix, rmsd_superposition
    rng = np.random.default_rng(14)
    F, N = 37, 150
    base = rng.normal(0, 9, (N, 3))
    traj = np.stack([(base @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 3, 3)
                      + rng.normal(0, 0.8, (N, 3))) for _ in range(F)]).astype(np.float32)
    w = rng.uniform(1, 32, N)
    fit = functools.partial(ref_rmsd, superposition=True)
    for metric, gpu_metric, weights, tol in ((ref_rmsd, ref_rmsd, None, 0.0), (ref_rmsd, ref_rmsd, w, 0.0),
                                             (fit, rmsd_superposition, None, 5e-6), (fit, fit, w, 5e-6)):
        want = rdm.DistanceMatrix(_universe(mda, traj, None), metric=metric, weights=weights).run().results.dist_matrix
        got = DistanceMatrix(traj, metric=gpu_metric, weights=weights).run().results.dist_matrix
        np.testing.assert_allclose(got, want, rtol=1e-7, atol=tol)
        obj = run_on_gpu(rdm.DistanceMatrix(_universe(mda, traj, None), metric=metric, weights=weights))
        np.testing.assert_allclose(obj.results.dist_matrix, want, rtol=1e-7, atol=tol)
        assert obj.n_frames == F and obj._calculated
    frames = [3, 4, 9, 30]
    want = rdm.DistanceMatrix(_universe(mda, traj, None), select="index 0:99").run(frames=frames).results.dist_matrix
    got = run_on_gpu(rdm.DistanceMatrix(_universe(mda, traj, None), select="index 0:99"), frames=frames)
    np.testing.assert_allclose(got.results.dist_matrix, want, rtol=1e-9)
    # the reference's DiffusionMap runs on top of a matrix computed here
    dmap = rdm.DiffusionMap(run_on_gpu(rdm.DistanceMatrix(_universe(mda, traj, None))), epsilon=2.0)
    dmap.run()
    ref_map = rdm.DiffusionMap(_universe(mda, traj, None), epsilon=2.0)
    ref_map.run()
    np.testing.assert_allclose(dmap.eigenvalues, ref_map.eigenvalues, rtol=1e-6, atol=1e-9)
