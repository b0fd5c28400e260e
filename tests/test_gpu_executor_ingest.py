"""-m gpu: the plug point of SURVEY.md section 8(b)4 and the ingest of section 8(f)1, against the
REAL reference (baseline/_ref, shipped to the GPU box): ``analysis.run(backend=cuda_backend(n),
unsupported_backend=True)`` for RMSD and for the gpu_enabled() InterRDF / InterRDF_s / Contacts /
AlignTraj objects equals the reference's serial run; a reader's current frame is respected and
restored; DCD trajectories go through the native bulk reader."""
import os
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
DATA = os.path.join(REF, "MDAnalysisTests", "data")


@pytest.fixture(scope="module")
def mda():
    if not os.path.isdir(os.path.join(REF, "MDAnalysis")):
        pytest.skip("baseline/_ref (the built reference) is not present")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import MDAnalysis
    return MDAnalysis


def _universe(mda, xyz, dims=None, masses=None):
    u = mda.Universe.empty(xyz.shape[1], trajectory=True)
    if masses is not None:
        u.add_TopologyAttr("masses", masses)
    u.load_new(xyz.copy(), order="fac", **({} if dims is None else {"dimensions": dims}))
    return u


def _n_workers():
    import torch
    return max(2, min(4, torch.cuda.device_count()))      # >= 2 parts even on one device


def test_executor_interrdf(mda):
    from MDAnalysis.analysis.rdf import InterRDF
    from mdanalysis_b200 import cuda_backend, gpu_enabled
    rng = np.random.default_rng(31)
    F, N = 9, 2500
    dims = np.tile(np.array([32, 30, 34, 90, 90, 90], dtype=np.float32), (F, 1))
    dims[3:, :3] *= 1.05
    xyz = np.stack([(rng.random((N, 3)) * dims[f, :3]).astype(np.float32) for f in range(F)])
    u = _universe(mda, xyz, dims)
    want = InterRDF(u.atoms[:1200], u.atoms[800:], nbins=60, range=(0.0, 9.0)).run()
    with pytest.raises(ValueError):                         # the reference refuses its own class
        InterRDF(u.atoms, u.atoms).run(backend=cuda_backend(2), unsupported_backend=True)
    nw = _n_workers()
    got = gpu_enabled(InterRDF(u.atoms[:1200], u.atoms[800:], nbins=60, range=(0.0, 9.0)))
    got.run(backend=cuda_backend(nw), n_parts=nw + 1, unsupported_backend=True)
    assert np.array_equal(got.results.count, want.results.count)
    assert np.allclose(got.results.rdf, want.results.rdf, rtol=1e-12)
    assert np.array_equal(got.frames, want.frames) and got.n_frames == F
    part = gpu_enabled(InterRDF(u.atoms, u.atoms, nbins=40, range=(0.0, 6.0), exclusion_block=(1, 1)))
    part.run(step=2, backend=cuda_backend(2), unsupported_backend=True)
    ref = InterRDF(u.atoms, u.atoms, nbins=40, range=(0.0, 6.0), exclusion_block=(1, 1)).run(step=2)
    assert np.array_equal(part.results.count, ref.results.count)
    assert np.allclose(part.results.rdf, ref.results.rdf, rtol=1e-12)


def test_executor_rmsd_contacts_aligntraj(mda):
    from MDAnalysis.analysis import align, contacts, rms
    from mdanalysis_b200 import cuda_backend, gpu_enabled
    rng = np.random.default_rng(32)
    F, N = 23, 900
    ref = rng.normal(0, 12, (N, 3))
    xyz = (ref[None] + rng.normal(0, 0.8, (F, N, 3)) + rng.normal(0, 3, (F, 1, 3))).astype(np.float32)
    masses = rng.uniform(1, 32, N)
    nw = _n_workers()
    u = _universe(mda, xyz, None, masses)
    want = rms.RMSD(u, select="all", weights="mass", groupselections=["index 0:99", "index 400:500"]).run()
    got = rms.RMSD(u, select="all", weights="mass", groupselections=["index 0:99", "index 400:500"])
    got.run(backend=cuda_backend(nw), n_parts=5, unsupported_backend=True)
    # frame 0 against itself: both sides return the noise floor of sqrt(2 (E0 - lambda) / N), ~1e-7
    assert np.allclose(got.results.rmsd, want.results.rmsd, rtol=1e-6, atol=5e-6)
    # Contacts
    sel_a, sel_b = u.atoms[:60], u.atoms[60:140]
    refg = (u.atoms[:60], u.atoms[60:140])
    want_c = contacts.Contacts(u, select=("index 0:59", "index 60:139"), refgroup=refg, radius=14.0).run()
    got_c = gpu_enabled(contacts.Contacts(u, select=("index 0:59", "index 60:139"), refgroup=refg, radius=14.0))
    got_c.run(backend=cuda_backend(nw), n_parts=4, unsupported_backend=True)
    assert np.allclose(got_c.results.timeseries, want_c.results.timeseries, rtol=1e-12, atol=1e-12)
    # AlignTraj on the universe itself, the reader parked on a LATER frame: the fit is to that frame
    # (analysis/align.py:836-845 snapshots ref_atoms where run() is called)
    u1, u2 = _universe(mda, xyz, None, masses), _universe(mda, xyz, None, masses)
    u1.trajectory[F - 4]
    u2.trajectory[F - 4]
    want_a = align.AlignTraj(u1, u1, select="all", in_memory=True).run()
    got_a = gpu_enabled(align.AlignTraj(u2, u2, select="all", in_memory=True))
    got_a.run(backend=cuda_backend(nw), n_parts=3, unsupported_backend=True)
    assert np.allclose(got_a.results.rmsd, want_a.results.rmsd, rtol=1e-6, atol=1e-6)
    a1 = u1.trajectory.timeseries(order="fac")
    a2 = u2.trajectory.timeseries(order="fac")
    assert np.abs(a1 - a2).max() <= 1.5e-4                  # the reference's own bar (test_align.py:450)
    assert want_a.results.rmsd[F - 4] < 1e-4 and got_a.results.rmsd[F - 4] < 1e-4


def test_run_on_gpu_respects_and_restores_current_frame(mda):
    from MDAnalysis.analysis import align, rms
    from mdanalysis_b200 import run_on_gpu
    rng = np.random.default_rng(33)
    F, N = 12, 500
    ref = rng.normal(0, 10, (N, 3))
    xyz = (ref[None] + rng.normal(0, 1.0, (F, N, 3))).astype(np.float32)
    u1, u2 = _universe(mda, xyz), _universe(mda, xyz)
    u1.trajectory[7]
    u2.trajectory[7]
    want = align.AlignTraj(u1, u1, select="all", in_memory=True).run()
    got = run_on_gpu(align.AlignTraj(u2, u2, select="all", in_memory=True))
    assert np.allclose(got.results.rmsd, want.results.rmsd, rtol=1e-6, atol=1e-6)
    assert got.results.rmsd[7] < 1e-4
    # RMSD with a selection and a stride reads only what it needs and leaves the reader where it was
    masses = rng.uniform(1, 32, N)                          # the reference's RMSD compares masses (rms.py:539)
    u3 = _universe(mda, xyz, None, masses)
    u3.trajectory[5]
    r = run_on_gpu(rms.RMSD(u3, select="index 10:200", ref_frame=3), step=3)
    w = rms.RMSD(_universe(mda, xyz, None, masses), select="index 10:200", ref_frame=3).run(step=3)
    assert np.allclose(r.results.rmsd, w.results.rmsd, rtol=1e-6, atol=5e-6)
    assert u3.trajectory.ts.frame == 5


def test_dcd_ingest_matches_reference_reader(mda):
    from MDAnalysis.analysis import rms
    from mdanalysis_b200 import FrameBlock, run_on_gpu
    from mdanalysis_b200.coordinates.dcd import DCDFile, read_dcd
    psf, dcd = os.path.join(DATA, "adk.psf"), os.path.join(DATA, "adk_dims.dcd")
    if not os.path.exists(dcd):
        pytest.skip("adk_dims.dcd fixture not shipped")
    u = mda.Universe(psf, dcd)
    blk = read_dcd(dcd)
    assert np.array_equal(blk.positions, u.trajectory.timeseries(order="fac"))
    assert np.allclose(blk.times, [ts.time for ts in u.trajectory], rtol=0, atol=1e-12)
    ca = u.select_atoms("name CA")
    sub = FrameBlock.from_atomgroup(ca, start=3, stop=90, step=7)
    assert np.array_equal(sub.positions, u.trajectory.timeseries(atomgroup=ca, start=3, stop=90, step=7, order="fac"))
    assert np.array_equal(sub.frames, np.arange(3, 90, 7))
    with DCDFile(dcd) as f:
        assert len(f) == 98 and f.header["natoms"] == 3341 and not f.header["is_periodic"]
    for name in ("tip125_tric_C36", "watdyn"):               # unit cells: new-CHARMM box matrix, plain lengths
        p, d = os.path.join(DATA, name + ".psf"), os.path.join(DATA, name + ".dcd")
        if not os.path.exists(d):
            continue
        u2 = mda.Universe(p, d)
        b2 = read_dcd(d)
        assert np.array_equal(b2.positions, u2.trajectory.timeseries(order="fac"))
        assert np.array_equal(b2.dimensions, np.array([ts.dimensions for ts in u2.trajectory], dtype=np.float32))
    # an analysis on the DCD universe: RMSD of the CA atoms, every frame, through the bulk reader
    want = rms.RMSD(u, select="name CA").run()
    got = run_on_gpu(rms.RMSD(mda.Universe(psf, dcd), select="name CA"))
    assert np.allclose(got.results.rmsd, want.results.rmsd, rtol=1e-6, atol=5e-6)     # frame 0 on itself: noise floor
