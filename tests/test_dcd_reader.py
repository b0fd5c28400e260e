"""not gpu: the native bulk DCD reader (csrc/dcd.cu, coordinates/dcd.py) is host code -- checked here
against the unmodified reference's own reader (baseline/_ref, built in this container; skipped when it
or its data files are absent): coordinates bit for bit, unit cells, times, frame ranges, atom
selections, and a trajectory written by the reference's DCDWriter with a triclinic cell."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
DATA = os.path.join(REF, "MDAnalysisTests", "data")


@pytest.fixture(scope="module")
def mda():
    if not os.path.isdir(os.path.join(REF, "MDAnalysis")):
        pytest.skip("baseline/_ref not built")
    sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import MDAnalysis
    yield MDAnalysis
    sys.path.remove(REF)


def test_adk_dcd_matches_reference_reader(mda):
    from mdanalysis_b200.coordinates.dcd import DCDFile, read_dcd
    psf, dcd = os.path.join(DATA, "adk.psf"), os.path.join(DATA, "adk_dims.dcd")
    if not os.path.exists(dcd):
        pytest.skip("adk_dims.dcd fixture absent")
    u = mda.Universe(psf, dcd)
    blk = read_dcd(dcd, pinned=False) if "pinned" in read_dcd.__code__.co_varnames else read_dcd(dcd)
    want = u.trajectory.timeseries(order="fac")
    assert blk.positions.dtype == np.float32 and np.array_equal(np.asarray(blk.positions), want)
    assert np.allclose(blk.times, [ts.time for ts in u.trajectory], rtol=0, atol=1e-12)
    with DCDFile(dcd) as f:
        assert len(f) == 98 and f.header["natoms"] == 3341 and not f.header["is_periodic"]
        idx = np.array([5, 17, 3000, 3340], dtype=np.int64)
        xyz, _ = f.readframes(3, 90, 7, indices=idx)
        assert np.array_equal(np.asarray(xyz), want[3:90:7][:, idx])
        with pytest.raises((IndexError, ValueError, RuntimeError)):
            f.readframes(0, 1, 1, indices=np.array([3341], dtype=np.int64))


@pytest.mark.parametrize("name", ["tip125_tric_C36", "watdyn"])
def test_unit_cells_match_reference_reader(mda, name):
    from mdanalysis_b200.coordinates.dcd import read_dcd
    p, d = os.path.join(DATA, name + ".psf"), os.path.join(DATA, name + ".dcd")
    if not os.path.exists(d):
        pytest.skip(name + ".dcd fixture absent")
    u = mda.Universe(p, d)
    blk = read_dcd(d)
    assert np.array_equal(np.asarray(blk.positions), u.trajectory.timeseries(order="fac"))
    assert np.array_equal(blk.dimensions, np.array([ts.dimensions for ts in u.trajectory], dtype=np.float32))


def test_round_trip_through_reference_writer(mda, tmp_path):
    """A DCD written by the reference (triclinic cell, 7 frames) read back by the native reader."""
    from mdanalysis_b200.coordinates.dcd import read_dcd
    rng = np.random.default_rng(3)
    n, F = 257, 7
    xyz = (rng.random((F, n, 3)) * 40).astype(np.float32)
    u = mda.Universe.empty(n, trajectory=True)
    dims = np.array([41.0, 39.5, 44.25, 70.0, 80.0, 100.0], dtype=np.float32)
    u.load_new(xyz, order="fac", dimensions=dims)
    path = str(tmp_path / "w.dcd")
    with mda.Writer(path, n_atoms=n) as w:
        for _ in u.trajectory:
            w.write(u.atoms)
    back = mda.Universe.empty(n, trajectory=True)
    back.load_new(path)
    blk = read_dcd(path)
    assert np.array_equal(np.asarray(blk.positions), back.trajectory.timeseries(order="fac"))
    assert np.allclose(blk.dimensions, [ts.dimensions for ts in back.trajectory], rtol=0, atol=1e-4)
    assert blk.n_frames == F and blk.n_atoms == n


def test_missing_file_is_an_error():
    from mdanalysis_b200.coordinates.dcd import DCDFile
    with pytest.raises((IOError, OSError, RuntimeError, ValueError)):
        DCDFile("/nonexistent/none.dcd")
