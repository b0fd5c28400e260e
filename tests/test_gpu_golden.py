"""-m gpu: the CUDA path (through the C ABI and the host mirror of the reference API)
against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Bars (BASELINE.json north_star): bit-exact for
distances / wraps / pair sets / pair distances / RDF counts; 1e-5 relative for
RMSD (measured: ~1e-12); angles 1e-12 absolute (atan2 is not correctly rounded);
aligned float32 coordinates within 1 ulp of the coordinate magnitude (4e-6 A here;
the reference's own check uses 1.5e-4 A, test_align.py:450)."""
import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu

RTOL_RMSD = 1e-9          # far inside the 1e-5 bar
POS_ATOL = 8e-6


class CudaImpl:
    def __init__(self):
        from mdanalysis_b200.analysis import align, rdf, rms
        from mdanalysis_b200.lib import distances, nsgrid, qcprot
        self.gd, self.rms, self.align, self.rdf = distances, rms, align, rdf
        self.FastNS = nsgrid.FastNS
        self.CalcRMSDRotationalMatrix = qcprot.CalcRMSDRotationalMatrix
        self.rmsd = rms.rmsd

    def distance_array(self, a, b, box):
        return self.gd.distance_array(a, b, box=box)

    def self_distance_array(self, a, box):
        return self.gd.self_distance_array(a, box=box)

    def calc_bonds(self, a, b, box):
        return self.gd.calc_bonds(a, b, box=box)

    def calc_angles(self, a, b, c, box):
        return self.gd.calc_angles(a, b, c, box=box)

    def calc_dihedrals(self, a, b, c, d, box):
        return self.gd.calc_dihedrals(a, b, c, d, box=box)

    def apply_PBC(self, a, box):
        return self.gd.apply_PBC(a, box)

    def transform_RtoS(self, a, box):
        return self.gd.transform_RtoS(a, box)

    def transform_StoR(self, a, box):
        return self.gd.transform_StoR(a, box)

    def capped_distance(self, *a, **k):
        return self.gd.capped_distance(*a, **k)

    def self_capped_distance(self, *a, **k):
        return self.gd.self_capped_distance(*a, **k)

    def rmsd_trajectory(self, traj, ref, weights=None):
        return self.rms.RMSD(traj, ref, weights=weights).run().results.rmsd[:, 2]

    def align_trajectory(self, traj, ref, select=None, weights=None):
        t = traj.copy()
        full_ref = ref
        if select is not None:                       # AlignTraj wants the reference of all atoms
            full_ref = np.zeros((traj.shape[1], 3), dtype=np.float32)
            full_ref[select] = ref
        a = self.align.AlignTraj(t, full_ref, select="all" if select is None else select, weights=weights,
                                 in_memory=True).run()
        return np.asarray(a.results.positions), np.asarray(a.results.rmsd)

    def interrdf(self, a, b, box, nbins, rrange, exclusion_block=None):
        r = self.rdf.InterRDF(a, b, nbins=nbins, range=rrange, exclusion_block=exclusion_block,
                              dimensions=box).run()
        return {"count": r.results.count, "edges": r.results.edges, "rdf": r.results.rdf}


@pytest.fixture(scope="module")
def impl():
    return CudaImpl()


@pytest.mark.parametrize("bname", list(gc.BOXES))
@pytest.mark.parametrize("tag", ["in", "far"])
def test_distances(impl, bname, tag):
    gc.check_distances(impl, bname, tag)


def test_edge_wraps(impl):
    gc.check_edge_wraps(impl)


@pytest.mark.parametrize("bname", list(gc.ADK_BOXES))
def test_adk_config_c1(impl, bname):
    gc.check_adk(impl, bname)


@pytest.mark.parametrize("case", ["ortho_4.0", "ortho_7.3", "tric_6.0", "tric60_5.0"])
def test_fastns(impl, case):
    gc.check_fastns(impl, case)


def test_fastns_nopbc(impl):
    gc.check_fastns_nopbc(impl)


@pytest.mark.parametrize("bname", list(gc.BOXES))
@pytest.mark.parametrize("method", ["bruteforce", "nsgrid", "pkdtree"])
@pytest.mark.parametrize("use_min", [False, True])
def test_capped(impl, bname, method, use_min):
    gc.check_capped(impl, bname, method, use_min)


def test_qcp(impl):
    gc.check_qcp(impl, rtol=RTOL_RMSD)


def test_rmsd_function(impl):
    gc.check_rmsd_fn(impl, rtol=RTOL_RMSD)


@pytest.mark.parametrize("wname", ["none", "mass"])
def test_rmsd_and_aligntraj(impl, wname):
    gc.check_trajectory(impl, wname, rtol=RTOL_RMSD, pos_atol=POS_ATOL)


def test_aligntraj_selection(impl):
    gc.check_trajectory_selection(impl, rtol=RTOL_RMSD, pos_atol=POS_ATOL)


@pytest.mark.parametrize("case", gc.rdf_cases())
def test_interrdf(impl, case):
    gc.check_rdf(impl, case)
