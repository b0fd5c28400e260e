"""-m "not gpu": the CPU oracle against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  This is what pins the oracle on the GPU box,
where /root/reference does not exist."""
import numpy as np
import pytest

import golden_cases as gc


class OracleImpl:
    """Adapter: oracle.py functions under the names golden_cases uses."""

    def __init__(self, orc):
        self.o = orc
        for name in ("distance_array", "self_distance_array", "calc_bonds", "calc_angles", "calc_dihedrals",
                     "apply_PBC", "transform_RtoS", "transform_StoR", "FastNS", "capped_distance",
                     "self_capped_distance", "CalcRMSDRotationalMatrix", "rmsd", "rmsd_trajectory",
                     "align_trajectory"):
            setattr(self, name, getattr(orc, name))

    def interrdf(self, a, b, box, nbins, rrange, exclusion_block=None):
        return self.o.interrdf(a, a if b is None else b, box, nbins=nbins, rrange=rrange,
                               exclusion_block=exclusion_block)


@pytest.fixture(scope="module")
def impl(oracle):
    return OracleImpl(oracle)


@pytest.mark.parametrize("bname", list(gc.BOXES))
@pytest.mark.parametrize("tag", ["in", "far"])
def test_distances(impl, bname, tag):
    gc.check_distances(impl, bname, tag, exact_angles=True)     # same libm as the reference: bit-exact


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
    gc.check_qcp(impl, rtol=0.0)            # bit-exact restatement


def test_rmsd_function(impl):
    gc.check_rmsd_fn(impl, rtol=0.0)


@pytest.mark.parametrize("wname", ["none", "mass"])
def test_rmsd_and_aligntraj(impl, wname):
    gc.check_trajectory(impl, wname, rtol=0.0, pos_atol=0.0)


def test_aligntraj_selection(impl):
    gc.check_trajectory_selection(impl, rtol=0.0, pos_atol=0.0)


@pytest.mark.parametrize("case", gc.rdf_cases())
def test_interrdf(impl, case):
    gc.check_rdf(impl, case)
