"""-m "not gpu": the CPU oracle against the second golden set (minimize_vectors, InterRDF_s,
Contacts) produced by the unmodified reference."""
import numpy as np
import pytest

import golden_cases_r01e as gc


@pytest.mark.parametrize("bname", ["ortho", "tric", "tric60"])
@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_minimize_vectors(oracle, bname, dtname):
    gc.check_minimize_vectors(oracle.minimize_vectors, bname, dtname)


@pytest.mark.parametrize("bname", ["ortho", "tric60"])
@pytest.mark.parametrize("norm", ["rdf", "density", "none"])
def test_interrdf_s(oracle, bname, norm):
    pos, dims, sites = gc.rdfs_inputs(bname)
    gc.check_rdfs(oracle.interrdf_s(pos, sites, dims, nbins=40, rrange=(0.0, 8.0), norm=norm), bname, norm)


@pytest.mark.parametrize("bname", ["none", "ortho", "tric"])
@pytest.mark.parametrize("method,kw", gc.CONTACT_CASES)
def test_contacts(oracle, bname, method, kw):
    pos, dims, ia, ib = gc.contacts_inputs(bname)
    ts = oracle.contacts_timeseries(pos, ia, ib, [(pos[0][ia], pos[0][ib])], dimensions=dims, method=method,
                                    radius=6.0, pbc=dims is not None,
                                    ref_box=None if dims is None else dims[0], kwargs=kw)
    gc.check_contacts(ts, bname, method, kw, rtol=1e-14)
