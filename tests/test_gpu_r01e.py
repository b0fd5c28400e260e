"""-m gpu: minimize_vectors, InterRDF_s and Contacts on the CUDA path (through the C ABI)
against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

import golden_cases_r01e as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bname", ["ortho", "tric", "tric60"])
@pytest.mark.parametrize("dtname", ["float32", "float64"])
@pytest.mark.parametrize("where", ["host", "device"])
def test_minimize_vectors_golden(bname, dtname, where):
    import torch
    from mdanalysis_b200.lib.distances import minimize_vectors

    def fn(v, box):
        if where == "host":
            return minimize_vectors(v, box)
        return minimize_vectors(torch.as_tensor(v).cuda(), box).cpu().numpy()
    gc.check_minimize_vectors(fn, bname, dtname)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_minimize_vectors_oracle_large(oracle, dt):
    from mdanalysis_b200.lib.distances import minimize_vectors
    rng = np.random.default_rng(11)
    for box in ([31.5, 27.25, 44.0, 90, 90, 90], [40, 35, 50, 70, 80, 100], [33, 33, 33, 60, 60, 90]):
        box = np.array(box, dtype=np.float32)
        v = rng.normal(0, 80, (200_003, 3)).astype(dt)
        assert np.array_equal(minimize_vectors(v, box), oracle.minimize_vectors(v, box))
    with pytest.raises(ValueError):
        minimize_vectors(v[:, :2], box)
    assert minimize_vectors(v[:0], box).shape == (0, 3)


@pytest.mark.parametrize("bname", ["ortho", "tric60"])
@pytest.mark.parametrize("norm", ["rdf", "density", "none"])
@pytest.mark.parametrize("where", ["host", "device"])
def test_interrdf_s_golden(bname, norm, where):
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rdf import InterRDF_s
    pos, dims, sites = gc.rdfs_inputs(bname)
    blk = FrameBlock(pos, dims)
    if where == "device":
        blk = blk.to_device()
    r = InterRDF_s(blk, sites, nbins=40, range=(0.0, 8.0), norm=norm).run()
    gc.check_rdfs({"count": r.results.count, "rdf": r.results.rdf, "cdf": r.get_cdf(), "edges": r.results.edges},
                  bname, norm)
    assert [list(map(len, ix)) for ix in r.results.indices] == [[2, 60], [3, 4]]


def test_interrdf_s_oracle_grid_path(oracle):
    """Groups large enough for the reference to pick nsgrid (n * m >= 1e8 is not needed: the
    method follows lib.distances._determine_method; here the brute-force and the grid path
    are both exercised by two pairs of groups of very different size)."""
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rdf import InterRDF_s
    rng = np.random.default_rng(3)
    box = np.array([42, 40, 44, 90, 90, 90], dtype=np.float32)
    pos = (rng.random((2, 6000, 3)) * box[:3] * 1.2).astype(np.float32)
    sites = [[np.arange(4), np.arange(5000)], [np.array([10, 11]), np.array([10, 11, 12])]]
    want = oracle.interrdf_s(pos, sites, box, nbins=30, rrange=(0.0, 6.0))
    got = InterRDF_s(FrameBlock(pos, box), sites, nbins=30, range=(0.0, 6.0)).run()
    for s in range(2):
        assert np.array_equal(got.results.count[s], want["count"][s])
        assert np.allclose(got.results.rdf[s], want["rdf"][s], rtol=1e-12)
    with pytest.raises(ValueError, match="invalid norm"):
        InterRDF_s(FrameBlock(pos, box), sites, norm="foo")


@pytest.mark.parametrize("bname", ["none", "ortho", "tric"])
@pytest.mark.parametrize("method,kw", gc.CONTACT_CASES)
@pytest.mark.parametrize("where", ["host", "device"])
def test_contacts_golden(bname, method, kw, where):
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.contacts import Contacts
    pos, dims, ia, ib = gc.contacts_inputs(bname)
    blk = FrameBlock(pos, dims)
    if where == "device":
        blk = blk.to_device()
    refA = FrameBlock(pos[:1, ia], None if dims is None else dims[:1])
    refB = FrameBlock(pos[:1, ib], None if dims is None else dims[:1])
    c = Contacts(blk, select=(ia, ib), refgroup=(refA, refB), method=method, radius=6.0,
                 pbc=dims is not None, kwargs=kw).run()
    gc.check_contacts(c.results.timeseries, bname, method, kw, rtol=1e-12)


def test_contacts_callable_and_per_frame_boxes(oracle):
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.contacts import Contacts, pair_list_distances
    rng = np.random.default_rng(8)
    F, N = 5, 400
    dims = np.tile(np.array([28, 30, 26, 80, 95, 70], dtype=np.float32), (F, 1))
    dims[:, :3] *= (1 + 0.03 * np.arange(F))[:, None]
    pos = (rng.random((F, N, 3)) * 60 - 15).astype(np.float32)
    ia, ib = np.arange(0, 150), np.arange(200, 390)
    ii, jj = rng.integers(0, 150, 3000), rng.integers(0, 190, 3000)
    r = pair_list_distances(pos[:, ia], pos[:, ib], ii, jj, dims)
    for f in range(F):
        d = oracle.distance_array(np.ascontiguousarray(pos[f, ia]), np.ascontiguousarray(pos[f, ib]), box=dims[f])
        assert np.array_equal(r[f], d[ii, jj])              # distance_array's arithmetic, bit for bit

    def q(r, r0, scale=1.0):
        return float(np.mean(r < scale * r0))
    c = Contacts(FrameBlock(pos, dims), select=(ia, ib),
                 refgroup=[(FrameBlock(pos[:1, ia], dims[:1]), FrameBlock(pos[:1, ib], dims[:1])),
                           (FrameBlock(pos[1:2, ia], dims[1:2]), FrameBlock(pos[1:2, ib], dims[1:2]))],
                 method=q, radius=7.0, kwargs={"scale": 1.1}).run()
    want0 = oracle.contacts_timeseries(pos, ia, ib, [(pos[0, ia], pos[0, ib])], dimensions=dims, method=q,
                                       radius=7.0, ref_box=dims[0], kwargs={"scale": 1.1})
    want1 = oracle.contacts_timeseries(pos, ia, ib, [(pos[1, ia], pos[1, ib])], dimensions=dims, method=q,
                                       radius=7.0, ref_box=dims[1], kwargs={"scale": 1.1})
    assert np.array_equal(c.results.timeseries[:, 1], want0[:, 1])
    assert np.array_equal(c.results.timeseries[:, 2], want1[:, 1])
    with pytest.raises(ValueError, match="callable"):
        Contacts(FrameBlock(pos, dims), select=(ia, ib), refgroup=(pos[0, ia], pos[0, ib]), method=3)
