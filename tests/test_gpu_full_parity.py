"""-m gpu: EXHAUSTIVE parity at BASELINE.json's full sizes for the index-exact rows.

north_star: "bit-exact for capped_distance/FastNS pair index sets and RDF bin counts" on every
config.  The sampled checks of test_gpu_nsgrid.py / test_gpu_analysis.py stay; here the whole
result is compared:

* C3 -- 1 000 000 atoms, 8 A, box [L, L, L, 60, 60, 90] and its orthorhombic twin: the FULL sorted
  set of unordered pairs (min, max) and the float64 distance of every one of the ~107 M pairs
  against the oracle's C port of lib/nsgrid.pyx (about 50 s of CPU per box).
* C2 -- 33 334 O sites, box 100 A, 150 bins to 15 A: ten WHOLE frames, each against the port
  (capped_distance -> np.histogram), their sum against one ten-frame launch.
* MDAGPU_FULL_PARITY=1 (run once per round, log under profiles/): the 1000-frame C2 total against
  the port on all host cores, and C3 against the unmodified reference's FastNS (baseline/_ref,
  ~130 s).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.environ.get("MDAGPU_FULL_PARITY", "") not in ("", "0")


def _c3_system(oracle, angles, nn=1_000_000, seed=20261019):
    vol_factor = 0.70710678 if angles[0] == 60 else 1.0
    L = (nn / 0.1 / vol_factor) ** (1.0 / 3.0)
    box = np.array([L, L, L, *angles], dtype=np.float32)
    rng = np.random.default_rng(seed)
    xyz = (rng.random((nn, 3)) @ oracle.triclinic_vectors(box).astype(np.float64)).astype(np.float32)
    return xyz, box


def _sorted_keys(torch, pairs, d, nn):
    """(key = min * n + max sorted, d in the same order) on the device."""
    p = torch.as_tensor(pairs).cuda() if not torch.is_tensor(pairs) else pairs
    dd = torch.as_tensor(d).cuda() if not torch.is_tensor(d) else d
    lo, hi = torch.minimum(p[:, 0], p[:, 1]), torch.maximum(p[:, 0], p[:, 1])
    key = lo * nn + hi
    del p, lo, hi
    key, order = torch.sort(key)
    return key, dd[order]


def _gpu_self_search(torch, xyz, box, cutoff=8.0):
    from mdanalysis_b200.lib.nsgrid import FastNS
    res = FastNS(cutoff, torch.as_tensor(xyz).cuda(), box).self_search()
    pairs, dist = res.to_tensors()
    del res
    key, d = _sorted_keys(torch, pairs, dist, len(xyz))
    del pairs, dist
    torch.cuda.empty_cache()
    return key, d


@pytest.mark.parametrize("angles", [(60, 60, 90), (90, 90, 90)])
def test_c3_full_pair_set_and_distances_vs_port(oracle, angles):
    import torch
    xyz, box = _c3_system(oracle, angles)
    nn = len(xyz)
    gkey, gd = _gpu_self_search(torch, xyz, box)
    r = oracle.FastNS(8.0, xyz, box).self_search()
    okey, od = _sorted_keys(torch, r.get_pairs(), r.get_pair_distances(), nn)
    del r
    assert gkey.shape == okey.shape and 1.06e8 < gkey.shape[0] < 1.09e8
    assert bool((gkey[1:] > gkey[:-1]).all())              # every unordered pair once
    assert torch.equal(gkey, okey)                          # the same ~107 M pairs
    assert torch.equal(gd, od)                              # bit-equal float64 distances, all of them
    assert float(gd.max()) <= 8.0


def test_c3_capped_distance_ordered_pairs_vs_port(oracle):
    """capped_distance(x, x) of the C3 system restricted to 250 000 query atoms: the ORDERED pair
    set including the self pairs, against the port's search()."""
    import torch
    from mdanalysis_b200.lib.nsgrid import FastNS
    xyz, box = _c3_system(oracle, (60, 60, 90))
    nn, nq = len(xyz), 250_000
    res = FastNS(8.0, torch.as_tensor(xyz).cuda(), box).search(torch.as_tensor(xyz[:nq]).cuda())
    p, d = res.to_tensors()
    del res
    gkey, order = torch.sort(p[:, 0] * nn + p[:, 1])
    gd = d[order]
    del p, d, order
    r = oracle.FastNS(8.0, xyz, box).search(xyz[:nq])
    op = torch.as_tensor(r.get_pairs()).cuda()
    od = torch.as_tensor(r.get_pair_distances()).cuda()
    okey, order = torch.sort(op[:, 0] * nn + op[:, 1])
    od = od[order]
    assert torch.equal(gkey, okey) and torch.equal(gd, od)
    assert int((gd == 0.0).sum()) >= nq                    # the self pairs are there


def _c2_frames(n_frames, seed=20261018):
    rng = np.random.default_rng(seed)
    return (rng.random((n_frames, 33334, 3), dtype=np.float32) * np.float32(100.0))


def _port_counts(oracle, pos, box):
    _, dist = oracle.capped_distance(pos, pos, 15.0, box=box, method="nsgrid")
    return np.histogram(dist, bins=150, range=(0.0, 15.0))[0]


def test_c2_ten_whole_frames_vs_port(oracle):
    import torch
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rdf import InterRDF
    box = np.array([100, 100, 100, 90, 90, 90], dtype=np.float32)
    pos = _c2_frames(10)
    dev = torch.as_tensor(pos).cuda()
    total = np.zeros(150, dtype=np.int64)
    for f in range(10):
        want = _port_counts(oracle, pos[f], box)
        got = InterRDF(FrameBlock(dev[f:f + 1], box), None, nbins=150, range=(0.0, 15.0)).run().results.count
        assert np.array_equal(got, want.astype(np.float64)), f"frame {f}"
        total += want
    both = InterRDF(FrameBlock(dev, box), None, nbins=150, range=(0.0, 15.0)).run().results.count
    assert np.array_equal(both, total.astype(np.float64))
    host = InterRDF(FrameBlock(pos, box), None, nbins=150, range=(0.0, 15.0)).run().results.count
    assert np.array_equal(host, total.astype(np.float64))
    assert total[0] >= 10 * 33334                          # the self pairs sit in bin 0 (rdf.py: no exclusion)


def _port_block(args):
    lo, hi = args
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    box = np.array([100, 100, 100, 90, 90, 90], dtype=np.float32)
    total = np.zeros(150, dtype=np.int64)
    pos = _c2_frames(1000)[lo:hi]
    for f in range(hi - lo):
        total += _port_counts(orc, pos[f], box)
    return total


@pytest.mark.skipif(not FULL, reason="MDAGPU_FULL_PARITY=1: 1000 whole frames on all host cores (~2 min on 16 cores)")
def test_c2_thousand_frame_total_vs_port(oracle):
    import multiprocessing as mp
    import torch
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rdf import InterRDF
    ncpu = max(1, len(os.sched_getaffinity(0)))
    cuts = np.linspace(0, 1000, 4 * ncpu + 1).astype(int)
    with mp.get_context("spawn").Pool(ncpu) as pool:
        parts = pool.map(_port_block, [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a])
    want = np.sum(parts, axis=0)
    box = np.array([100, 100, 100, 90, 90, 90], dtype=np.float32)
    got = InterRDF(FrameBlock(torch.as_tensor(_c2_frames(1000)).cuda(), box), None, nbins=150,
                   range=(0.0, 15.0)).run().results.count
    assert np.array_equal(got, want.astype(np.float64))
    print("C2 1000-frame total: %d pairs, counts identical to the port" % int(want.sum()))


@pytest.mark.skipif(not FULL, reason="MDAGPU_FULL_PARITY=1: the unmodified reference's FastNS on 1 M atoms (~130 s)")
def test_c3_full_pair_set_vs_real_reference(oracle):
    import torch
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "MDAnalysis")):
        pytest.skip("baseline/_ref (the built reference) is not present")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import warnings
    warnings.filterwarnings("ignore")
    from MDAnalysis.lib.nsgrid import FastNS as RefNS
    xyz, box = _c3_system(oracle, (60, 60, 90))
    gkey, gd = _gpu_self_search(torch, xyz, box)
    r = RefNS(8.0, xyz, box).self_search()
    okey, od = _sorted_keys(torch, r.get_pairs(), r.get_pair_distances(), len(xyz))
    assert torch.equal(gkey, okey) and torch.equal(gd, od)
    print("C3: %d pairs, pair set and distances identical to MDAnalysis FastNS" % gkey.shape[0])
