"""-m gpu: caller-owned PAGEABLE host arrays larger than the direct-copy threshold go through the
pinned ring of csrc/hostcopy.cu (chunked DMA + host thread pool); the results must be the bytes the
device holds -- compared with the device-resident call and with pinned host arrays."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_distance_array_pageable_result_equals_device_result():
    import torch
    from mdanalysis_b200.lib import distances as gd
    rng = np.random.default_rng(5)
    a = (rng.random((3100, 3)) * 40).astype(np.float32)
    b = (rng.random((2900, 3)) * 40).astype(np.float32)
    box = np.array([40, 41, 42, 90, 90, 90], dtype=np.float32)
    dev = gd.distance_array(torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda(), box=box).cpu().numpy()
    page = np.full((3100, 2900), -1.0)                         # 72 MB: nine chunks of the ring, the last one partial
    gd.distance_array(a, b, box=box, result=page)
    assert np.array_equal(page, dev)
    fresh = gd.distance_array(a, b, box=box)                   # np.empty inside: pages not yet touched
    assert np.array_equal(fresh, dev)
    pin = torch.empty((3100, 2900), dtype=torch.float64).pin_memory().numpy()
    gd.distance_array(a, b, box=box, result=pin)
    assert np.array_equal(pin, dev)
    # odd sizes around the chunk size (8 MiB = 1 048 576 doubles)
    for n, m in ((1024, 1024), (1025, 1023), (1449, 1449)):
        got = gd.distance_array(a[:n], b[:m])
        assert np.array_equal(got, gd.distance_array(torch.as_tensor(a[:n]).cuda(), torch.as_tensor(b[:m]).cuda()).cpu().numpy())


def test_pageable_trajectories_rdf_and_rmsd(oracle):
    import torch
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rdf import InterRDF
    from mdanalysis_b200.analysis.rms import RMSD
    rng = np.random.default_rng(6)
    box = np.array([60, 60, 60, 90, 90, 90], dtype=np.float32)
    pos = (rng.random((260, 9000, 3)) * 60).astype(np.float32)             # 28 MB, several pieces of the upload ring
    want = InterRDF(FrameBlock(torch.as_tensor(pos).cuda(), box), None, nbins=90, range=(0.0, 9.0)).run().results.count
    got = InterRDF(FrameBlock(pos, box), None, nbins=90, range=(0.0, 9.0)).run().results.count
    assert np.array_equal(got, want)
    pin = torch.as_tensor(pos).pin_memory().numpy()
    assert np.array_equal(InterRDF(FrameBlock(pin, box), None, nbins=90, range=(0.0, 9.0)).run().results.count, want)
    ref = pos[0]
    r_dev = RMSD(FrameBlock(torch.as_tensor(pos).cuda()), ref).run().results.rmsd
    r_page = RMSD(FrameBlock(pos), ref).run().results.rmsd
    assert np.array_equal(r_dev, r_page)


def test_fastns_export_to_pageable_arrays(oracle):
    import torch
    from mdanalysis_b200.lib.nsgrid import FastNS
    rng = np.random.default_rng(7)
    box = np.array([90, 90, 90, 90, 90, 90], dtype=np.float32)
    xyz = (rng.random((70000, 3)) * 90).astype(np.float32)
    res = FastNS(9.0, torch.as_tensor(xyz).cuda(), box).self_search()
    p_dev, d_dev = res.to_tensors()
    pairs, dist = res.get_pairs(), res.get_pair_distances()                 # ~10 M pairs: 160 MB + 80 MB to NumPy arrays
    assert len(pairs) > 5_000_000
    assert np.array_equal(pairs, p_dev.cpu().numpy()) and np.array_equal(dist, d_dev.cpu().numpy())
