"""InterRDF: the fine-cell traversal (csrc/rowfine.cuh: planner + walker kernels) against the
per-pair row kernel (csrc/rowtile.cuh, itself pinned to the oracle in test_gpu_analysis.py) and,
where the oracle finishes in seconds, against the oracle (FastNS + np.histogram restatement of
analysis/rdf.py:253-316).  Bit-exact counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _counts(blk, b2=None, env=None, monkeypatch=None, **kw):
    from mdanalysis_b200.analysis.rdf import InterRDF
    for k in ("MDAGPU_RDF_NO_FINE", "MDAGPU_RDF_NO_QUAD"):
        monkeypatch.delenv(k, raising=False)
    if env:
        monkeypatch.setenv(env, "1")
    out = InterRDF(blk, b2, **kw).run().results.count
    if env:
        monkeypatch.delenv(env)
    return out


CASES = ["c2_like", "unwrapped", "two_groups", "dense_fallback_groups", "dense_fallback_rows", "lattice_on_edges",
         "blob", "upper_face", "many_bins", "per_frame_boxes", "segments"]


@pytest.mark.parametrize("case", CASES)
def test_fine_traversal_counts(oracle, case, monkeypatch):
    import torch
    from mdanalysis_b200 import FrameBlock
    import zlib
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    box = np.array([100, 100, 100, 90, 90, 90], dtype=np.float32)
    pos2 = None
    check_oracle = False
    kw = dict(nbins=150, range=(0.0, 15.0))
    if case == "c2_like":                       # the headline geometry: a whole row of home cells per CTA
        pos = (rng.random((3, 33334, 3)) * 100).astype(np.float32)
    elif case == "unwrapped":                   # coordinates outside the cell, unequal edges
        box = np.array([47, 52, 61, 90, 90, 90], dtype=np.float32)
        pos = (rng.random((2, 9000, 3)) * box[:3] * 1.3 - 0.15 * box[:3]).astype(np.float32)
        kw = dict(nbins=140, range=(0.0, 7.0))
        check_oracle = True
    elif case == "two_groups":                  # g1 != g2: full shell, 25 rows
        box = np.array([47, 52, 61, 90, 90, 90], dtype=np.float32)
        p = (rng.random((2, 9000, 3)) * box[:3]).astype(np.float32)
        pos, pos2 = np.ascontiguousarray(p[:, :4500]), np.ascontiguousarray(p[:, 3000:])
        kw = dict(nbins=140, range=(0.0, 7.0))
        check_oracle = True
    elif case == "dense_fallback_groups":       # rows do not fit the tile: the planner declines, rowquad.cuh takes over
        box = np.array([30, 31, 29, 90, 90, 90], dtype=np.float32)
        pos = (rng.random((2, 60000, 3)) * box[:3] * 1.1).astype(np.float32)
        kw = dict(nbins=80, range=(0.0, 4.0))
    elif case == "dense_fallback_rows":
        box = np.array([20, 22, 24, 90, 90, 90], dtype=np.float32)
        pos = (rng.random((1, 400000, 3)) * box[:3]).astype(np.float32)
        kw = dict(nbins=60, range=(0.0, 3.0))
    elif case == "lattice_on_edges":            # every distance sits on a bin edge: all pairs go through float64
        n1, a = 24, 0.5
        g = np.arange(n1, dtype=np.float32) * a
        pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(1, -1, 3).astype(np.float32)
        box = np.array([n1 * a] * 3 + [90, 90, 90], dtype=np.float32)
        kw = dict(nbins=16, range=(0.0, 2.0))
        check_oracle = True
    elif case == "blob":                        # everything in a few cells: overflow of queries / tile / list
        pos = rng.normal(50, 3, (2, 20000, 3)).astype(np.float32)
        kw = dict(nbins=100, range=(0.0, 10.0))
    elif case == "upper_face":                  # -1e-6 + L == L in float32: an atom ON the upper face of the cell
        box = np.array([47, 52, 61, 90, 90, 90], dtype=np.float32)
        pos = (rng.random((2, 9000, 3)) * box[:3]).astype(np.float32)
        pos[0, :50, 0] = -1e-6
        pos[1, :50, 1] = -1e-7
        pos[1, 50:80, 2] = -3e-7
        kw = dict(nbins=140, range=(0.0, 7.0))
        check_oracle = True
    elif case == "many_bins":                   # fewer lane columns per bin
        pos = (rng.random((2, 33334, 3)) * 100).astype(np.float32)
        kw = dict(nbins=1000, range=(0.0, 15.0))
    elif case == "per_frame_boxes":
        dims = np.tile(np.array([47, 52, 61, 90, 90, 90], dtype=np.float32), (3, 1))
        dims[1, :3] *= 1.13
        dims[2, :3] *= 0.91
        pos = np.stack([(rng.random((9000, 3)) * dims[f, :3]).astype(np.float32) for f in range(3)])
        box = dims
        kw = dict(nbins=140, range=(0.0, 7.0))
        check_oracle = True
    else:                                       # a long box: several x segments per row
        box = np.array([400, 40, 40, 90, 90, 90], dtype=np.float32)
        pos = (rng.random((2, 40000, 3)) * box[:3]).astype(np.float32)
        kw = dict(nbins=100, range=(0.0, 8.0))
    blk = FrameBlock(torch.as_tensor(pos).cuda(), box)
    b2 = None if pos2 is None else FrameBlock(torch.as_tensor(pos2).cuda(), box)
    got = _counts(blk, b2, monkeypatch=monkeypatch, **kw)
    want = _counts(blk, b2, env="MDAGPU_RDF_NO_QUAD", monkeypatch=monkeypatch, **kw)
    assert got.sum() > 0
    assert np.array_equal(got, want)
    # host-resident frames go through the upload ring in pieces
    got_h = _counts(FrameBlock(pos, box), None if pos2 is None else FrameBlock(pos2, box), monkeypatch=monkeypatch, **kw)
    assert np.array_equal(got_h, want)
    if check_oracle:
        ref = oracle.interrdf(pos, pos if pos2 is None else pos2, box, nbins=kw["nbins"], rrange=kw["range"])
        assert np.array_equal(got, ref["count"])
