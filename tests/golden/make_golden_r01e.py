#!/usr/bin/env python
"""Second set of golden fixtures (minimize_vectors, InterRDF_s, Contacts) from the REAL
reference.  TEST INFRASTRUCTURE; same recipe as make_golden.py (reference built out of tree
in /tmp/mda_ref, stub modules for mmtf / fasteners / mda_xdrlib / matplotlib):

    MDA_REF_PATH=/tmp/mda_ref/package:/tmp/mda_ref/stubs python tests/golden/make_golden_r01e.py

Writes tests/golden/reference_vectors_r01e.npz: inputs handed to and outputs returned by the
unmodified reference (MDAnalysis 2.8.0-dev0).  Nothing at test time reads /root/reference.
"""
import os
import sys
import warnings

import numpy as np

for p in os.environ.get("MDA_REF_PATH", "/tmp/mda_ref/package:/tmp/mda_ref/stubs").split(":"):
    sys.path.insert(0, p)
warnings.filterwarnings("ignore")

import MDAnalysis as mda                                   # noqa: E402
import MDAnalysis.lib.distances as rd                      # noqa: E402
from MDAnalysis.lib import mdamath as rmath                # noqa: E402
from MDAnalysis.analysis.rdf import InterRDF_s             # noqa: E402
from MDAnalysis.analysis import contacts as rcontacts      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261018)
BOXES = {
    "ortho": np.array([30, 25, 40, 90, 90, 90], dtype=np.float32),
    "tric": np.array([40, 35, 50, 70, 80, 100], dtype=np.float32),
    "tric60": np.array([30, 30, 30, 60, 60, 90], dtype=np.float32),
}
out = {}

# ---- minimize_vectors: both dtypes, vectors up to several boxes long ---------------------------
for bname, box in BOXES.items():
    for dt in (np.float32, np.float64):
        v = rng.normal(0, 45, (400, 3)).astype(dt)
        v[:4] = [[0, 0, 0], [15, 12.5, 20], [-15, -12.5, -20], [30, 25, 40]]
        k = f"minvec/{bname}/{np.dtype(dt).name}/"
        out[k + "v"] = v
        out[k + "out"] = rd.minimize_vectors(v, box)


def universe(xyz, dims):
    uu = mda.Universe.empty(xyz.shape[1], trajectory=True)
    uu.load_new(xyz.copy(), order="fac", dimensions=dims)
    return uu


# ---- InterRDF_s --------------------------------------------------------------------------------
for bname in ("ortho", "tric60"):
    box = BOXES[bname]
    tv = rmath.triclinic_vectors(box).astype(np.float64)
    F, N = 4, 300
    pos = np.stack([(rng.random((N, 3)) * 1.3 - 0.15) @ tv for _ in range(F)]).astype(np.float32)
    dims = np.tile(box, (F, 1))
    dims[2, :3] *= 1.05
    uu = universe(pos, dims)
    sites = [[np.array([0, 5]), np.arange(100, 160)], [np.array([7, 8, 9]), np.array([7, 8, 9, 200])]]
    ags = [[uu.atoms[a], uu.atoms[b]] for a, b in sites]
    k = f"rdfs/{bname}/"
    out[k + "pos"], out[k + "dims"] = pos, dims
    for s, (a, b) in enumerate(sites):
        out[k + f"sites{s}a"], out[k + f"sites{s}b"] = a, b
    for norm in ("rdf", "density", "none"):
        r = InterRDF_s(uu, ags, nbins=40, range=(0.0, 8.0), norm=norm).run()
        cdf = r.get_cdf()
        for s in range(len(sites)):
            out[k + f"{norm}/count{s}"] = r.results.count[s]
            out[k + f"{norm}/rdf{s}"] = r.results.rdf[s]
            out[k + f"{norm}/cdf{s}"] = cdf[s]
        out[k + f"{norm}/edges"] = r.results.edges

# ---- Contacts ----------------------------------------------------------------------------------
for bname in ("none", "ortho", "tric"):
    box = BOXES.get(bname)
    F, N = 6, 160
    base = rng.normal(0, 6, (N, 3)) + 15
    pos = np.stack([base + rng.normal(0, 0.8, (N, 3)) for _ in range(F)]).astype(np.float32)
    dims = None if box is None else np.tile(box, (F, 1))
    uu = mda.Universe.empty(N, trajectory=True)
    uu.load_new(pos.copy(), order="fac", **({} if dims is None else {"dimensions": dims}))
    gA, gB = uu.atoms[:70], uu.atoms[90:150]
    k = f"contacts/{bname}/"
    out[k + "pos"] = pos
    if dims is not None:
        out[k + "dims"] = dims
    out[k + "idxA"], out[k + "idxB"] = np.arange(70), np.arange(90, 150)
    uu.trajectory[0]
    for method, kw in (("hard_cut", None), ("soft_cut", None), ("radius_cut", None),
                       ("soft_cut", {"beta": 3.0, "lambda_constant": 1.5})):
        c = rcontacts.Contacts(uu, select=(gA, gB), refgroup=(gA, gB), method=method, radius=6.0,
                               pbc=box is not None, kwargs=kw).run()
        tag = method + ("_kw" if kw else "")
        out[k + tag] = c.results.timeseries.copy()
        uu.trajectory[0]

np.savez_compressed(os.path.join(HERE, "reference_vectors_r01e.npz"), **out)
print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_vectors_r01e.npz")) // 1024, "kB")
