#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ from the REAL reference.

TEST INFRASTRUCTURE.  Runs only in the build container, where the reference has
been built out of tree (recipe in DESIGN.md, "Oracle"):

    cp -r /root/reference/package /tmp/mda_ref/ && cd /tmp/mda_ref/package
    CC=/usr/bin/gcc CXX=/usr/bin/g++ python setup.py build_ext --inplace   (+ copy build/lib.*/**/*.so in place)
    stubs for mmtf / fasteners / mda_xdrlib / matplotlib in /tmp/mda_ref/stubs

    MDA_REF_PATH=/tmp/mda_ref/package:/tmp/mda_ref/stubs:/root/reference/testsuite \
        python tests/golden/make_golden.py

Every array written here is an input handed to, or an output returned by, the
unmodified reference (MDAnalysis 2.8.0-dev0).  The tests compare the oracle
(-m "not gpu") and the CUDA path (-m gpu) against these files; nothing at test
time reads /root/reference.
"""
import json
import os
import sys
import warnings

import numpy as np

for p in os.environ.get("MDA_REF_PATH", "/tmp/mda_ref/package:/tmp/mda_ref/stubs:/root/reference/testsuite").split(":"):
    sys.path.insert(0, p)
warnings.filterwarnings("ignore")

import MDAnalysis as mda                                   # noqa: E402
import MDAnalysis.lib.distances as rd                      # noqa: E402
from MDAnalysis.lib import nsgrid as rns                   # noqa: E402
from MDAnalysis.lib import qcprot as rq                    # noqa: E402
from MDAnalysis.lib import mdamath as rmath                # noqa: E402
from MDAnalysis.analysis import rms as rrms, align as ralign  # noqa: E402
from MDAnalysis.analysis.rdf import InterRDF               # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261017)

BOXES = {
    "none": None,
    "ortho": np.array([30, 25, 40, 90, 90, 90], dtype=np.float32),
    "tric": np.array([40, 35, 50, 70, 80, 100], dtype=np.float32),
    "tric60": np.array([30, 30, 30, 60, 60, 90], dtype=np.float32),
}


def cloud(n, scale, spread=1.0, offset=0.0):
    return ((rng.random((n, 3)) - offset) * scale * spread).astype(np.float32)


def canon(pairs, d, self_pairs=False):
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if self_pairs:
        p = np.sort(p, axis=1)
    o = np.lexsort((p[:, 1], p[:, 0]))
    return p[o], np.asarray(d)[o]


out = {}
meta = {"reference": f"MDAnalysis {mda.__version__}", "numpy": np.__version__, "kat": {}}

# ---- lib.distances: every kernel x every box flavour, inside and far outside the cell ----
for bname, box in BOXES.items():
    for tag, spread, off in (("in", 1.0, 0.0), ("far", 7.0, 0.5)):
        a, b, c, d = cloud(61, 40, spread, off), cloud(47, 40, spread, off), cloud(61, 40, spread, off), cloud(61, 40, spread, off)
        k = f"dist/{bname}/{tag}/"
        out[k + "a"], out[k + "b"], out[k + "c"], out[k + "d"] = a, b, c, d
        out[k + "distance_array"] = rd.distance_array(a, b, box=box)
        out[k + "self_distance_array"] = rd.self_distance_array(a, box=box)
        out[k + "calc_bonds"] = rd.calc_bonds(a, c, box=box)
        out[k + "calc_angles"] = rd.calc_angles(a, c, d, box=box)
        out[k + "calc_dihedrals"] = rd.calc_dihedrals(a, c, d, a[::-1].copy(), box=box)
        if box is not None:
            out[k + "apply_PBC"] = rd.apply_PBC(a, box)
            out[k + "transform_RtoS"] = rd.transform_RtoS(a, box)
            out[k + "transform_StoR"] = rd.transform_StoR(a, box)

# points on faces / tiny negatives / huge values through the wraps
edge = np.array([[0, 0, 0], [30, 25, 40], [-1e-7, 25, -0.0], [29.999998, 1e-9, 40.000004],
                 [-30, -25, -40], [60, 50, 80], [1e9, -1e9, 5]], dtype=np.float32)
out["edge/xyz"] = edge
for bname in ("ortho", "tric", "tric60"):
    out[f"edge/{bname}/apply_PBC"] = rd.apply_PBC(edge, BOXES[bname])

# ---- adk (config C1): frames 0 and 10 of the reference's own test trajectory ---------------
from MDAnalysisTests.datafiles import PSF, DCD             # noqa: E402
u = mda.Universe(PSF, DCD)
u.trajectory[0]
x0 = u.atoms.positions.copy()
u.trajectory[10]
x1 = u.atoms.positions.copy()
out["adk/x0"], out["adk/x1"] = x0, x1
samp = rng.integers(0, 3341 * 3341, 4096)
out["adk/sample_index"] = samp
for bname, box in (("none", None), ("ortho", np.array([50, 50, 50, 90, 90, 90], dtype=np.float32)),
                   ("tric", np.array([50, 50, 50, 70.53571, 109.48542, 70.518196], dtype=np.float32))):
    d = rd.distance_array(x0, x1, box=box)
    s = rd.self_distance_array(x1, box=box)
    out[f"adk/{bname}/distance_array_sample"] = d.ravel()[samp]
    out[f"adk/{bname}/distance_array_rowsum"] = d.sum(axis=1)
    out[f"adk/{bname}/self_distance_array_sample"] = s[samp % len(s)]
    out[f"adk/{bname}/minmax"] = np.array([d.min(), d.max(), s.min(), s.max()])
    if box is not None:
        out[f"adk/{bname}/box"] = box
# constants asserted by the reference's own tests (testsuite/MDAnalysisTests/...)
meta["kat"]["adk_distance_array_min_max (lib/test_distances.py:451-454)"] = [0.11981228170520701, 53.572192429459619]
meta["kat"]["qcp_dummy (lib/test_qcprot.py:44)"] = 0.7174389197644676
meta["kat"]["qcp_regression (lib/test_qcprot.py:59-67)"] = 0.6057544485785074
meta["kat"]["rmsd_adk_first_last (analysis/test_rms.py:101)"] = 6.820321761927005

# QCP regression inputs (test_qcprot.py:47-67): protein of adk, frames 0 and 1, mass weights
prot = u.select_atoms("protein")
w = prot.masses.astype(np.float64)
w /= np.mean(w)
u.trajectory[0]
p1 = prot.positions.astype(np.float64)
u.trajectory[1]
p2 = prot.positions.astype(np.float64)
rot = np.zeros(9)
out["qcp/reg/p1"], out["qcp/reg/p2"], out["qcp/reg/w"] = p1, p2, w
out["qcp/reg/rmsd"] = np.array(rq.CalcRMSDRotationalMatrix(p1, p2, len(prot), rot, w))
out["qcp/reg/rot"] = rot.copy()
# rms.rmsd first vs last frame of the backbone, superposition (test_rms.py:91-101)
ca = u.select_atoms("backbone")
u.trajectory[0]
A = ca.positions.copy()
u.trajectory[-1]
B = ca.positions.copy()
out["rmsd/ca/A"], out["rmsd/ca/B"] = A, B
out["rmsd/ca/superposition"] = np.array(rrms.rmsd(A, B, superposition=True))
out["rmsd/ca/centered"] = np.array(rrms.rmsd(A, B, center=True))
out["rmsd/ca/plain"] = np.array(rrms.rmsd(A, B))

# ---- QCP on synthetic structures, float64, with and without weights ----------------------------
for n in (3, 50, 1000):
    x = rng.normal(0, 10, (n, 3))
    y = x @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 1, (n, 3))
    x -= x.mean(0)
    y -= y.mean(0)
    w = rng.uniform(1, 16, n)
    out[f"qcp/{n}/x"], out[f"qcp/{n}/y"], out[f"qcp/{n}/w"] = x, y, w
    for wname, ww in (("now", None), ("w", w)):
        r = np.zeros(9)
        out[f"qcp/{n}/{wname}/rmsd"] = np.array(rq.CalcRMSDRotationalMatrix(x, y, n, r, ww))
        out[f"qcp/{n}/{wname}/rot"] = r.copy()

# ---- FastNS + capped_distance ------------------------------------------------------------------
for bname, cutoff, n in (("ortho", 4.0, 1200), ("ortho", 7.3, 800), ("tric", 6.0, 1200), ("tric60", 5.0, 1000)):
    box = BOXES[bname]
    tv = rmath.triclinic_vectors(box).astype(np.float64)
    xyz = ((rng.random((n, 3)) * 3 - 1) @ tv).astype(np.float32)
    q = ((rng.random((200, 3)) * 3 - 1) @ tv).astype(np.float32)
    k = f"ns/{bname}_{cutoff}/"
    out[k + "xyz"], out[k + "q"] = xyz, q
    g = rns.FastNS(cutoff, xyz, box)
    r = g.self_search()
    out[k + "self_pairs"], out[k + "self_dist"] = canon(r.get_pairs(), r.get_pair_distances(), True)
    r = g.search(q)
    out[k + "search_pairs"], out[k + "search_dist"] = canon(r.get_pairs(), r.get_pair_distances())
g = rns.FastNS(5.0, out["ns/ortho_4.0/xyz"], BOXES["ortho"], pbc=False)
r = g.self_search()
out["ns/nopbc/self_pairs"], out["ns/nopbc/self_dist"] = canon(r.get_pairs(), r.get_pair_distances(), True)

for bname, box in BOXES.items():
    a, b = cloud(300, 30), cloud(200, 30)
    k = f"capped/{bname}/"
    out[k + "a"], out[k + "b"] = a, b
    for method in ("bruteforce", "nsgrid", "pkdtree"):
        for mn in (None, 1.5):
            kk = f"{k}{method}/{'min' if mn else 'nomin'}/"
            out[kk + "pairs"], out[kk + "dist"] = canon(*rd.capped_distance(a, b, 5.0, min_cutoff=mn, box=box, method=method))
            out[kk + "self_pairs"], out[kk + "self_dist"] = canon(
                *rd.self_capped_distance(a, 5.0, min_cutoff=mn, box=box, method=method), self_pairs=True)

# ---- analysis drivers through Universes --------------------------------------------------------
F, N = 5, 250
ref_xyz = rng.normal(0, 8, (N, 3))
traj = np.stack([(ref_xyz @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 3, 3)
                  + rng.normal(0, 0.5, (N, 3))) for _ in range(F)]).astype(np.float32)
masses = rng.uniform(1, 32, N)
out["traj/xyz"], out["traj/ref"], out["traj/masses"] = traj, ref_xyz.astype(np.float32), masses


def universe(xyz, dims=None):
    uu = mda.Universe.empty(xyz.shape[1], trajectory=True)
    uu.add_TopologyAttr("masses", masses[:xyz.shape[1]])
    uu.load_new(xyz.copy(), order="fac", **({} if dims is None else {"dimensions": dims}))
    return uu


r = universe(ref_xyz.astype(np.float32)[None])
for wname in ("none", "mass"):
    ww = None if wname == "none" else "mass"
    R = rrms.RMSD(universe(traj), r, select="all", weights=ww).run()
    out[f"traj/{wname}/rmsd"] = R.results.rmsd.copy()
    u2 = universe(traj)
    Al = ralign.AlignTraj(u2, r, select="all", weights=ww, in_memory=True).run()
    out[f"traj/{wname}/align_rmsd"] = np.asarray(Al.results.rmsd).copy()
    out[f"traj/{wname}/aligned"] = u2.trajectory.timeseries(order="fac").copy()
# fit on a sub-selection, move all atoms
sel = "index 10:120"
u2 = universe(traj)
Al = ralign.AlignTraj(u2, r, select=sel, in_memory=True).run()
out["traj/sel/index"] = np.arange(10, 121)
out["traj/sel/align_rmsd"] = np.asarray(Al.results.rmsd).copy()
out["traj/sel/aligned"] = u2.trajectory.timeseries(order="fac").copy()

for bname in ("ortho", "tric60"):
    box = BOXES[bname]
    tv = rmath.triclinic_vectors(box).astype(np.float64)
    F, N = 3, 500
    pos = np.stack([(rng.random((N, 3)) * 1.4 - 0.2) @ tv for _ in range(F)]).astype(np.float32)
    out[f"rdf/{bname}/pos"] = pos
    uu = mda.Universe.empty(N, trajectory=True)
    uu.load_new(pos.copy(), order="fac", dimensions=np.tile(box, (F, 1)))
    for tag, ga, gb, kw in (("ab", uu.atoms[:300], uu.atoms[200:], {}),
                            ("self", uu.atoms, uu.atoms, {}),
                            ("excl", uu.atoms, uu.atoms, {"exclusion_block": (5, 5)})):
        for rmax, nb in ((6.0, 75), (12.0, 150)):
            try:
                res = InterRDF(ga, gb, nbins=nb, range=(0.0, rmax), **kw).run()
            except ValueError:
                continue                     # cutoff too large for this box (FastNS refuses)
            k = f"rdf/{bname}/{tag}/{rmax}_{nb}/"
            out[k + "count"], out[k + "edges"], out[k + "rdf"] = res.results.count, res.results.edges, res.results.rdf

np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
with open(os.path.join(HERE, "reference_vectors.json"), "w") as fh:
    json.dump(meta, fh, indent=1)
print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_vectors.npz")) // 1024, "kB")
