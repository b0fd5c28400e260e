#!/usr/bin/env python
"""Pin the oracle (oracle.c / oracle.py) against the real reference.

TEST INFRASTRUCTURE.  Runs only in the build container, where the reference
has been built out-of-tree (see DESIGN.md "Oracle"):

    cp -r /root/reference/package /tmp/mda_ref/ && cd /tmp/mda_ref/package
    CC=/usr/bin/gcc CXX=/usr/bin/g++ python setup.py build_ext --inplace
    (+ stub modules mmtf, fasteners, mda_xdrlib, matplotlib in /tmp/mda_ref/stubs)

    MDA_REF_PATH=/tmp/mda_ref/package:/tmp/mda_ref/stubs python oracle/check_against_reference.py

Every comparison is bit-for-bit unless a tolerance is printed.
"""
import os
import sys

import numpy as np

for p in os.environ.get("MDA_REF_PATH", "/tmp/mda_ref/package:/tmp/mda_ref/stubs").split(":"):
    sys.path.insert(0, p)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import warnings
warnings.filterwarnings("ignore")
import MDAnalysis.lib.distances as rd          # noqa: E402
from MDAnalysis.lib import nsgrid as rns       # noqa: E402
from MDAnalysis.lib import qcprot as rq        # noqa: E402
from MDAnalysis.analysis import rms as rrms, align as ralign  # noqa: E402

from oracle import oracle as orc               # noqa: E402

FAIL = 0


def check(name, ok, extra=""):
    global FAIL
    print(("ok   " if ok else "FAIL ") + name + (" " + extra if extra else ""))
    if not ok:
        FAIL += 1


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def canon(pairs, d, self_pairs=False):
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if self_pairs:
        p = np.sort(p, axis=1)
    o = np.lexsort((p[:, 1], p[:, 0]))
    return p[o], np.asarray(d)[o]


rng = np.random.default_rng(1234)
BOXES = {
    "none": None,
    "ortho": np.array([30, 25, 40, 90, 90, 90], dtype=np.float32),
    "tric": np.array([40, 35, 50, 70, 80, 100], dtype=np.float32),
    "tric2": np.array([30, 30, 30, 60, 60, 90], dtype=np.float32),
}


def cloud(n, scale, spread=1.0, offset=0.0):
    return ((rng.random((n, 3)) - offset) * scale * spread).astype(np.float32)


# ---- wraps, distance arrays, bonds/angles/dihedrals -------------------------
for impl in ("port", "reference"):
    for bname, box in BOXES.items():
        for spread, off in ((1.0, 0.0), (3.0, 0.5), (9.0, 0.5)):
            a = cloud(211, 40, spread, off)
            b = cloud(173, 40, spread, off)
            c = cloud(211, 40, spread, off)
            d = cloud(211, 40, spread, off)
            tag = f"[{impl}] {bname} spread={spread}"
            check("distance_array " + tag, same(orc.distance_array(a, b, box, impl=impl), rd.distance_array(a, b, box=box)))
            check("self_distance_array " + tag, same(orc.self_distance_array(a, box, impl=impl), rd.self_distance_array(a, box=box)))
            check("calc_bonds " + tag, same(orc.calc_bonds(a, c, box, impl=impl), rd.calc_bonds(a, c, box=box)))
            check("calc_angles " + tag, same(orc.calc_angles(a, c, d, box, impl=impl), rd.calc_angles(a, c, d, box=box)))
            check("calc_dihedrals " + tag, same(orc.calc_dihedrals(a, c, d, a[::-1], box, impl=impl), rd.calc_dihedrals(a, c, d, a[::-1], box=box)))
            if box is not None:
                check("apply_PBC " + tag, same(orc.apply_PBC(a, box, impl=impl), rd.apply_PBC(a, box)))
                check("transform_RtoS " + tag, same(orc.transform_RtoS(a, box, impl=impl), rd.transform_RtoS(a, box)))
                check("transform_StoR " + tag, same(orc.transform_StoR(a, box, impl=impl), rd.transform_StoR(a, box)))

# points exactly on faces / tiny negatives
edge = np.array([[0, 0, 0], [30, 25, 40], [-1e-7, 25, -0.0], [29.999998, 1e-9, 40.000004],
                 [-30, -25, -40], [60, 50, 80], [1e9, -1e9, 5]], dtype=np.float32)
for bname in ("ortho", "tric", "tric2"):
    check("apply_PBC edge " + bname, same(orc.apply_PBC(edge, BOXES[bname]), rd.apply_PBC(edge, BOXES[bname])))

# ---- FastNS -------------------------------------------------------------------
for bname, box, cutoff, n in (("ortho", BOXES["ortho"], 4.0, 3000), ("ortho", BOXES["ortho"], 7.3, 1500),
                              ("tric", BOXES["tric"], 6.0, 3000), ("tric2", BOXES["tric2"], 5.0, 2500),
                              ("tric2", BOXES["tric2"], 2.0, 2500)):
    tv = orc.triclinic_vectors(box).astype(np.float64)
    xyz = ((rng.random((n, 3)) * 3 - 1) @ tv).astype(np.float32)
    q = ((rng.random((500, 3)) * 3 - 1) @ tv).astype(np.float32)
    g_ref = rns.FastNS(cutoff, xyz, box)
    g_orc = orc.FastNS(cutoff, xyz, box)
    r1, r2 = g_ref.self_search(), g_orc.self_search()
    p1, d1 = canon(r1.get_pairs(), r1.get_pair_distances(), True)
    p2, d2 = canon(r2.get_pairs(), r2.get_pair_distances(), True)
    check(f"FastNS.self_search {bname} rc={cutoff}", same(p1, p2) and same(d1, d2), f"{len(p1)} pairs")
    r1, r2 = g_ref.search(q), g_orc.search(q)
    p1, d1 = canon(r1.get_pairs(), r1.get_pair_distances())
    p2, d2 = canon(r2.get_pairs(), r2.get_pair_distances())
    check(f"FastNS.search {bname} rc={cutoff}", same(p1, p2) and same(d1, d2), f"{len(p1)} pairs")

# ---- capped_distance front ends -----------------------------------------------
for bname, box in BOXES.items():
    a = cloud(400, 30)
    b = cloud(300, 30)
    for method in (None, "bruteforce", "nsgrid"):
        for mn in (None, 1.5):
            p1, d1 = canon(*rd.capped_distance(a, b, 5.0, min_cutoff=mn, box=box, method=method))
            p2, d2 = canon(*orc.capped_distance(a, b, 5.0, min_cutoff=mn, box=box, method=method))
            check(f"capped_distance {bname} {method} min={mn}", same(p1, p2) and same(d1, d2), f"{len(p1)}")
            p1, d1 = canon(*rd.self_capped_distance(a, 5.0, min_cutoff=mn, box=box, method=method), self_pairs=True)
            p2, d2 = canon(*orc.self_capped_distance(a, 5.0, min_cutoff=mn, box=box, method=method), self_pairs=True)
            check(f"self_capped_distance {bname} {method} min={mn}", same(p1, p2) and same(d1, d2), f"{len(p1)}")

# ---- QCP -------------------------------------------------------------------
for dt in (np.float64, np.float32):
    for n in (3, 50, 1000):
        x = rng.normal(0, 10, (n, 3)).astype(dt)
        y = (x + rng.normal(0, 1, (n, 3))).astype(dt)
        x -= x.mean(0)
        y -= y.mean(0)
        w = rng.uniform(1, 16, n).astype(dt)
        for ww in (None, w):
            r1 = np.zeros(9, dtype=dt)
            r2 = np.zeros(9, dtype=dt)
            v1 = rq.CalcRMSDRotationalMatrix(x, y, n, r1, ww)
            v2 = orc.CalcRMSDRotationalMatrix(x, y, n, r2, ww)
            check(f"qcp {dt.__name__} n={n} w={'y' if ww is not None else 'n'}", v1 == v2 and same(r1, r2), f"rmsd={v1!r}")
            v1 = rq.CalcRMSDRotationalMatrix(x, y, n, None, ww)
            v2 = orc.CalcRMSDRotationalMatrix(x, y, n, None, ww)
            check("qcp norot", v1 == v2)
a = rng.normal(0, 10, (100, 3))
b = rng.normal(0, 10, (100, 3))
w = rng.uniform(1, 3, 100)
check("rmsd superposition", rrms.rmsd(a, b, weights=w, superposition=True) == orc.rmsd(a, b, weights=w, superposition=True))
check("rmsd center", rrms.rmsd(a, b, center=True) == orc.rmsd(a, b, center=True))
R1, v1 = ralign.rotation_matrix(a, b, weights=w)
R2, v2 = orc.rotation_matrix(a, b, weights=w)
check("rotation_matrix", same(R1, R2) and v1 == v2)

# ---- analysis drivers through a Universe --------------------------------------
import MDAnalysis as mda                                    # noqa: E402
from MDAnalysis.analysis.rdf import InterRDF              # noqa: E402

F, N = 6, 300
ref_xyz = rng.normal(0, 8, (N, 3))
traj = np.stack([(ref_xyz @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 3, 3)
                  + rng.normal(0, 0.5, (N, 3))) for _ in range(F)]).astype(np.float32)
masses = rng.uniform(1, 32, N)
u = mda.Universe.empty(N, trajectory=True)
u.add_TopologyAttr("masses", masses)
u.load_new(traj.copy(), order="fac")
r = mda.Universe.empty(N, trajectory=True)
r.add_TopologyAttr("masses", masses)
r.load_new(ref_xyz.astype(np.float32)[None].copy(), order="fac")
for wname, ww in (("none", None), ("mass", masses)):
    R = rrms.RMSD(u, r, select="all", weights=None if ww is None else "mass").run()
    mine = orc.rmsd_trajectory(traj, ref_xyz.astype(np.float32), weights=ww)
    check(f"RMSD.run weights={wname}", np.array_equal(R.results.rmsd[:, 2], mine), f"max={mine.max():.4f}")
    u2 = mda.Universe.empty(N, trajectory=True)
    u2.add_TopologyAttr("masses", masses)
    u2.load_new(traj.copy(), order="fac")
    A = ralign.AlignTraj(u2, r, select="all", weights=None if ww is None else "mass", in_memory=True).run()
    al = u2.trajectory.timeseries(order="fac")
    mine_xyz, mine_rmsd = orc.align_trajectory(traj, ref_xyz.astype(np.float32), weights=ww)
    check(f"AlignTraj rmsd weights={wname}", np.array_equal(A.results.rmsd, mine_rmsd))
    check(f"AlignTraj coords weights={wname}", np.array_equal(al, mine_xyz),
          f"maxdiff={np.abs(al - mine_xyz).max():.3g}")

for bname in ("ortho", "tric2"):
    box = BOXES[bname]
    tv = orc.triclinic_vectors(box).astype(np.float64)
    F, N = 3, 600
    pos = np.stack([(rng.random((N, 3)) * 1.4 - 0.2) @ tv for _ in range(F)]).astype(np.float32)
    u = mda.Universe.empty(N, trajectory=True)
    u.load_new(pos.copy(), order="fac", dimensions=np.tile(box, (F, 1)))
    for rmax, nb in ((6.0, 75), (12.0, 150)):
        try:
            ref = InterRDF(u.atoms[:400], u.atoms[200:], nbins=nb, range=(0.0, rmax)).run()
        except ValueError as e:
            print("skip InterRDF", bname, rmax, e)
            continue
        mine = orc.interrdf(pos[:, :400], pos[:, 200:], box, nbins=nb, rrange=(0.0, rmax))
        check(f"InterRDF {bname} rmax={rmax}", same(ref.results.count, mine["count"]) and
              same(ref.results.edges, mine["edges"]) and np.allclose(ref.results.rdf, mine["rdf"], rtol=1e-13),
              f"sum={mine['count'].sum():.0f}")

print("FAILURES:", FAIL)
sys.exit(1 if FAIL else 0)
