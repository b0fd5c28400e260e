#!/usr/bin/env python
"""Third set of golden fixtures: DistanceMatrix (analysis/diffusionmap.py) from the REAL reference.
TEST INFRASTRUCTURE.  The reference is the out-of-tree build baseline/build_ref.sh installs:

    MDA_REF_PATH=baseline/_ref python tests/golden/make_golden_r01h.py

Writes tests/golden/reference_vectors_r01h.npz: the coordinates handed to and the matrices returned
by the unmodified reference (MDAnalysis 2.8.0-dev0) on frames of its own adk PSF/DCD fixture.
Nothing at test time reads /root/reference.
"""
import functools
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in os.environ.get("MDA_REF_PATH", os.path.join(HERE, "..", "..", "baseline", "_ref")).split(":"):
    sys.path.insert(0, os.path.abspath(p))
warnings.filterwarnings("ignore")

import MDAnalysis as mda                                         # noqa: E402
from MDAnalysis.analysis.diffusionmap import DistanceMatrix     # noqa: E402
from MDAnalysis.analysis.rms import rmsd                         # noqa: E402
from MDAnalysisTests.datafiles import DCD, PSF                   # noqa: E402

u = mda.Universe(PSF, DCD)
ca = u.select_atoms("name CA")
frames = np.arange(0, 98, 7)                                     # 14 frames
out = {"frames": frames, "ca_indices": ca.indices.astype(np.int64),
       "xyz": np.stack([u.atoms.positions.copy() for _ in u.trajectory[frames]]).astype(np.float32)}
masses = ca.masses.astype(np.float64)
wts = np.random.default_rng(20261019).uniform(0.5, 3.0, len(ca))
out["weights"] = wts
out["dm/ca"] = DistanceMatrix(u, select="name CA").run(frames=frames).results.dist_matrix
out["dm/all"] = DistanceMatrix(u, select="all").run(frames=frames).results.dist_matrix
out["dm/ca_weighted"] = DistanceMatrix(u, select="name CA", weights=wts).run(frames=frames).results.dist_matrix
out["dm/ca_cutoff"] = DistanceMatrix(u, select="name CA", cutoff=1.5).run(frames=frames).results.dist_matrix
fit = functools.partial(rmsd, superposition=True)
out["dmfit/ca"] = DistanceMatrix(u, select="name CA", metric=fit).run(frames=frames).results.dist_matrix
out["dmfit/ca_weighted"] = DistanceMatrix(u, select="name CA", metric=fit, weights=wts).run(frames=frames).results.dist_matrix
np.savez_compressed(os.path.join(HERE, "reference_vectors_r01h.npz"), **out)
for k, v in out.items():
    print(k, v.shape, v.dtype, float(np.max(v)) if v.size else "")
