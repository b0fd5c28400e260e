"""-m gpu: the REFERENCE'S OWN test modules for the hot path, run against the CUDA backend.

baseline/_ref holds the unmodified reference and the part of its MDAnalysisTests package the
hot path needs (baseline/build_ref.sh).  Each module is run by pytest in a subprocess with
tests/ref_gpu_plugin.py loaded: the CUDA module masquerades as the 'serial' and 'openmp'
backends of lib.distances, capped_distance / self_capped_distance default to it, and
nsgrid.FastNS / qcprot.CalcRMSDRotationalMatrix are the GPU twins.  What must hold: every test
of the module passes except the listed ones, which probe Cython-level behaviour the Python twin
does not have (documented next to each).  Skipped when baseline/_ref is absent.
"""
import json
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

# module -> test ids (substrings) allowed to fail, with the reason
MODULES = {
    "lib/test_distances.py": {},
    "lib/test_nsgrid.py": {},
    "lib/test_qcprot.py": {},
    "analysis/test_rdf.py": {},
    "analysis/test_rdf_s.py": {},
    # client_RMSD0 = the reference's multiprocessing executor: it fork()s workers, and a CUDA context
    # does not survive fork ("initialization error"); the serial and dask-less variants pass.  The
    # executor route for the GPU is adapters.cuda_backend() (tests/test_gpu_reference_dropin.py).
    "analysis/test_rms.py": {"client_RMSD0": "multiprocessing fork + CUDA"},
    # these three load adk_oplsaa.tpr; the TPR parser needs mda_xdrlib, which is a stub in baseline/_ref
    # (they fail the same way with the reference's own backend in this environment)
    "analysis/test_contacts.py": {"test_distance_box": "TPR parser stub", "test_n_initial_contacts[datafiles1": "TPR parser stub"},
    "analysis/test_distances.py": {},
    "analysis/test_align.py": {},
    "lib/test_neighborsearch.py": {},      # AtomNeighborSearch -> capped_distance (lib/NeighborSearch.py:105)
    # DistanceMatrix.run -> mdagpu_pairwise_rmsd (register.install(patch_analysis=True)); DiffusionMap on top
    "analysis/test_diffusionmap.py": {},
    # distance-based selections (around, sphlayer, sphzone, cylayer, cyzone, point, isolayer) call
    # lib.distances.capped_distance (core/selection.py:321-531); "-k": only those tests, the other
    # selection tests need data files that are not part of baseline/_ref
    "core/test_atomselections.py": {"-k": "DistanceSelections or around or sphlayer or sphzone or isolayer or point"},
    # callers after the path: hydrogen bonds and water bridges (capped_distance +
    # calc_angles), bond / angle / dihedral values of topology objects (calc_bonds / calc_angles / calc_dihedrals),
    # guess_bonds (self_capped_distance), Dihedral / Ramachandran / Janin (calc_dihedrals)
    "analysis/test_hydrogenbonds_analysis.py": {"client_HydrogenBondAnalysis0": "multiprocessing fork + CUDA"},
    "analysis/test_wbridge.py": {},
    "analysis/test_atomicdistances.py": {},
    "core/test_topologyobjects.py": {},
    "guesser/test_default_guesser.py": {},
    # client_*0 = the multiprocessing executor (fork + CUDA, as for test_rms.py)
    "analysis/test_dihedrals.py": {"test_plot": "matplotlib is a stub in baseline/_ref", "test_janin_tpr": "TPR parser stub",
                                   "client_Ramachandran0": "multiprocessing fork + CUDA", "client_Janin0": "multiprocessing fork + CUDA",
                                   "client_Dihedral0": "multiprocessing fork + CUDA"},
}

# The substrings above only say WHY a family of tests cannot pass here.  WHICH tests may fail is the exact
# list of pytest ids in tests/golden/reference_allowed_failures.json (recorded on the GPU box with
# MDAGPU_RECORD_REF_FAILURES=<file>, every id matching one of the reasons above): a new failure whose
# id merely contains "test_distance_box" or "client_RMSD0" is NOT excused.
EXACT_FILE = os.path.join(ROOT, "tests", "golden", "reference_allowed_failures.json")
try:
    with open(EXACT_FILE) as _fh:
        EXACT = json.load(_fh)
except OSError:
    EXACT = None


@pytest.mark.parametrize("module", list(MODULES))
def test_reference_module_on_cuda_backend(module):
    path = os.path.join(REF, "MDAnalysisTests", module)
    if not os.path.exists(path):
        pytest.skip("baseline/_ref/MDAnalysisTests is not present")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REF, os.path.join(ROOT, "tests"), ROOT])
    cmd = [sys.executable, "-m", "pytest", path, "-q", "-p", "ref_gpu_plugin", "-p", "no:cacheprovider",
           "-W", "ignore", "--no-header", "-rfE", "--disable-warnings"]
    if "-k" in MODULES[module]:
        cmd += ["-k", MODULES[module]["-k"]]
    out = subprocess.run(cmd, cwd=REF, env=env, capture_output=True, text=True, timeout=1500)
    text = out.stdout + out.stderr
    tail = "\n".join(text.splitlines()[-40:])
    m = re.search(r"(\d+) passed", text)
    assert m, f"no tests ran:\n{tail}"
    launches = re.search(r"mdagpu kernel launches during this session: (\d+)", text)
    assert launches and int(launches.group(1)) > 0, f"the CUDA backend was not exercised:\n{tail}"
    failed = re.findall(r"^(?:FAILED|ERROR) (\S+)", text, flags=re.M)    # fixture errors count as failures
    allowed = {k: v for k, v in MODULES[module].items() if k != "-k"}
    unexpected = [f for f in failed if not any(a in f for a in allowed)]
    record = os.environ.get("MDAGPU_RECORD_REF_FAILURES")
    if record:
        seen = {}
        if os.path.exists(record):
            with open(record) as fh:
                seen = json.load(fh)
        seen[module] = {f: next(v for a, v in allowed.items() if a in f) for f in sorted(failed)
                        if any(a in f for a in allowed)}
        with open(record, "w") as fh:
            json.dump(seen, fh, indent=1, sort_keys=True)
    elif EXACT is not None:
        unexpected = [f for f in failed if f not in EXACT.get(module, {})]
    assert not unexpected, f"{len(unexpected)} unexpected failures, e.g. {unexpected[:5]}\n{tail}"
    print(f"{module}: {m.group(1)} passed, {len(failed)} allowed failures, {launches.group(1)} kernel launches")
