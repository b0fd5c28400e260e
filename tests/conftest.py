import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def hostcheck():
    """CPU build of the device arithmetic in csrc/geom.cuh (checker only)."""
    src = os.path.join(ROOT, "tests", "host_check.cu")
    out = os.path.join(ROOT, "tests", "_build", "libhostcheck.so")
    hdr = os.path.join(ROOT, "mdanalysis_b200", "csrc", "geom.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if (not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        nvcc = "/usr/local/cuda/bin/nvcc"
        subprocess.check_call([nvcc, "-O2", "--fmad=false", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                               "-fPIC,-ffp-contract=off", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    lib.hc_qcp_solve.restype = C.c_double
    return lib


BOXES = {
    "none": None,
    "ortho": np.array([30, 25, 40, 90, 90, 90], dtype=np.float32),
    "tric": np.array([40, 35, 50, 70, 80, 100], dtype=np.float32),
    "tric60": np.array([30, 30, 30, 60, 60, 90], dtype=np.float32),
}


def canon_pairs(pairs, d, self_pairs=False):
    """Sort a pair list (and its distances) into a canonical order."""
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    d = np.asarray(d)
    if self_pairs:
        p = np.sort(p, axis=1)
    o = np.lexsort((p[:, 1], p[:, 0]))
    return p[o], d[o]
