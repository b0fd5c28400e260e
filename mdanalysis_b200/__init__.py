"""mdanalysis_b200 -- a B200 (sm_100a) backend for MDAnalysis's geometry hot path.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (``include/mdagpu.h``) -> libmdagpu.so
  _cabi.py     ctypes binding of that ABI
  lib/         host-side mirror of ``MDAnalysis.lib`` for the path
               (distances, nsgrid.FastNS, qcprot, mdamath, util)
  analysis/    block-of-frames drivers: RMSD, AlignTraj, InterRDF, InterRDF_s, Contacts
  frames.py    FrameBlock: [F, n, 3] float32 block of trajectory frames
  parallel.py  frame sharding over ranks + NCCL reductions (torch.distributed)
  register.py  plugs the backend into an installed MDAnalysis (``backend='cuda'``)
  adapters.py  run_on_gpu(analysis): an MDAnalysis analysis object computed by the GPU classes

There is no CPU fallback: every numerical result comes from libmdagpu.so.
"""
from . import _cabi
from .frames import FrameBlock
from . import lib, analysis, parallel, register  # noqa: F401
from .adapters import cuda_backend, gpu_enabled, run_on_gpu

__version__ = "0.1.0"
__all__ = ["FrameBlock", "lib", "analysis", "parallel", "register", "library_path", "run_on_gpu", "cuda_backend", "gpu_enabled"]


def library_path():
    return _cabi.LIB_PATH
