"""Bulk DCD reader -- the GPU pipeline's twin of ``DCDReader.timeseries``.

Reference: ``coordinates/DCD.py:279-344`` (``timeseries`` -> ``DCDFile.readframes``,
``lib/formats/libdcd.pyx:583-700``) and ``DCD.py:211-256`` (``_frame_to_ts``: unit cell
order, cosines vs degrees, time of a frame).  The file is parsed by the native reader in
``csrc/dcd.cu`` (``mdagpu_dcd_*``): frames land directly in one ``[n, nsel, 3]`` float32
buffer -- pinned host memory by default, so the upload that follows runs at PCIe rate --
and the unit cells / times come out of the same bulk read (no per-frame Python loop).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .._cabi import check, load, ptr

# AKMA time unit -> ps (MDAnalysis.units.timeUnit_factor['AKMA'], DCDReader converts dt with it:
# coordinates/DCD.py:143-150 via convert_time_from_native)
AKMA_TO_PS = 4.888821e-2


def _dims_from_unitcell(uc):
    """[n, 6] file-order cells -> [n, 6] float32 [A, B, C, alpha, beta, gamma] (DCD.py:225-256)."""
    uc = np.asarray(uc, dtype=np.float64)
    out = uc[:, [0, 2, 5, 4, 3, 1]].copy()
    ang = out[:, 3:]
    cosines = np.all((ang >= -1.0) & (ang <= 1.0), axis=1)
    pi_2 = np.pi / 2
    out[cosines, 3:] = 90.0 - np.arcsin(ang[cosines]) * 90.0 / pi_2
    weird = ~cosines & (np.any(out < 0.0, axis=1) | np.any(out[:, 3:] > 180.0, axis=1))
    if weird.any():
        # new CHARMM: the six numbers are the lower triangle of the box matrix (DCD.py:241-246)
        from ..lib.mdamath import triclinic_box
        for k in np.flatnonzero(weird):
            H = uc[k]
            out[k] = triclinic_box(H[[0, 1, 3]], H[[1, 2, 4]], H[[3, 4, 5]])
    return out.astype(np.float32)


class DCDFile:
    """Read-only DCD file; ``readframes`` mirrors ``libdcd.DCDFile.readframes`` for order 'fac'."""

    def __init__(self, filename):
        self.filename = str(filename)
        self._h = C.c_void_p()
        check(load().mdagpu_dcd_open(self.filename.encode(), C.byref(self._h)))
        na, nf = C.c_int64(), C.c_int64()
        cell, charmm = C.c_int(), C.c_int()
        istart, nsavc = C.c_int32(), C.c_int32()
        delta = C.c_double()
        check(load().mdagpu_dcd_header(self._h, C.byref(na), C.byref(nf), C.byref(cell), C.byref(istart),
                                       C.byref(nsavc), C.byref(delta), C.byref(charmm)))
        self.header = {"natoms": na.value, "istart": istart.value, "nsavc": nsavc.value, "delta": delta.value,
                       "is_periodic": bool(cell.value), "charmm": charmm.value}
        self.n_frames = nf.value
        self.natoms = na.value

    def __len__(self):
        return self.n_frames

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            load().mdagpu_dcd_close(h)

    __del__ = close

    def readframes(self, start=None, stop=None, step=None, indices=None, out=None, pinned=True):
        """Frames ``range(start, stop, step)`` -> (xyz float32 [n, nsel, 3], unitcell float64 [n, 6]).

        ``unitcell`` is in the file's order ``[A, gamma, B, beta, alpha, C]`` like the reference's
        ``readframes``.  ``out``: destination array (e.g. a slice of a pinned staging buffer);
        otherwise a new array, backed by pinned memory when torch + CUDA are available.
        """
        if self._h is None:
            raise IOError("No file open")
        if self.n_frames == 0:
            raise IOError("opened empty file. No frames are saved")
        if start is None and stop is None and step is not None and step < 0:
            start, stop = self.n_frames - 1, -1
        rng = range(*slice(start, stop, step).indices(self.n_frames)) if not (stop == -1 and (step or 1) < 0) \
            else range(start, -1, step)
        n = len(rng)
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int64)
        nsel = self.natoms if idx is None else len(idx)
        if out is None:
            out = _host_buffer((n, nsel, 3), pinned)
        elif out.shape != (n, nsel, 3) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float32 array of shape {(n, nsel, 3)}")
        cell = np.zeros((n, 6), dtype=np.float64)
        if n:
            st = rng.step
            check(load().mdagpu_dcd_readframes(self._h, rng.start, rng.start + n * st, st, ptr(idx), nsel,
                                               ptr(out), ptr(cell)))
        return out, cell


def _host_buffer(shape, pinned):
    if pinned:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
        except Exception:
            pass
    return np.empty(shape, dtype=np.float32)


def read_dcd(filename, start=None, stop=None, step=None, indices=None, pinned=True):
    """One bulk read -> ``FrameBlock`` with dimensions, frame numbers and times like ``DCDReader``."""
    from ..frames import FrameBlock
    with DCDFile(filename) as f:
        frames = np.arange(f.n_frames)[slice(start, stop, step)]
        xyz, cell = f.readframes(start, stop, step, indices=indices, pinned=pinned)
        h = f.header
    dims = _dims_from_unitcell(cell) if h["is_periodic"] else None
    # DCD.py:143-150, 215: dt = delta * nsavc (AKMA -> ps); time = (frame + istart / nsavc) * dt
    nsavc = h["nsavc"] if h["nsavc"] else 1
    dt = h["delta"] * nsavc * AKMA_TO_PS
    times = (frames + h["istart"] / nsavc) * dt
    if dims is not None and not np.any(dims):
        dims = None
    return FrameBlock(xyz, dims, frames, times)
