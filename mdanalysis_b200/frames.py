"""A block of trajectory frames: what the GPU analyses consume.

The reference analyses pull one ``Timestep`` at a time from a reader inside a
Python loop (``analysis/base.py:515-523``).  The GPU path wants the frames of
its shard as ONE array ``[F, n_atoms, 3]`` float32 (plus ``[F, 6]`` unit cells),
which is exactly what ``ReaderBase.timeseries(order='fac')`` /
``MemoryReader.get_array()`` deliver (``coordinates/base.py:992``,
``coordinates/memory.py:480-491``).  ``FrameBlock`` is that array with the few
attributes the analyses read; ``FrameBlock.from_atomgroup`` builds it from a
real MDAnalysis AtomGroup when MDAnalysis is importable.
"""
from __future__ import annotations

import threading

import numpy as np

from ._cabi import is_tensor

# a generic reader is moved frame by frame: one ingest at a time (the multi-GPU executor of
# adapters.py reads the parts of a run from several threads)
_READER_LOCK = threading.Lock()


class FrameBlock:
    """positions [F, n, 3] float32 (NumPy or torch CUDA), dimensions [F, 6] float32 or None."""

    def __init__(self, positions, dimensions=None, frames=None, times=None, masses=None):
        if is_tensor(positions):
            import torch
            if positions.ndim == 2:
                positions = positions[None]
            self.positions = positions.to(torch.float32).contiguous()
        else:
            positions = np.asarray(positions)
            if positions.ndim == 2:
                positions = positions[None]
            self.positions = np.ascontiguousarray(positions, dtype=np.float32)
        if self.positions.ndim != 3 or self.positions.shape[2] != 3:
            raise ValueError(f"positions must have shape (F, n, 3), got {tuple(self.positions.shape)}")
        F = self.n_frames
        if dimensions is None:
            self.dimensions = None
        else:
            dims = np.asarray(dimensions, dtype=np.float32)
            if dims.ndim == 1:
                dims = np.broadcast_to(dims, (F, 6))
            if dims.shape != (F, 6):
                raise ValueError(f"dimensions must have shape (6,) or ({F}, 6), got {dims.shape}")
            self.dimensions = np.ascontiguousarray(dims)
        self.frames = np.arange(F) if frames is None else np.asarray(frames)
        self.times = self.frames.astype(np.float64) if times is None else np.asarray(times, dtype=np.float64)
        self.masses = None if masses is None else np.asarray(masses, dtype=np.float64)

    @property
    def n_frames(self):
        return int(self.positions.shape[0])

    @property
    def n_atoms(self):
        return int(self.positions.shape[1])

    @property
    def on_device(self):
        return is_tensor(self.positions)

    def __len__(self):
        return self.n_atoms

    def constant_box(self):
        if self.dimensions is None or self.n_frames == 0:
            return True
        return bool((self.dimensions == self.dimensions[0]).all())

    def slice_frames(self, sl):
        """Frames ``sl`` (slice or index array) as a new block sharing memory where possible."""
        dims = None if self.dimensions is None else self.dimensions[sl]
        return FrameBlock(self.positions[sl], dims, self.frames[sl], self.times[sl], self.masses)

    def select(self, indices):
        """Atoms ``indices`` of every frame (copy)."""
        idx = np.asarray(indices)
        if self.on_device:
            import torch
            pos = self.positions[:, torch.as_tensor(idx, device=self.positions.device)]
        else:
            pos = self.positions[:, idx]
        m = None if self.masses is None else self.masses[idx]
        return FrameBlock(pos, self.dimensions, self.frames, self.times, m)

    def to_device(self, device="cuda"):
        if self.on_device:
            return self
        from ._cabi import upload
        return FrameBlock(upload(self.positions, device), self.dimensions, self.frames, self.times, self.masses)

    @classmethod
    def from_atomgroup(cls, atomgroup, start=None, stop=None, step=None, frames=None):
        """Bulk ingest of the atoms of ``atomgroup`` for the frames ``[start:stop:step]`` (or the
        explicit ``frames``) of its universe's trajectory -- the GPU twin of the reference's
        per-frame loop (``analysis/base.py:515-523``).

        * ``DCDReader``: the native bulk reader (``coordinates/dcd.py`` -> ``csrc/dcd.cu``) on the
          reader's file: only the selected atoms of the selected frames are read, straight into
          pinned memory; unit cells and times come out of the same read (``coordinates/DCD.py:211-256``).
        * ``MemoryReader``: the coordinate array itself (``coordinates/memory.py:480-484``), unit
          cells from ``dimensions_array``, ``time = frame * dt`` (``memory.py:606``).
        * any other reader (or a reader with on-the-fly transformations): ``timeseries`` with the
          frame range pushed down, unit cells and times from one pass over those frames.

        The reader is left on the frame it was on (the reference's analyses snapshot
        ``ts`` where ``run()`` is called, e.g. ``analysis/align.py:836-845``)."""
        u = atomgroup.universe
        traj = u.trajectory
        n_all = len(traj)
        if frames is not None:
            fidx = np.asarray(frames, dtype=np.int64)
            regular = None
        else:
            regular = slice(start, stop, step)
            fidx = np.arange(n_all)[regular]
        aidx = np.asarray(atomgroup.indices, dtype=np.int64)
        whole = len(aidx) == u.atoms.n_atoms and bool(np.array_equal(aidx, np.arange(len(aidx))))
        masses = getattr(atomgroup, "masses", None)
        kind = type(traj).__name__
        plain = not getattr(traj, "transformations", ())
        if kind == "DCDReader" and plain and getattr(traj, "convert_units", True) is not None:
            from .coordinates.dcd import DCDFile, _dims_from_unitcell, AKMA_TO_PS
            with DCDFile(traj.filename) as f:
                if regular is not None:
                    r = range(*regular.indices(f.n_frames))
                    xyz, cell = f.readframes(r.start, r.stop, r.step, indices=None if whole else aidx)
                else:
                    from .coordinates.dcd import _host_buffer
                    xyz = _host_buffer((len(fidx), len(aidx), 3), True)
                    cell = np.zeros((len(fidx), 6))
                    for k, fr in enumerate(fidx):          # arbitrary frame lists: one read each
                        f.readframes(int(fr), int(fr) + 1, 1, indices=None if whole else aidx, out=xyz[k:k + 1])
                        cell[k] = f.readframes(int(fr), int(fr) + 1, 1, indices=aidx[:1])[1][0]
                h = f.header
            dims = _dims_from_unitcell(cell) if h["is_periodic"] else None
            if dims is not None and not np.any(dims):
                dims = None
            nsavc = h["nsavc"] if h["nsavc"] else 1
            times = (fidx + h["istart"] / nsavc) * (h["delta"] * nsavc * AKMA_TO_PS)
            return cls(xyz, dims, fidx, times, masses)
        if kind == "MemoryReader" and plain:
            arr = traj.get_array()
            order = traj.stored_order
            view = np.transpose(arr, (order.find("f"), order.find("a"), order.find("c")))
            # a regular frame range is a VIEW of the reader's array (no copy of the trajectory)
            pos = view[regular] if regular is not None else view[fidx]
            if not whole:
                pos = pos[:, aidx]
            dims = np.asarray(traj.dimensions_array, dtype=np.float32)[fidx]
            dims = None if not np.any(dims) else dims
            return cls(pos, dims, fidx, fidx * float(traj.dt), masses)
        # generic reader: bulk coordinates, one pass for unit cells and times
        _READER_LOCK.acquire()
        cur = traj.ts.frame
        try:
            if regular is not None and len(fidx):
                st = fidx[1] - fidx[0] if len(fidx) > 1 else 1
                pos = np.asarray(traj.timeseries(atomgroup=atomgroup, start=int(fidx[0]), stop=int(fidx[-1]) + 1,
                                                 step=int(st), order="fac"))
                walk = traj[int(fidx[0]):int(fidx[-1]) + 1:int(st)]
            else:
                pos = np.empty((len(fidx), len(aidx), 3), dtype=np.float32)
                walk = traj[fidx] if len(fidx) else []
            dims, times = [], []
            for k, ts in enumerate(walk):
                if regular is None:
                    pos[k] = atomgroup.positions
                dims.append(None if ts.dimensions is None else np.array(ts.dimensions, dtype=np.float32))
                times.append(ts.time)
        finally:
            if 0 <= cur < n_all:
                traj[cur]
            _READER_LOCK.release()
        dimensions = None if (not dims or any(d is None for d in dims)) else np.stack(dims)
        return cls(pos, dimensions, fidx, np.asarray(times), masses)


def as_block(x, dimensions=None):
    """FrameBlock | array [F,n,3] | array [n,3] | AtomGroup-like -> FrameBlock."""
    if isinstance(x, FrameBlock):
        return x
    if hasattr(x, "universe") and hasattr(x, "positions"):
        return FrameBlock.from_atomgroup(x)
    return FrameBlock(x, dimensions)
