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

import numpy as np

from ._cabi import is_tensor


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
        return self.dimensions is None or bool((self.dimensions == self.dimensions[0]).all())

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
        import torch
        return FrameBlock(torch.as_tensor(self.positions).to(device), self.dimensions, self.frames,
                          self.times, self.masses)

    @classmethod
    def from_atomgroup(cls, atomgroup, start=None, stop=None, step=None):
        """Bulk ingest from an MDAnalysis AtomGroup (needs MDAnalysis)."""
        u = atomgroup.universe
        traj = u.trajectory
        sl = slice(start, stop, step)
        frames = np.arange(len(traj))[sl]
        # one bulk read of every frame (MemoryReader.timeseries has inclusive-stop semantics and
        # rejects stop=None, coordinates/memory.py:545-567), sliced here
        pos = np.asarray(traj.timeseries(atomgroup, order="fac"))[sl]
        dims, times = [], []
        for ts in traj[sl]:
            dims.append(None if ts.dimensions is None else np.array(ts.dimensions, dtype=np.float32))
            times.append(ts.time)
        dimensions = None if any(d is None for d in dims) else np.stack(dims)
        masses = getattr(atomgroup, "masses", None)
        return cls(pos, dimensions, frames, np.asarray(times), masses)


def as_block(x, dimensions=None):
    """FrameBlock | array [F,n,3] | array [n,3] | AtomGroup-like -> FrameBlock."""
    if isinstance(x, FrameBlock):
        return x
    if hasattr(x, "universe") and hasattr(x, "positions"):
        return FrameBlock.from_atomgroup(x)
    return FrameBlock(x, dimensions)
