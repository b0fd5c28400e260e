"""Minimal run()/results plumbing shared by the GPU analyses.

Mirrors what user code touches on ``MDAnalysis.analysis.base.AnalysisBase``
(``analysis/base.py:697-833``): ``run(start, stop, step, frames)``,
``.results`` (attribute + item access, ``analysis/results.py:56-140``),
``.frames``, ``.times``, ``.n_frames``.  The frame loop itself does not exist
here: a whole block of frames goes to the GPU in one call.
"""
import numpy as np


class Results(dict):
    """dict with attribute access (analysis/results.py:56)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(f"'Results' object has no attribute '{k}'") from e

    def __setattr__(self, k, v):
        self[k] = v


class BlockAnalysis:
    def __init__(self, block, verbose=False):
        self._block = block
        self._verbose = verbose
        self.results = Results()

    def _slice(self, start=None, stop=None, step=None, frames=None):
        """Same frame selection rules as AnalysisBase._setup_frames (base.py:420-478)."""
        if frames is not None:
            if not all(opt is None for opt in (start, stop, step)):
                raise ValueError("start/stop/step cannot be combined with frames")
            idx = np.asarray(frames)
            if idx.dtype == bool:
                idx = np.flatnonzero(idx)
            return idx
        return np.arange(self._block.n_frames)[slice(start, stop, step)]

    def run(self, start=None, stop=None, step=None, frames=None, verbose=None, **kwargs):
        idx = self._slice(start, stop, step, frames)
        if len(idx) == self._block.n_frames and np.array_equal(idx, np.arange(len(idx))):
            blk = self._block
        else:
            blk = self._block.slice_frames(idx)
        self.n_frames = blk.n_frames
        self.frames = blk.frames.copy()
        self.times = blk.times.copy()
        self._prepare(blk)
        self._compute(blk)
        self._conclude()
        return self

    def _prepare(self, blk):
        pass

    def _compute(self, blk):
        raise NotImplementedError

    def _conclude(self):
        pass
