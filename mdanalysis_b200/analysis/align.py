"""Trajectory alignment on the GPU -- mirror of ``MDAnalysis/analysis/align.py``
(rotation_matrix :224-307, _fit_to :310-366, AlignTraj :666-867).
"""
import numpy as np

from .._cabi import DEVICE, check, current_stream, download, is_tensor, load, tensor_ptr, upload
from ..frames import as_block
from ..lib import qcprot as qcp
from .base import BlockAnalysis
from .rms import _resolve_selection, _whole, get_weights


def rotation_matrix(a, b, weights=None):
    """Rotation that minimises the RMSD of ``a`` onto ``b`` (both centred).

    Returns ``(R, rmsd)``; ``R`` acts as ``a_rotated = a @ R.T`` (align.py:284-307),
    including the ``np.allclose`` early exit (:290-291).
    """
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        raise ValueError("'a' and 'b' must have same shape")
    if np.allclose(a, b) and weights is None:
        return np.eye(3, dtype=np.float64), 0.0
    N = b.shape[0]
    if weights is not None:
        weights = np.asarray(weights, dtype=np.float64) / np.mean(weights)
    rot = np.zeros(9, dtype=np.float64)
    val = qcp.CalcRMSDRotationalMatrix(a, b, N, rot, weights)
    return rot.reshape(3, 3), val


class AlignTraj(BlockAnalysis):
    """RMS-fit every frame of ``mobile`` onto ``reference`` (in memory).

    mobile     FrameBlock | [F, n, 3] array / CUDA tensor | AtomGroup; ALL its
               atoms are moved, ``select`` picks the atoms that define the fit
    reference  [n, 3] positions (or a block; frame ``ref_frame``) of the same atoms
    weights    None | 'mass' | array over the fitted atoms
    After ``run()``: ``results.rmsd`` (float64 [F]) and ``results.positions``
    (aligned coordinates, same kind as the input; ``in_memory=True`` with a
    FrameBlock input rewrites the block's positions, as the reference rewrites
    its MemoryReader).
    """

    def __init__(self, mobile, reference, select="all", weights=None, in_memory=True,
                 ref_frame=0, tol_mass=0.1, **kwargs):
        blk = as_block(mobile)
        super().__init__(blk, **kwargs)
        self._ref = as_block(reference)
        self._sel_m = _whole(_resolve_selection(mobile, select["mobile"] if isinstance(select, dict) else select), blk.n_atoms)
        self._sel_r = _whole(_resolve_selection(reference, select["reference"] if isinstance(select, dict) else select),
                             self._ref.n_atoms)
        self._weights = weights
        self._in_memory = in_memory
        self.ref_frame = ref_frame
        n_m = blk.n_atoms if self._sel_m is None else len(self._sel_m)
        n_r = self._ref.n_atoms if self._sel_r is None else len(self._sel_r)
        if n_m != n_r:
            raise ValueError("Reference and trajectory atom selections do not contain the same "
                             f"number of atoms: N_ref={n_r:d}, N_traj={n_m:d}")

    def _prepare(self, blk):
        ref = self._ref.positions[self.ref_frame]
        if is_tensor(ref):
            ref = ref.detach().cpu().numpy()
        ref = np.asarray(ref, dtype=np.float32)
        self._ref_sel = ref if self._sel_r is None else ref[self._sel_r]
        self.results.rmsd = np.zeros((blk.n_frames,))

    def _compute(self, blk):
        w = None
        if self._weights is not None:
            if isinstance(self._weights, str):
                w = get_weights(self._ref, self._weights)
                w = w if self._sel_r is None else w[self._sel_r]
            else:
                w = np.asarray(self._weights, dtype=np.float64)
            w = w / np.mean(w)            # align.py:297
        vals, _, moved = qcp.qcp_trajectory(
            blk.positions, self._ref_sel, select=self._sel_m, weights=w, align=True, convention=1,
            allclose_shortcut=(w is None), out=blk.positions if self._in_memory else None)
        self.results.rmsd = vals.detach().cpu().numpy() if is_tensor(vals) else vals
        self.results.positions = moved
        if self._in_memory:
            blk.positions = moved


class AverageStructure(BlockAnalysis):
    """Average structure after superposing every frame on a reference
    (``analysis/align.py:870-1118``).

    mobile     FrameBlock | [F, n, 3] array / CUDA tensor | AtomGroup; ALL atoms are moved
               (``in_memory=True`` rewrites the block, as the reference rewrites its reader)
    reference  block / positions of the reference (default: ``mobile``), frame ``ref_frame``
    select     atoms that define the fit AND whose positions are averaged
    After ``run()``: ``results.positions`` (float64 ``[n_sel, 3]``, the mean of the fitted
    selected atoms over the frames) and ``results.rmsd`` (mean best-fit RMSD).  One
    ``mdagpu_qcp_trajectory(align=1)`` launch superposes all frames; the mean is a float64
    reduction over the aligned block.
    """

    def __init__(self, mobile, reference=None, select="all", weights=None, ref_frame=0, in_memory=True,
                 **kwargs):
        blk = as_block(mobile)
        super().__init__(blk, **kwargs)
        self._ref = blk if reference is None else as_block(reference)
        self._sel_m = _resolve_selection(mobile, select["mobile"] if isinstance(select, dict) else select)
        self._sel_r = _resolve_selection(reference if reference is not None else mobile,
                                         select["reference"] if isinstance(select, dict) else select)
        self._weights = weights
        self._in_memory = in_memory
        self.ref_frame = ref_frame
        n_m = blk.n_atoms if self._sel_m is None else len(self._sel_m)
        n_r = self._ref.n_atoms if self._sel_r is None else len(self._sel_r)
        if n_m != n_r:
            raise ValueError("Reference and trajectory atom selections do not contain the same "
                             f"number of atoms: N_ref={n_r:d}, N_traj={n_m:d}")

    def _prepare(self, blk):
        ref = self._ref.positions[self.ref_frame]
        if is_tensor(ref):
            ref = ref.detach().cpu().numpy()
        ref = np.array(ref, dtype=np.float32)                  # a copy: `mobile` may be rewritten below
        self._ref_sel = ref if self._sel_r is None else ref[self._sel_r]

    def _compute(self, blk):
        w = None
        if self._weights is not None:
            if isinstance(self._weights, str):
                w = get_weights(self._ref, self._weights)
                w = w if self._sel_r is None else w[self._sel_r]
            else:
                w = np.asarray(self._weights, dtype=np.float64)
            w = w / np.mean(w)
        import torch
        host_input = not blk.on_device
        pos = upload(blk.positions) if host_input else blk.positions     # reduce on the device
        vals, _, moved = qcp.qcp_trajectory(
            pos, self._ref_sel, select=self._sel_m, weights=w, align=True, convention=1,
            allclose_shortcut=(w is None), out=pos if (self._in_memory or host_input) else None)
        if self._in_memory:
            if host_input:
                download(moved, out=blk.positions)
            else:
                blk.positions = moved
        # float64 mean over the frames of the fitted atoms: mdagpu_frame_mean (csrc/reduce.cu)
        sel = None if self._sel_m is None else torch.as_tensor(np.ascontiguousarray(self._sel_m, dtype=np.int32),
                                                               device=moved.device)
        nsel = moved.shape[1] if sel is None else int(sel.shape[0])
        mean = torch.empty((nsel, 3), dtype=torch.float64, device=moved.device)
        check(load().mdagpu_frame_mean(tensor_ptr(moved), moved.shape[0], moved.shape[1], tensor_ptr(sel), nsel,
                                       tensor_ptr(mean), DEVICE, current_stream(moved)))
        self.results.positions = mean.cpu().numpy()
        self.results.rmsd = float(vals.sum().item() / vals.shape[0])
