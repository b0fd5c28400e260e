"""RMSD on the GPU -- mirror of ``MDAnalysis/analysis/rms.py`` (rmsd :178-283, RMSD :335-736).

``RMSD(mobile, reference, select=..., weights=...).run()`` fills
``results.rmsd`` with rows ``[frame, time, rmsd, (group rmsds...)]`` exactly like
the reference, but the whole block of frames is centred, superposed (QCP) and
reduced in ONE kernel launch (``mdagpu_qcp_trajectory``) instead of a Python
loop over ``_single_frame`` (rms.py:685-727).

Because this package has no selection language, ``select`` is an index array
(or ``None`` / ``'all'``); with real MDAnalysis AtomGroups a selection string
is resolved through ``select_atoms``.
"""
import numpy as np

from .._cabi import DEVICE, check, current_stream, is_tensor, load, tensor_ptr, upload
from ..frames import FrameBlock, as_block
from ..lib import qcprot as qcp
from .base import BlockAnalysis


def rmsd(a, b, weights=None, center=False, superposition=False):
    """RMSD between two coordinate sets (rms.py:178-283); QCP runs on the GPU."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    N = b.shape[0]
    if a.shape != b.shape:
        raise ValueError("a and b must have same shape")
    if center or superposition:
        a = a - np.average(a, axis=0, weights=weights)
        b = b - np.average(b, axis=0, weights=weights)
    if weights is not None:
        if len(weights) != len(a):
            raise ValueError("weights must have same length as a and b")
        weights = np.asarray(weights, dtype=np.float64) / np.mean(weights)
    if superposition:
        return qcp.CalcRMSDRotationalMatrix(a, b, N, None, weights)
    if weights is not None:
        return np.sqrt(np.sum(weights[:, np.newaxis] * ((a - b) ** 2)) / N)
    return np.sqrt(np.sum((a - b) ** 2) / N)


def _resolve_selection(group, select):
    """index array into ``group``'s atoms (None = all)."""
    if select is None or (isinstance(select, str) and select == "all"):
        return None
    if isinstance(select, str):
        if hasattr(group, "select_atoms"):
            sel = group.select_atoms(select)
            return np.searchsorted(group.indices, sel.indices)
        raise TypeError("selection strings need MDAnalysis AtomGroups; pass an index array")
    return np.asarray(select, dtype=np.int64)


def _whole(sel, n_atoms):
    """``None`` when the index list is every atom in order: the kernels then take the contiguous
    (TMA-streamed) path instead of a gather through an index list that changes nothing."""
    if sel is not None and len(sel) == n_atoms and np.array_equal(sel, np.arange(n_atoms)):
        return None
    return sel


def get_weights(block, weights):
    """util.get_weights: None | 'mass' | array (lib/util.py:1329-1397)."""
    if weights is None:
        return None
    if isinstance(weights, str):
        if weights != "mass":
            raise ValueError("weights must be {'mass', None} or an iterable of the same size as "
                             "the atomgroup.")
        if block.masses is None:
            raise TypeError("weights='mass' needs masses on the frame block")
        return block.masses
    w = np.asarray(weights)
    if w.ndim != 1:
        raise ValueError("weights must be a 1D array, not with shape {0}".format(w.shape))
    if len(w) != block.n_atoms:
        raise ValueError("weights (length {0}) must be of same length as the atoms ({1})"
                         .format(len(w), block.n_atoms))
    return w.astype(np.float64)


class RMSD(BlockAnalysis):
    """Best-fit RMSD of every frame of ``mobile`` to ``reference``.

    mobile     FrameBlock | [F, n, 3] array / CUDA tensor | AtomGroup
    reference  same kinds (default: ``mobile``); frame ``ref_frame`` is used
    select     indices of the atoms to superpose and measure (default all);
               a dict ``{'mobile': idx, 'reference': idx}`` selects each side
    groupselections  list of index arrays (or such dicts): RMSD of these groups
               after the superposition of ``select`` (no further fitting)
    weights    None | 'mass' | array over the selected atoms
    """

    def __init__(self, atomgroup, reference=None, select="all", groupselections=None,
                 weights=None, weights_groupselections=False, tol_mass=0.1, ref_frame=0, **kwargs):
        mobile = as_block(atomgroup)
        super().__init__(mobile, **kwargs)
        refblock = mobile if reference is None else as_block(reference)
        if isinstance(select, dict):
            sel_m = _resolve_selection(atomgroup, select["mobile"])
            sel_r = _resolve_selection(reference if reference is not None else atomgroup,
                                       select["reference"])
        else:
            sel_m = _resolve_selection(atomgroup, select)
            sel_r = _resolve_selection(reference if reference is not None else atomgroup, select)
        sel_m, sel_r = _whole(sel_m, mobile.n_atoms), _whole(sel_r, refblock.n_atoms)
        self._sel_m, self._sel_r = sel_m, sel_r
        self._refblock = refblock
        self.ref_frame = ref_frame
        self.weights = weights
        self.tol_mass = tol_mass
        n_m = mobile.n_atoms if sel_m is None else len(sel_m)
        n_r = refblock.n_atoms if sel_r is None else len(sel_r)
        if n_m != n_r:
            raise ValueError("Reference and trajectory atom selections do not contain the same "
                             f"number of atoms: N_ref={n_r:d}, N_traj={n_m:d}")
        if mobile.masses is not None and refblock.masses is not None:
            mm = mobile.masses if sel_m is None else mobile.masses[sel_m]
            mr = refblock.masses if sel_r is None else refblock.masses[sel_r]
            if np.any(np.absolute(mr - mm) > tol_mass):
                raise ValueError("Inconsistent selections, masses differ by more than"
                                 f"{tol_mass:f}")
        self._groups = []
        for g in (groupselections or []):
            if isinstance(g, dict):
                gm, gr = np.asarray(g["mobile"]), np.asarray(g["reference"])
            else:
                gm = gr = np.asarray(g)
            if len(gm) != len(gr):
                raise ValueError("Group selection: Reference and trajectory atom selections do "
                                 "not contain the same number of atoms")
            self._groups.append((gm, gr))
        self.weights_groupselections = weights_groupselections
        if weights_groupselections and len(weights_groupselections) != len(self._groups):
            raise ValueError("Length of weights_groupselections is not equal to length of "
                             "groupselections ")

    def _selected_weights(self, blk):
        sub = blk if self._sel_m is None else None
        if self.weights is None:
            return None
        if isinstance(self.weights, str):
            w = get_weights(blk, self.weights)
            w = w if self._sel_m is None else w[self._sel_m]
        else:
            w = np.asarray(self.weights, dtype=np.float64)
            n = blk.n_atoms if self._sel_m is None else len(self._sel_m)
            if w.ndim != 1 or len(w) != n:
                raise ValueError("weights must be a 1D array over the selected atoms")
        return np.asarray(w, dtype=np.float64) / np.mean(w)      # rms.py:639-643

    def _prepare(self, blk):
        ref_pos = self._refblock.positions[self.ref_frame]
        if is_tensor(ref_pos):
            ref_pos = ref_pos.detach().cpu().numpy()
        self._ref_all = np.asarray(ref_pos, dtype=np.float32)
        self._ref_sel = self._ref_all if self._sel_r is None else self._ref_all[self._sel_r]
        self.results.rmsd = np.zeros((blk.n_frames, 3 + len(self._groups)))

    def _compute(self, blk):
        w = self._selected_weights(blk)
        out = self.results.rmsd
        out[:, 0] = blk.frames
        out[:, 1] = blk.times
        if not self._groups:
            vals, _, _ = qcp.qcp_trajectory(blk.positions, self._ref_sel, select=self._sel_m,
                                            weights=w, convention=0)
            out[:, 2] = vals.detach().cpu().numpy() if is_tensor(vals) else vals
            return
        # secondary groups: superpose whole frames (rms.py:697-721), then plain RMSDs of the groups,
        # reduced on the device in float64
        import torch
        pos = blk.positions if blk.on_device else upload(blk.positions)
        vals, _, moved = qcp.qcp_trajectory(pos, self._ref_sel, select=self._sel_m,
                                            weights=w, align=True, convention=0)
        out[:, 2] = vals.detach().cpu().numpy()
        wg = self.weights_groupselections
        if not wg:
            wg = [self.weights if not np.iterable(self.weights) or isinstance(self.weights, str)
                  else None] * len(self._groups)
        for k, (gm, gr) in enumerate(self._groups):
            gw = wg[k]
            if isinstance(gw, str) and gw == "mass":
                gw = blk.masses[gm]
            # plain RMSD of the group on the superposed frames, one CTA per frame (csrc/reduce.cu)
            refpos = torch.as_tensor(np.ascontiguousarray(self._ref_all[gr], dtype=np.float64), device=moved.device)
            idx = torch.as_tensor(np.ascontiguousarray(gm, dtype=np.int32), device=moved.device)
            w_d = None
            if gw is not None:
                gw = np.asarray(gw, dtype=np.float64)
                w_d = torch.as_tensor(np.ascontiguousarray(gw / np.mean(gw)), device=moved.device)
            vals_g = torch.empty((moved.shape[0],), dtype=torch.float64, device=moved.device)
            check(load().mdagpu_group_rmsd(tensor_ptr(moved), moved.shape[0], moved.shape[1], tensor_ptr(idx), len(gm),
                                           tensor_ptr(refpos), tensor_ptr(w_d), tensor_ptr(vals_g), DEVICE,
                                           current_stream(moved)))
            out[:, 3 + k] = vals_g.cpu().numpy()
