"""Contacts on the GPU -- the pieces of ``MDAnalysis/analysis/contacts.py`` that sit
on the hot path: ``contact_matrix`` (:315-338), ``hard_cut_q`` (:268-297) and a
fused ``self_distance_array -> d <= radius`` reduction that never materialises
the n(n-1)/2 distances (config "contacts on a 100k-atom
selection").
"""
import functools

import numpy as np

from .._cabi import (BOX_NONE, BOX_ORTHO, BOX_TRICLINIC, DEVICE, HOST, check, current_stream, is_tensor,
                     load, ptr, tensor_ptr, upload)
from ..frames import FrameBlock, as_block
from ..lib.distances import _box_args, distance_array
from ..lib.util import check_box
from .base import BlockAnalysis
from .rdf import sqrt_le_threshold


def contact_matrix(d, radius, out=None):
    """Boolean contacts ``d <= radius`` from a distance matrix (NumPy or torch)."""
    if out is not None:
        out[:] = d <= radius
    else:
        out = d <= radius
    return out


def hard_cut_q(r, cutoff):
    """Fraction of distances ``r`` within ``cutoff`` (contacts.py:268-297)."""
    r = np.asarray(r)
    cutoff = np.asarray(cutoff)
    y = r <= cutoff
    return y.sum() / r.size


def self_contacts(positions, radius, box=None, return_degree=False):
    """Number of atom pairs (i < j) with ``self_distance_array(x, box) <= radius``, per frame.

    positions  [F, n, 3] / [n, 3] float32 NumPy array or torch CUDA tensor (or FrameBlock)
    box        None or ``[lx, ly, lz, alpha, beta, gamma]`` (one cell for all frames)
    Returns counts uint64 [F] (and per-atom contact degree int32 [F, n]).
    The decision is taken on the reference's float64 distance, so counts equal
    ``(self_distance_array(x, box) <= radius).sum()`` exactly.
    """
    blk = as_block(positions)
    F, n = blk.n_frames, blk.n_atoms
    mode, b = _box_args(box)
    d2_max = float(sqrt_le_threshold(radius))
    if blk.on_device:
        import torch
        x = blk.positions
        counts = torch.zeros(F, dtype=torch.int64, device=x.device)
        degree = torch.zeros((F, n), dtype=torch.int32, device=x.device) if return_degree else None
        check(load().mdagpu_contacts_self(tensor_ptr(x), n, F, mode, ptr(b), d2_max, float(radius),
                                          tensor_ptr(counts), tensor_ptr(degree), DEVICE,
                                          current_stream(x)))
    else:
        x = blk.positions
        counts = np.zeros(F, dtype=np.uint64)
        degree = np.zeros((F, n), dtype=np.int32) if return_degree else None
        check(load().mdagpu_contacts_self(ptr(x), n, F, mode, ptr(b), d2_max, float(radius),
                                          ptr(counts), ptr(degree), HOST, None))
    return (counts, degree) if return_degree else counts


def soft_cut_q(r, r0, beta=5.0, lambda_constant=1.8):
    """Soft-cut fraction of native contacts (contacts.py:231-265)."""
    r = np.asarray(r)
    r0 = np.asarray(r0)
    result = 1 / (1 + np.exp(beta * (r - lambda_constant * r0)))
    return result.sum() / len(r0)


def radius_cut_q(r, r0, radius):
    """Contacts closer than ``radius`` (contacts.py:300-312)."""
    return hard_cut_q(r, radius)


def _frame_boxes(dimensions, n_frames):
    """FrameBlock.dimensions -> (box_mode, float32 boxes [F or 1, 3 | 9], per_frame)."""
    if dimensions is None:
        return BOX_NONE, None, 0
    dims = np.asarray(dimensions, dtype=np.float32).reshape(-1, 6)
    kinds = [check_box(d) for d in dims]
    if all(k == "ortho" for k, _ in kinds):
        mode, boxes = BOX_ORTHO, np.stack([b for _, b in kinds])
    else:   # a mixed trajectory runs the triclinic kernel on every frame (same distances)
        from ..lib.mdamath import triclinic_vectors
        mode, boxes = BOX_TRICLINIC, np.stack([triclinic_vectors(d).reshape(9) for d in dims])
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    if len(boxes) > 1 and (boxes == boxes[0]).all():
        boxes = boxes[:1]
    return mode, boxes, int(len(boxes) > 1 and len(boxes) == n_frames)


def pair_list_distances(A, B, idx_i, idx_j, dimensions=None):
    """``distance_array(A[f], B[f], box[f])[idx_i, idx_j]`` for every frame f, float64 [F, P].

    A, B: [F, n, 3] float32 (NumPy or torch CUDA); only the listed elements of each frame's
    distance matrix are evaluated (``analysis/contacts.py:486-499`` takes ``d[initial_contacts]``).
    """
    F, na, nb = A.shape[0], A.shape[1], B.shape[1]
    mode, boxes, per_frame = _frame_boxes(dimensions, F)
    ii = np.ascontiguousarray(idx_i, dtype=np.int64)
    jj = np.ascontiguousarray(idx_j, dtype=np.int64)
    P = len(ii)
    if P and (ii.min() < 0 or ii.max() >= na or jj.min() < 0 or jj.max() >= nb):
        raise IndexError("pair index out of range")
    if is_tensor(A):
        import torch
        Ad, Bd = A.to(torch.float32).contiguous(), B.to(torch.float32).contiguous()
        out = torch.empty((F, P), dtype=torch.float64, device=Ad.device)
        iid, jjd = torch.as_tensor(ii, device=Ad.device), torch.as_tensor(jj, device=Ad.device)
        check(load().mdagpu_pair_list_distances(tensor_ptr(Ad), na, tensor_ptr(Bd), nb, F, tensor_ptr(iid),
                                                tensor_ptr(jjd), P, mode, ptr(boxes), per_frame,
                                                tensor_ptr(out), DEVICE, current_stream(Ad)))
        return out
    Ah = np.ascontiguousarray(A, dtype=np.float32)
    Bh = np.ascontiguousarray(B, dtype=np.float32)
    out = np.empty((F, P), dtype=np.float64)
    check(load().mdagpu_pair_list_distances(ptr(Ah), na, ptr(Bh), nb, F, ptr(ii), ptr(jj), P, mode,
                                            ptr(boxes), per_frame, ptr(out), HOST, None))
    return out


class Contacts(BlockAnalysis):
    """Fraction of native contacts per frame (``analysis/contacts.py:341-499``).

    u         FrameBlock | [F, n, 3] array / CUDA tensor holding every atom
    select    (selA, selB): index arrays into ``u`` (or "all")
    refgroup  (refA, refB) or a list of such tuples: reference positions ``(n, 3)`` of the two
              groups (arrays, or one-frame FrameBlocks whose ``dimensions`` give the box)
    method    'hard_cut' | 'soft_cut' | 'radius_cut' | callable(r, r0, **kwargs)
    The per-frame distances of the native contacts are evaluated for all frames in one
    kernel launch (only the native pairs, not the full matrix); ``results.timeseries`` is
    ``[frame, q_1, ...]`` as in the reference.
    """

    def __init__(self, u, select, refgroup, method="hard_cut", radius=4.5, pbc=True, kwargs=None, **basekwargs):
        blk = as_block(u)
        super().__init__(blk, **basekwargs)
        self.fraction_kwargs = kwargs if kwargs is not None else {}
        self._method_name = method if isinstance(method, str) else None
        if method == "hard_cut":
            self.fraction_contacts = hard_cut_q
        elif method == "soft_cut":
            self.fraction_contacts = soft_cut_q
        elif method == "radius_cut":
            self.fraction_contacts = functools.partial(radius_cut_q, radius=radius)
        else:
            if not callable(method):
                raise ValueError("method has to be callable")
            self.fraction_contacts = method
        self.select = select
        self._idx = []
        for sel in select:
            if isinstance(sel, str):
                if sel != "all":
                    raise TypeError("selection strings need a topology: pass index arrays")
                self._idx.append(np.arange(blk.n_atoms))
            else:
                self._idx.append(np.asarray(sel, dtype=np.int64))
        self.pbc = pbc
        self.radius = radius
        self.r0, self.initial_contacts = [], []
        groups = [refgroup] if not isinstance(refgroup[0], (tuple, list)) else list(refgroup)
        for refA, refB in groups:
            box = None
            if isinstance(refA, FrameBlock):
                if pbc and refA.dimensions is not None:
                    box = refA.dimensions[0]
                refA = refA.positions[0]
            if isinstance(refB, FrameBlock):
                refB = refB.positions[0]
            if is_tensor(refA):
                refA = refA.detach().cpu().numpy()
            if is_tensor(refB):
                refB = refB.detach().cpu().numpy()
            self.r0.append(distance_array(np.asarray(refA, dtype=np.float32), np.asarray(refB, dtype=np.float32),
                                          box=box))
            self.initial_contacts.append(contact_matrix(self.r0[-1], radius))
        self.n_initial_contacts = self.initial_contacts[0].sum()

    def _prepare(self, blk):
        self.results.timeseries = np.empty((blk.n_frames, len(self.r0) + 1))

    def _compute(self, blk):
        ts = self.results.timeseries
        ts[:, 0] = blk.frames
        # the native-pair distances stay on the device; only q[F] comes back (csrc/reduce.cu).  A
        # user-supplied callable (or extra kwargs) is evaluated on the host, frame by frame, as in
        # the reference (contacts.py:486-499).
        import torch
        dev = blk.positions.device if blk.on_device else torch.device("cuda", torch.cuda.current_device())
        A = blk.select(self._idx[0]).positions
        B = blk.select(self._idx[1]).positions
        if not blk.on_device:
            A, B = upload(A, dev), upload(B, dev)
        dims = blk.dimensions if self.pbc else None
        kw = self.fraction_kwargs
        builtin = {"hard_cut": 0, "radius_cut": 1, "soft_cut": 2}.get(self._method_name)
        if builtin == 2:
            if set(kw) - {"beta", "lambda_constant"}:
                builtin = None
        elif builtin is not None and kw:
            builtin = None
        for k, (mask, r0) in enumerate(zip(self.initial_contacts, self.r0), 1):
            ii, jj = np.nonzero(mask)                    # row-major, the order of d[mask]
            r = pair_list_distances(A, B, ii, jj, dims)
            r0s = np.ascontiguousarray(r0[mask], dtype=np.float64)
            if builtin is not None and r0s.size:
                q = torch.empty((blk.n_frames,), dtype=torch.float64, device=dev)
                r0d = torch.as_tensor(r0s, device=dev)
                check(load().mdagpu_contact_fraction(tensor_ptr(r), blk.n_frames, r0s.size, tensor_ptr(r0d), builtin,
                                                     float(self.radius), float(kw.get("beta", 5.0)),
                                                     float(kw.get("lambda_constant", 1.8)), tensor_ptr(q), DEVICE,
                                                     current_stream(r)))
                ts[:, k] = q.cpu().numpy()
                continue
            r = r.detach().cpu().numpy()
            with np.errstate(invalid="ignore", divide="ignore"):
                for f in range(blk.n_frames):
                    ts[f, k] = self.fraction_contacts(r[f], r0s, **self.fraction_kwargs)
