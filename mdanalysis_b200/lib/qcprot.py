"""QCP RMSD / optimal rotation -- mirror of ``MDAnalysis/lib/qcprot.pyx``.

``CalcRMSDRotationalMatrix(ref, conf, N, rot, weights)`` (qcprot.pyx:282-320)
keeps its call signature: centred ``(N, 3)`` structures, ``rot`` a flat
9-element array filled in place (or ``None``), ``weights`` ``(N,)`` or ``None``;
returns the RMSD.  ``R = rot.reshape(3, 3)`` satisfies ``conf @ R ~ ref``.

float32 inputs are accepted like in the reference; the GPU accumulates in
float64 in both cases (the reference accumulates in the input precision and
itself documents float32 errors of ~1e-3, rms.py:690-692), so float32 results
agree with the reference's float32 ones to about 1e-4 relative and with its
float64 ones to 1e-12.

``qcp_trajectory`` is the batched twin driving ``RMSD`` / ``AlignTraj``.
"""
import ctypes as C

import numpy as np

from .._cabi import DEVICE, HOST, check, current_stream, is_tensor, load, ptr, tensor_ptr


def CalcRMSDRotationalMatrix(ref, conf, N, rot, weights):
    """RMSD after optimal superposition and (optionally) the rotation matrix."""
    if ref is None or conf is None:
        raise TypeError("Argument 'ref'/'conf' must not be None")
    ref64 = np.ascontiguousarray(ref, dtype=np.float64)
    conf64 = np.ascontiguousarray(conf, dtype=np.float64)
    if ref64.ndim != 2 or conf64.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2)")
    N = int(N)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    r = None if rot is None else np.zeros(9, dtype=np.float64)
    val = C.c_double(0.0)
    check(load().mdagpu_qcp(ptr(ref64), ptr(conf64), N, ptr(r), ptr(w), C.byref(val), HOST, None))
    if rot is not None:
        rot[:] = r
    return float(val.value)


def InnerProduct(A, coords1, coords2, N, weight):
    """``A[9]`` (filled in place) ``= sum_i w_i coords1_i (x) coords2_i`` and the return value
    ``E0 = (sum w |coords1|^2 + sum w |coords2|^2) / 2`` -- qcprot.pyx:179-277.  The first ``N``
    atoms of the two centred ``(n, 3)`` structures are used; float32 or float64 like the
    reference (accumulated in float64 on the GPU)."""
    if A is None or coords1 is None or coords2 is None:
        raise TypeError("Argument 'A'/'coords1'/'coords2' must not be None")
    N = int(N)
    c1 = np.ascontiguousarray(np.asarray(coords1)[:N], dtype=np.float64)
    c2 = np.ascontiguousarray(np.asarray(coords2)[:N], dtype=np.float64)
    if c1.ndim != 2 or c2.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2)")
    w = None if weight is None else np.ascontiguousarray(np.asarray(weight)[:N], dtype=np.float64)
    a = np.zeros(9, dtype=np.float64)
    e0 = C.c_double(0.0)
    check(load().mdagpu_qcp_inner_product(ptr(c1), ptr(c2), N, ptr(w), ptr(a), C.byref(e0), HOST, None))
    A[:] = a
    return float(e0.value)


def FastCalcRMSDAndRotation(rot, A, E0, N):
    """RMSD (returned) and rotation matrix (``rot[9]`` filled in place unless ``None``) from the
    inner-product matrix ``A[9]``, ``E0`` and the atom count -- qcprot.pyx:326-543."""
    if A is None:
        raise TypeError("Argument 'A' must not be None")
    a = np.ascontiguousarray(A, dtype=np.float64).reshape(-1)
    if a.shape != (9,):
        raise ValueError("A must have 9 elements")
    r = None if rot is None else np.zeros(9, dtype=np.float64)
    val = C.c_double(0.0)
    check(load().mdagpu_qcp_solve(ptr(a), float(E0), float(N), ptr(r), C.byref(val)))
    if rot is not None:
        rot[:] = r
    return float(val.value)


def qcp_trajectory(traj, reference, select=None, weights=None, want_rotation=False,
                   align=False, convention=0, allclose_shortcut=False, out=None):
    """Per-frame centring + QCP for a whole trajectory block in one launch.

    traj       float32 [F, natoms, 3] NumPy array or torch CUDA tensor
    reference  float32 [nsel, 3] un-centred reference positions of the fitted atoms
    select     int indices of the fitted atoms within a frame (None: all atoms)
    weights    relative weights [nsel] (already divided by their mean) or None
    align      also move every atom of every frame onto the reference
               (``out`` = destination with traj's shape; ``out is traj`` = in place)
    convention 0: RMSD (rot: mobile @ R ~ ref, rms.py:699-711)
               1: AlignTraj (rot = rotation_matrix(mobile, ref), applied as x @ R.T)
    Returns (rmsd [F], rot [F,3,3] or None, aligned or None) of the input kind.
    """
    dev = is_tensor(traj)
    F, natoms = int(traj.shape[0]), int(traj.shape[1])
    if tuple(traj.shape[2:]) != (3,):
        raise ValueError(f"traj must have shape (F, n, 3), got {tuple(traj.shape)}")
    sel = None if select is None else np.ascontiguousarray(select, dtype=np.int32)
    nsel = natoms if sel is None else len(sel)
    if sel is not None and nsel and (sel.min() < 0 or sel.max() >= natoms):
        raise IndexError("selection index out of range")
    ref = np.ascontiguousarray(reference, dtype=np.float32)
    if ref.shape != (nsel, 3):
        raise ValueError(f"reference must have shape ({nsel}, 3), got {ref.shape}")
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    if w is not None and w.shape != (nsel,):
        raise ValueError("weights must have the same length as the selection")
    lib = load()
    if dev:
        import torch
        t = traj if (traj.dtype == torch.float32 and traj.is_contiguous()) else traj.to(torch.float32).contiguous()
        d = t.device
        rmsd = torch.empty((F,), dtype=torch.float64, device=d)
        rot = torch.empty((F, 3, 3), dtype=torch.float64, device=d) if want_rotation else None
        dst = None
        if align:
            dst = t if out is traj else (out if out is not None else torch.empty_like(t))
        sel_d = None if sel is None else torch.as_tensor(sel, device=d)
        ref_d = torch.as_tensor(ref, device=d)
        w_d = None if w is None else torch.as_tensor(w, device=d)
        check(lib.mdagpu_qcp_trajectory(tensor_ptr(t), F, natoms, tensor_ptr(sel_d), nsel,
                                        tensor_ptr(ref_d), tensor_ptr(w_d), tensor_ptr(rmsd),
                                        tensor_ptr(rot), tensor_ptr(dst), int(convention),
                                        int(bool(allclose_shortcut)), DEVICE, current_stream(t)))
        return rmsd, rot, dst
    t = np.ascontiguousarray(traj, dtype=np.float32)
    rmsd = np.empty((F,), dtype=np.float64)
    rot = np.empty((F, 3, 3), dtype=np.float64) if want_rotation else None
    dst = None
    if align:
        if out is traj and t is traj:
            dst = t
        else:
            dst = out if out is not None else np.empty_like(t)
    check(lib.mdagpu_qcp_trajectory(ptr(t), F, natoms, ptr(sel), nsel, ptr(ref), ptr(w), ptr(rmsd),
                                    ptr(rot), ptr(dst), int(convention),
                                    int(bool(allclose_shortcut)), HOST, None))
    return rmsd, rot, dst


def pairwise_rmsd(traj, select=None, weights=None, superposition=False, cutoff=-4.0):
    """RMSD between every pair of frames in one launch (``mdagpu_pairwise_rmsd``):
    ``rms.rmsd(traj[i, select], traj[j, select], weights, superposition=superposition)`` for all
    i, j -- the loop of ``diffusionmap.DistanceMatrix._single_frame`` (diffusionmap.py:248-267).

    traj     float32 [F, natoms, 3] NumPy array or torch CUDA tensor
    weights  any non-negative weights over the selected atoms (normalised here by their mean,
             rms.py:270) or None
    Returns the symmetric float64 [F, F] matrix, of the input kind."""
    dev = is_tensor(traj)
    F, natoms = int(traj.shape[0]), int(traj.shape[1])
    if tuple(traj.shape[2:]) != (3,):
        raise ValueError(f"traj must have shape (F, n, 3), got {tuple(traj.shape)}")
    sel = None if select is None else np.ascontiguousarray(select, dtype=np.int32)
    nsel = natoms if sel is None else len(sel)
    if sel is not None and nsel and (sel.min() < 0 or sel.max() >= natoms):
        raise IndexError("selection index out of range")
    w = None
    if weights is not None:
        w = np.asarray(weights, dtype=np.float64)
        if w.shape != (nsel,):
            raise ValueError("weights must have same length as a and b")
        w = np.ascontiguousarray(w / np.mean(w))
    lib = load()
    if dev:
        import torch
        t = traj if (traj.dtype == torch.float32 and traj.is_contiguous()) else traj.to(torch.float32).contiguous()
        out = torch.zeros((F, F), dtype=torch.float64, device=t.device)
        sel_d = None if sel is None else torch.as_tensor(sel, device=t.device)
        w_d = None if w is None else torch.as_tensor(w, device=t.device)
        check(lib.mdagpu_pairwise_rmsd(tensor_ptr(t), F, natoms, tensor_ptr(sel_d), nsel, tensor_ptr(w_d),
                                       int(bool(superposition)), float(cutoff), tensor_ptr(out), DEVICE,
                                       current_stream(t)))
        return out
    t = np.ascontiguousarray(traj, dtype=np.float32)
    out = np.zeros((F, F), dtype=np.float64)
    check(lib.mdagpu_pairwise_rmsd(ptr(t), F, natoms, ptr(sel), nsel, ptr(w), int(bool(superposition)),
                                   float(cutoff), ptr(out), HOST, None))
    return out
