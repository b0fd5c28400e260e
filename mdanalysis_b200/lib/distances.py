"""Front end of the distance functions -- mirror of ``MDAnalysis/lib/distances.py``.

Same function names, argument meaning, return types and error behaviour as the
reference, with ``backend="cuda"`` as the (only) backend of this package:

==========================  =======================================================
function                    reference (package/MDAnalysis/lib/distances.py)
==========================  =======================================================
``distance_array``          :235-321
``self_distance_array``     :324-404
``capped_distance``         :407-513  (+ ``_determine_method`` :516-586,
                            ``_bruteforce_capped`` :589-673, ``_nsgrid_capped`` :768-874)
``self_capped_distance``    :877-982  (+ ``_determine_method_self`` :985-1047,
                            ``_bruteforce_capped_self`` :1050-1130,
                            ``_nsgrid_capped_self`` :1217-1317)
``transform_RtoS/StoR``     :1320-1411
``calc_bonds``              :1414-1515
``calc_angles``             :1518-1617
``calc_dihedrals``          :1620-1735
``apply_PBC``               :1738-1787
==========================  =======================================================

NumPy arrays (or AtomGroup-likes with ``.positions``) go through the reference's
checks and the ``_distances[backend]`` module (:mod:`.c_distances_cuda`), i.e.
exactly what happens inside MDAnalysis once the backend is registered.  In
addition every function accepts **torch CUDA tensors** (float32, contiguous,
``(n, 3)``): the computation then stays on the device, on torch's current
stream, and torch tensors are returned ("tensor twin" of SURVEY.md section 8b).

``capped_distance`` / ``self_capped_distance`` gain a ``backend=`` keyword (the
one API addition; the reference selects by ``method=`` only).  ``method`` keeps
its meaning: it selects *which reference algorithm's arithmetic and pair set* is
reproduced ('bruteforce' or 'nsgrid'; 'pkdtree' is answered with the brute-force
predicate, whose distances it shares).
"""
from __future__ import annotations

import numpy as np

from .. import _cabi
from .._cabi import (BOX_NONE, BOX_ORTHO, BOX_TRICLINIC, DEVICE, HOST, check, current_stream,
                     is_tensor, load, ptr, tensor_ptr)
from . import c_distances_cuda
from .mdamath import triclinic_vectors
from .util import check_box, prepare_coords

# backend registry, same shape as the reference's (distances.py:139-151)
_distances = {"cuda": c_distances_cuda}
_UINT64_MAX = c_distances_cuda._UINT64_MAX


def _run(funcname, args=None, kwargs=None, backend="cuda"):
    """Select ``funcname`` from the backend module (distances.py:154-166)."""
    args = args if args is not None else tuple()
    kwargs = kwargs if kwargs is not None else dict()
    backend = backend.lower()
    try:
        func = getattr(_distances[backend], funcname)
    except KeyError:
        raise ValueError(f"Function {funcname} not available with backend {backend} "
                         f"try one of: {_distances.keys()}") from None
    return func(*args, **kwargs)


def _check_result_array(result, shape):
    """distances.py:193-232."""
    if result is None:
        return np.zeros(shape, dtype=np.float64)
    if result.shape != shape:
        raise ValueError(f"Result array has incorrect shape, should be {shape}, got "
                         f"{result.shape}.")
    if result.dtype != np.float64:
        raise TypeError(f"Result array must be of type numpy.float64, got {result.dtype}.")
    return result


# ---------------------------------------------------------------------------
# device-resident ("tensor twin") helpers
# ---------------------------------------------------------------------------
def _check_backend(backend):
    if backend.lower() not in _distances:
        raise ValueError(f"backend {backend} not available, try one of: {_distances.keys()}")


def _dev_coords(fname, name, t):
    import torch
    if not is_tensor(t) or not t.is_cuda:
        raise TypeError(f"{fname}(): '{name}' must be a torch CUDA tensor when any input is one")
    if t.ndim == 1 and t.shape[0] == 3:
        t = t[None, :]
    if t.ndim != 2 or t.shape[1] != 3:
        raise ValueError(f"{fname}(): {name}.shape must be (3,) or (n, 3), got {tuple(t.shape)}")
    return t.to(torch.float32).contiguous()


def _box_args(box):
    """box -> (mode, float32 host array or None) for the C ABI."""
    if box is None:
        return BOX_NONE, None
    boxtype, b = check_box(box)
    if boxtype == "ortho":
        return BOX_ORTHO, np.ascontiguousarray(b, dtype=np.float32)
    return BOX_TRICLINIC, np.ascontiguousarray(b, dtype=np.float32).reshape(9)


def _dev_result(result, shape, like):
    import torch
    if result is None:
        return torch.empty(shape, dtype=torch.float64, device=like.device)
    if tuple(result.shape) != tuple(shape):
        raise ValueError(f"Result array has incorrect shape, should be {shape}, got "
                         f"{tuple(result.shape)}.")
    if result.dtype != torch.float64:
        raise TypeError(f"Result array must be of type float64, got {result.dtype}.")
    if not result.is_contiguous():
        raise ValueError("a device result tensor must be contiguous")
    return result


# ---------------------------------------------------------------------------
# distance_array / self_distance_array
# ---------------------------------------------------------------------------
def distance_array(reference, configuration, box=None, result=None, backend="cuda"):
    """All distances between ``reference`` (n) and ``configuration`` (m) -> float64 (n, m)."""
    if is_tensor(reference) or is_tensor(configuration):
        _check_backend(backend)
        ref = _dev_coords("distance_array", "reference", reference)
        conf = _dev_coords("distance_array", "configuration", configuration)
        out = _dev_result(result, (ref.shape[0], conf.shape[0]), ref)
        if out.numel():
            mode, b = _box_args(box)
            check(load().mdagpu_distance_array(tensor_ptr(ref), ref.shape[0], tensor_ptr(conf),
                                               conf.shape[0], mode, ptr(b), tensor_ptr(out), 1,
                                               DEVICE, current_stream(ref)))
        return out
    (reference, configuration), _ = prepare_coords(
        "distance_array", [("reference", reference), ("configuration", configuration)],
        reduce_result_if_single=False, check_lengths_match=False, allow_atomgroup=True)
    confnum, refnum = configuration.shape[0], reference.shape[0]
    if refnum * confnum > _UINT64_MAX:
        raise ValueError(f"Size of resulting array {refnum * confnum} elements larger than size "
                         "of maximum integer")
    distances = _check_result_array(result, (refnum, confnum))
    if len(distances) == 0:
        return distances
    if box is not None:
        boxtype, box = check_box(box)
        if boxtype == "ortho":
            _run("calc_distance_array_ortho", args=(reference, configuration, box, distances),
                 backend=backend)
        else:
            _run("calc_distance_array_triclinic", args=(reference, configuration, box, distances),
                 backend=backend)
    else:
        _run("calc_distance_array", args=(reference, configuration, distances), backend=backend)
    return distances


def self_distance_array(reference, box=None, result=None, backend="cuda"):
    """All distances within ``reference`` -> float64 (n*(n-1)/2,), packed upper triangle."""
    if is_tensor(reference):
        _check_backend(backend)
        ref = _dev_coords("self_distance_array", "reference", reference)
        n = ref.shape[0]
        out = _dev_result(result, (n * (n - 1) // 2,), ref)
        if out.numel():
            mode, b = _box_args(box)
            check(load().mdagpu_self_distance_array(tensor_ptr(ref), n, mode, ptr(b),
                                                    tensor_ptr(out), 1, DEVICE, current_stream(ref)))
        return out
    (reference,), _ = prepare_coords("self_distance_array", [("reference", reference)],
                                     reduce_result_if_single=False, allow_atomgroup=True)
    refnum = reference.shape[0]
    distnum = refnum * (refnum - 1) // 2
    if distnum > _UINT64_MAX:
        raise ValueError(f"Size of resulting array {distnum} elements larger than size of "
                         "maximum integer")
    distances = _check_result_array(result, (distnum,))
    if len(distances) == 0:
        return distances
    if box is not None:
        boxtype, box = check_box(box)
        if boxtype == "ortho":
            _run("calc_self_distance_array_ortho", args=(reference, box, distances), backend=backend)
        else:
            _run("calc_self_distance_array_triclinic", args=(reference, box, distances),
                 backend=backend)
    else:
        _run("calc_self_distance_array", args=(reference, distances), backend=backend)
    return distances


# ---------------------------------------------------------------------------
# capped distances
# ---------------------------------------------------------------------------
def _determine_method(reference, configuration, max_cutoff, min_cutoff=None, box=None,
                      method=None):
    """Which reference algorithm would run (distances.py:516-586) -> its name."""
    if method is not None:
        method = method.lower()
        if method not in ("bruteforce", "pkdtree", "nsgrid"):
            raise KeyError(method)
        return method
    if len(reference) < 10 or len(configuration) < 10:
        return "bruteforce"
    if len(reference) * len(configuration) >= 1e8:
        return "nsgrid"
    if box is None:
        min_dim = np.array([reference.min(axis=0), configuration.min(axis=0)])
        max_dim = np.array([reference.max(axis=0), configuration.max(axis=0)])
        size = max_dim.max(axis=0) - min_dim.min(axis=0)
    elif np.all(box[3:] == 90.0):
        size = box[:3]
    else:
        tribox = triclinic_vectors(box)
        size = tribox.max(axis=0) - tribox.min(axis=0)
    if np.any(max_cutoff > 0.3 * size):
        return "bruteforce"
    return "nsgrid"


def _determine_method_self(reference, max_cutoff, min_cutoff=None, box=None, method=None):
    """distances.py:985-1047."""
    if method is not None:
        method = method.lower()
        if method not in ("bruteforce", "pkdtree", "nsgrid"):
            raise KeyError(method)
        return method
    if len(reference) < 100:
        return "bruteforce"
    if box is None:
        size = reference.max(axis=0) - reference.min(axis=0)
    elif np.all(box[3:] == 90.0):
        size = box[:3]
    else:
        tribox = triclinic_vectors(box)
        size = tribox.max(axis=0) - tribox.min(axis=0)
    if max_cutoff < 0.03 * size.min():
        return "pkdtree"
    return "nsgrid"


def _export_pairs(handle, return_distances=True):
    lib = load()
    n = int(lib.mdagpu_pairs_count(handle))
    pairs = np.empty((n, 2), dtype=np.intp)
    dist = np.empty((n,), dtype=np.float64) if return_distances else None
    try:
        if n:
            check(lib.mdagpu_pairs_export(handle, ptr(pairs), ptr(dist), HOST, None))
    finally:
        lib.mdagpu_pairs_free(handle)
    return pairs, dist


def _bruteforce_capped(reference, configuration, max_cutoff, min_cutoff=None, box=None,
                       return_distances=True):
    """Pairs with ``min_cutoff < d <= max_cutoff``, d as in distance_array (:589-673)."""
    import ctypes as C
    pairs = np.empty((0, 2), dtype=np.intp)
    distances = np.empty((0,), dtype=np.float64)
    if len(reference) > 0 and len(configuration) > 0:
        mode, b = _box_args(box)
        h = C.c_void_p()
        check(load().mdagpu_capped_bruteforce(ptr(reference), len(reference), ptr(configuration),
                                              len(configuration), float(max_cutoff),
                                              -1.0 if min_cutoff is None else float(min_cutoff),
                                              mode, ptr(b), HOST, None, C.byref(h)))
        pairs, distances = _export_pairs(h)
    return (pairs, distances) if return_distances else pairs


def _bruteforce_capped_self(reference, max_cutoff, min_cutoff=None, box=None,
                            return_distances=True):
    """distances.py:1050-1130 (pairs oriented i < j)."""
    import ctypes as C
    pairs = np.empty((0, 2), dtype=np.intp)
    distances = np.empty((0,), dtype=np.float64)
    if len(reference) > 1:
        mode, b = _box_args(box)
        h = C.c_void_p()
        check(load().mdagpu_capped_bruteforce(ptr(reference), len(reference), None, 0,
                                              float(max_cutoff),
                                              -1.0 if min_cutoff is None else float(min_cutoff),
                                              mode, ptr(b), HOST, None, C.byref(h)))
        pairs, distances = _export_pairs(h)
    return (pairs, distances) if return_distances else pairs


def _nsgrid_capped(reference, configuration, max_cutoff, min_cutoff=None, box=None,
                   return_distances=True):
    """FastNS-based search (distances.py:768-874)."""
    from .nsgrid import FastNS
    pairs = np.empty((0, 2), dtype=np.intp)
    distances = np.empty((0,), dtype=np.float64)
    if len(reference) > 0 and len(configuration) > 0:
        if box is None:
            # pseudo box around all coordinates, shifted to the origin (:838-859)
            pseudobox = np.zeros(6, dtype=np.float32)
            all_coords = np.concatenate([reference, configuration])
            lmax = all_coords.max(axis=0)
            lmin = all_coords.min(axis=0)
            boxsize = (lmax - lmin).max()
            boxsize = np.maximum(boxsize, 2 * max_cutoff)
            pseudobox[:3] = boxsize + 2.2 * max_cutoff
            pseudobox[3:] = 90.0
            shiftref, shiftconf = reference.copy(), configuration.copy()
            shiftref -= lmin - 0.1 * max_cutoff
            shiftconf -= lmin - 0.1 * max_cutoff
            gridsearch = FastNS(max_cutoff, shiftconf, box=pseudobox, pbc=False)
            results = gridsearch.search(shiftref)
        else:
            gridsearch = FastNS(max_cutoff, configuration, box=box)
            results = gridsearch.search(reference)
        pairs = results.get_pairs()
        if return_distances or (min_cutoff is not None):
            distances = results.get_pair_distances()
            if min_cutoff is not None:
                idx = distances > min_cutoff
                pairs, distances = pairs[idx], distances[idx]
    return (pairs, distances) if return_distances else pairs


def _nsgrid_capped_self(reference, max_cutoff, min_cutoff=None, box=None, return_distances=True):
    """distances.py:1217-1317."""
    from .nsgrid import FastNS
    pairs = np.empty((0, 2), dtype=np.intp)
    distances = np.empty((0,), dtype=np.float64)
    if len(reference) > 1:
        if box is None:
            pseudobox = np.zeros(6, dtype=np.float32)
            lmax = reference.max(axis=0)
            lmin = reference.min(axis=0)
            boxsize = (lmax - lmin).max()
            if boxsize < 2 * max_cutoff:
                sizefactor = 2.2 * max_cutoff / boxsize
            else:
                sizefactor = 1.2
            pseudobox[:3] = sizefactor * boxsize
            pseudobox[3:] = 90.0
            shiftref = reference.copy()
            shiftref -= lmin - 0.1 * boxsize
            gridsearch = FastNS(max_cutoff, shiftref, box=pseudobox, pbc=False)
        else:
            gridsearch = FastNS(max_cutoff, reference, box=box)
        results = gridsearch.self_search()
        pairs = results.get_pairs()
        if return_distances or (min_cutoff is not None):
            distances = results.get_pair_distances()
            if min_cutoff is not None:
                idx = distances > min_cutoff
                pairs, distances = pairs[idx], distances[idx]
    return (pairs, distances) if return_distances else pairs


def _as_host_coords(fname, name, x):
    """capped distances take host arrays; device tensors are brought over once."""
    if is_tensor(x):
        x = x.detach().to("cpu").numpy()
    (x,), _ = prepare_coords(fname, [(name, x)], enforce_copy=False, reduce_result_if_single=False,
                             check_lengths_match=False, allow_atomgroup=True)
    return x


def capped_distance(reference, configuration, max_cutoff, min_cutoff=None, box=None, method=None,
                    return_distances=True, backend="cuda"):
    """Pairs (i, j) with ``min_cutoff < |reference[i] - configuration[j]| <= max_cutoff``.

    Returns ``pairs`` intp (n, 2) and, optionally, ``distances`` float64 (n,).
    Row order is unspecified (as in the reference).
    """
    _check_backend(backend)
    reference = _as_host_coords("capped_distance", "reference", reference)
    configuration = _as_host_coords("capped_distance", "configuration", configuration)
    if box is not None:
        box = np.asarray(box, dtype=np.float32)
        if box.shape[0] != 6:
            raise ValueError("Box Argument is of incompatible type. The dimension should be "
                             "either None or of the form [lx, ly, lz, alpha, beta, gamma]")
    method = _determine_method(reference, configuration, max_cutoff, min_cutoff=min_cutoff,
                               box=box, method=method)
    function = _nsgrid_capped if method == "nsgrid" else _bruteforce_capped
    return function(reference, configuration, max_cutoff, min_cutoff=min_cutoff, box=box,
                    return_distances=return_distances)


def self_capped_distance(reference, max_cutoff, min_cutoff=None, box=None, method=None,
                         return_distances=True, backend="cuda"):
    """Pairs (i, j), each once, with ``min_cutoff < |r_i - r_j| <= max_cutoff``."""
    _check_backend(backend)
    reference = _as_host_coords("self_capped_distance", "reference", reference)
    if box is not None:
        box = np.asarray(box, dtype=np.float32)
        if box.shape[0] != 6:
            raise ValueError("Box Argument is of incompatible type. The dimension should be "
                             "either None or of the form [lx, ly, lz, alpha, beta, gamma]")
    method = _determine_method_self(reference, max_cutoff, min_cutoff=min_cutoff, box=box,
                                    method=method)
    function = _nsgrid_capped_self if method == "nsgrid" else _bruteforce_capped_self
    return function(reference, max_cutoff, min_cutoff=min_cutoff, box=box,
                    return_distances=return_distances)


# ---------------------------------------------------------------------------
# transforms, bonds, angles, dihedrals, wrapping
# ---------------------------------------------------------------------------
def _transform(fname, coords, box, backend, inverse):
    dev = is_tensor(coords)
    if dev:
        _check_backend(backend)
        single = coords.ndim == 1
        work = _dev_coords(fname, "coords", coords).clone()
    else:
        (work,), single = prepare_coords(fname, [("coords", coords)])
    if len(work) == 0:
        return work
    boxtype, b = check_box(box)
    if boxtype == "ortho":
        b = np.diag(b)
    b = b.astype(np.float64)
    if inverse:
        b = np.array(np.linalg.inv(b), order="C")
    if dev:
        m = np.ascontiguousarray(b, dtype=np.float64)
        check(load().mdagpu_coord_transform(tensor_ptr(work), work.shape[0], ptr(m), DEVICE,
                                            current_stream(work)))
    else:
        _run("coord_transform", args=(work, b), backend=backend)
    return work[0] if single else work


def transform_RtoS(coords, box, backend="cuda"):
    """Real space -> fractional ("S") space: ``coords . inv(box)``."""
    return _transform("transform_RtoS", coords, box, backend, True)


def transform_StoR(coords, box, backend="cuda"):
    """Fractional -> real space: ``coords . box``."""
    return _transform("transform_StoR", coords, box, backend, False)


_TUPLE_FN = {2: ("calc_bond_distance", "mdagpu_calc_bonds"),
             3: ("calc_angle", "mdagpu_calc_angles"),
             4: ("calc_dihedral", "mdagpu_calc_dihedrals")}


def _tuple_front(fname, names, coords, box, result, backend):
    shim, cfn = _TUPLE_FN[len(coords)]
    if any(is_tensor(c) for c in coords):
        _check_backend(backend)
        single = all(c.ndim == 1 for c in coords)
        arrs = [_dev_coords(fname, n, c) for n, c in zip(names, coords)]
        if len({a.shape[0] for a in arrs}) != 1:
            raise ValueError(f"{fname}(): {', '.join(names)} must contain the same number of "
                             f"coordinates, got {[a.shape[0] for a in arrs]}.")
        n = arrs[0].shape[0]
        out = _dev_result(result, (n,), arrs[0])
        if n:
            mode, b = _box_args(box)
            check(getattr(load(), cfn)(*[tensor_ptr(a) for a in arrs], n, mode, ptr(b),
                                       tensor_ptr(out), DEVICE, current_stream(arrs[0])))
        return out[0] if single else out
    arrs, single = prepare_coords(fname, list(zip(names, coords)), allow_atomgroup=True)
    n = arrs[0].shape[0]
    out = _check_result_array(result, (n,))
    if n > 0:
        if box is not None:
            boxtype, b = check_box(box)
            suffix = "_ortho" if boxtype == "ortho" else "_triclinic"
            _run(shim + suffix, args=(*arrs, b, out), backend=backend)
        else:
            _run(shim, args=(*arrs, out), backend=backend)
    return out[0] if single else out


def calc_bonds(coords1, coords2, box=None, result=None, backend="cuda"):
    """``|coords1[i] - coords2[i]|`` with optional minimum image."""
    return _tuple_front("calc_bonds", ("coords1", "coords2"), (coords1, coords2), box, result,
                        backend)


def calc_angles(coords1, coords2, coords3, box=None, result=None, backend="cuda"):
    """Angle (radians) at ``coords2[i]`` between ``coords1[i]`` and ``coords3[i]``."""
    return _tuple_front("calc_angles", ("coords1", "coords2", "coords3"),
                        (coords1, coords2, coords3), box, result, backend)


def calc_dihedrals(coords1, coords2, coords3, coords4, box=None, result=None, backend="cuda"):
    """Dihedral (radians) of the four points; NaN for collinear triples."""
    return _tuple_front("calc_dihedrals", ("coords1", "coords2", "coords3", "coords4"),
                        (coords1, coords2, coords3, coords4), box, result, backend)


def apply_PBC(coords, box, backend="cuda"):
    """Move coordinates into the primary unit cell (a wrapped copy is returned)."""
    if is_tensor(coords):
        _check_backend(backend)
        single = coords.ndim == 1
        work = _dev_coords("apply_PBC", "coords", coords).clone()
        if len(work):
            mode, b = _box_args(box)
            check(load().mdagpu_wrap(tensor_ptr(work), work.shape[0], mode, ptr(b), DEVICE,
                                     current_stream(work)))
        return work[0] if single else work
    (work,), single = prepare_coords("apply_PBC", [("coords", coords)], allow_atomgroup=True)
    if len(work) == 0:
        return work[0] if single else work
    boxtype, b = check_box(box)
    if boxtype == "ortho":
        _run("ortho_pbc", args=(work, b), backend=backend)
    else:
        _run("triclinic_pbc", args=(work, b), backend=backend)
    return work[0] if single else work


def minimize_vectors(vectors, box):
    """Apply the minimum image convention to an array of vectors (lib/distances.py:1790-1830).

    ``vectors`` is ``(n, 3)`` float32 or float64 (NumPy array or torch CUDA tensor); the
    result has the same dtype, as in the reference (``check_coords(enforce_copy=False,
    enforce_dtype=False)``).  The box is converted to the vectors' dtype first (:1822).
    """
    boxtype, b = check_box(box)
    if is_tensor(vectors):
        import torch
        if vectors.ndim != 2 or vectors.shape[1] != 3:
            raise ValueError(f"minimize_vectors(): vectors.shape must be (n, 3) got {tuple(vectors.shape)}")
        if vectors.dtype not in (torch.float32, torch.float64):
            raise TypeError("minimize_vectors(): vectors must be float32 or float64")
        v = vectors.contiguous()
        out = torch.empty_like(v)
        bb = np.ascontiguousarray(b.astype(np.float64 if v.dtype == torch.float64 else np.float32).ravel())
        check(load().mdagpu_minimize_vectors(tensor_ptr(v), v.shape[0], int(v.dtype == torch.float64),
                                             BOX_ORTHO if boxtype == "ortho" else BOX_TRICLINIC, ptr(bb),
                                             tensor_ptr(out), DEVICE, current_stream(v)))
        return out
    (v,), _ = prepare_coords("minimize_vectors", [("vectors", vectors)], enforce_copy=False, enforce_dtype=False,
                             allow_single=False)
    if v.dtype not in (np.float32, np.float64):
        # the reference's fused-type kernels accept nothing else (c_distances.pyx:372, 394)
        raise TypeError(f"minimize_vectors(): no kernel for dtype {v.dtype}")
    vc = np.ascontiguousarray(v)
    out = np.empty_like(vc)
    bb = np.ascontiguousarray(b.astype(vc.dtype).ravel())
    check(load().mdagpu_minimize_vectors(ptr(vc), vc.shape[0], int(vc.dtype == np.float64),
                                         BOX_ORTHO if boxtype == "ortho" else BOX_TRICLINIC, ptr(bb),
                                         ptr(out), HOST, None))
    return out
