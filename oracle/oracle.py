"""CPU oracle for the MDAnalysis geometry hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``mdanalysis_b200/`` imports this module.  It may be used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and only as the checker / the timed CPU baseline.

Two back ends are exposed through the same Python functions:

* ``impl="port"``      -> ``oracle/liboracle.so``: our plain-C restatement
  (``oracle/oracle.c``) of the reference arithmetic.
* ``impl="reference"`` -> ``oracle/_ref/libmda_ref.so`` (or ``_omp``): the
  reference's own ``calc_distances.h`` compiled through ``oracle/ref_shim.c``.
  Covers the ``lib.distances`` kernels only; FastNS / QCP are Cython in the
  reference and therefore always come from the port (which is pinned against
  the reference by ``oracle/check_against_reference.py`` and ``tests/golden``).

Parity status: pinned (see header of oracle.c).

Python-level logic restated here follows (paths relative to
/root/reference/package/MDAnalysis/):
  lib/mdamath.py:309-444      triclinic_vectors, box_volume
  lib/util.py:2471-2524       check_box
  lib/distances.py:407-1317   capped_distance / self_capped_distance front ends
  analysis/rdf.py:253-316     InterRDF
  analysis/rms.py:178-283,620-727   rmsd, RMSD
  analysis/align.py:224-366,836-854 rotation_matrix, _fit_to, AlignTraj
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_F32P = C.POINTER(C.c_float)
_F64P = C.POINTER(C.c_double)
_I32P = C.POINTER(C.c_int)
_I64P = C.POINTER(C.c_int64)

BOX_NONE, BOX_ORTHO, BOX_TRICLINIC = 0, 1, 2


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when the reference checkout exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref = os.path.join(_HERE, "_ref", "libmda_ref.so")
    if os.path.isdir("/root/reference/package") and (force or not os.path.exists(ref)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_libs: dict = {}


def _port():
    if "port" not in _libs:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_grid_create.restype = C.c_void_p
        L.orc_pairs_new.restype = C.c_void_p
        L.orc_pairs_count.restype = C.c_uint64
        L.orc_grid_max_cutoff.restype = C.c_float
        for name in ("orc_inner_product_f64", "orc_qcp_solve", "orc_qcp_f64", "orc_qcp_f32"):
            getattr(L, name).restype = C.c_double
        L.orc_inner_product_f32.restype = C.c_float
        _libs["port"] = L
    return _libs["port"]


def have_reference(openmp: bool = False) -> bool:
    name = "libmda_ref_omp.so" if openmp else "libmda_ref.so"
    return os.path.exists(os.path.join(_HERE, "_ref", name))


def _ref(openmp: bool = False):
    key = "ref_omp" if openmp else "ref"
    if key not in _libs:
        name = "libmda_ref_omp.so" if openmp else "libmda_ref.so"
        _libs[key] = C.CDLL(os.path.join(_HERE, "_ref", name))
    return _libs[key]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------
# box handling (host-side NumPy in the reference)
# --------------------------------------------------------------------------
def triclinic_vectors(dimensions, dtype=np.float32):
    """lib/mdamath.py:309-405 -- [lx,ly,lz,a,b,g] -> lower-triangular 3x3."""
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim
    if not (np.all(dim > 0.0) and alpha < 180.0 and beta < 180.0 and gamma < 180.0):
        return np.zeros((3, 3), dtype=dtype)
    if alpha == beta == gamma == 90.0:
        return np.diag(dim[:3].astype(dtype, copy=False))
    m = np.zeros((3, 3), dtype=np.float64)
    m[0, 0] = lx
    cos_a = 0.0 if alpha == 90.0 else np.cos(np.deg2rad(alpha))
    cos_b = 0.0 if beta == 90.0 else np.cos(np.deg2rad(beta))
    if gamma == 90.0:
        cos_g, sin_g = 0.0, 1.0
    else:
        g = np.deg2rad(gamma)
        cos_g, sin_g = np.cos(g), np.sin(g)
    m[1, 0] = ly * cos_g
    m[1, 1] = ly * sin_g
    m[2, 0] = lz * cos_b
    m[2, 1] = lz * (cos_a - cos_b * cos_g) / sin_g
    m[2, 2] = np.sqrt(lz * lz - m[2, 0] ** 2 - m[2, 1] ** 2)
    if m[2, 2] > 0.0:
        return m.astype(dtype, copy=False)
    return np.zeros((3, 3), dtype=dtype)


def box_volume(dimensions) -> float:
    """lib/mdamath.py:408-444."""
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim
    if alpha == beta == gamma == 90.0 and lx > 0 and ly > 0 and lz > 0:
        return lx * ly * lz
    t = triclinic_vectors(dim, dtype=np.float64)
    return t[0, 0] * t[1, 1] * t[2, 2]


def check_box(box):
    """lib/util.py:2515-2524 -> (mode, float32 box for the C layer)."""
    if box is None:
        raise ValueError("Box is None")
    box = np.asarray(box, dtype=np.float32, order="C")
    if box.shape != (6,):
        raise ValueError("Invalid box information. Must be of the form "
                         "[lx, ly, lz, alpha, beta, gamma].")
    if np.all(box[3:] == 90.0):
        return BOX_ORTHO, np.ascontiguousarray(box[:3])
    return BOX_TRICLINIC, np.ascontiguousarray(triclinic_vectors(box).reshape(9))


def _mode_box(box):
    if box is None:
        return BOX_NONE, np.zeros(9, dtype=np.float32)
    return check_box(box)


# --------------------------------------------------------------------------
# lib.distances kernels
# --------------------------------------------------------------------------
def _dist_lib(impl, openmp=False):
    if impl == "reference":
        return _ref(openmp), "ref_"
    return _port(), "orc_"


def apply_PBC(coords, box, impl="port"):
    """lib/distances.py:1738-1787."""
    xyz = _f32(coords).copy()
    if xyz.size == 0:
        return xyz
    mode, b = check_box(box)
    L, pre = _dist_lib(impl)
    if impl == "reference":
        fn = L.ref_ortho_pbc if mode == BOX_ORTHO else L.ref_triclinic_pbc
    else:
        fn = L.orc_wrap_ortho if mode == BOX_ORTHO else L.orc_wrap_triclinic
    fn(_p(xyz, _F32P), C.c_uint64(len(xyz)), _p(b, _F32P))
    return xyz


def distance_array(ref, conf, box=None, impl="port", openmp=False):
    """lib/distances.py:235-321 -> include/calc_distances.h:381-446."""
    r, c = _f32(ref).copy(), _f32(conf).copy()
    out = np.zeros((len(r), len(c)), dtype=np.float64)
    if out.size == 0:
        return out
    mode, b = _mode_box(box)
    L, pre = _dist_lib(impl, openmp)
    getattr(L, pre + "distance_array")(_p(r, _F32P), C.c_uint64(len(r)), _p(c, _F32P),
                                       C.c_uint64(len(c)), mode, _p(b, _F32P), _p(out, _F64P))
    return out


def self_distance_array(ref, box=None, impl="port", openmp=False):
    """lib/distances.py:324-404 -> include/calc_distances.h:448-531."""
    r = _f32(ref).copy()
    n = len(r)
    out = np.zeros((n * (n - 1) // 2,), dtype=np.float64)
    if out.size == 0:
        return out
    mode, b = _mode_box(box)
    L, pre = _dist_lib(impl, openmp)
    getattr(L, pre + "self_distance_array")(_p(r, _F32P), C.c_uint64(n), mode, _p(b, _F32P),
                                            _p(out, _F64P))
    return out


def _tuple_fn(name, arrays, box, impl, openmp=False):
    arrs = [_f32(a).copy() for a in arrays]
    n = len(arrs[0])
    out = np.zeros((n,), dtype=np.float64)
    if n == 0:
        return out
    mode, b = _mode_box(box)
    L, pre = _dist_lib(impl, openmp)
    getattr(L, pre + name)(*[_p(a, _F32P) for a in arrs], C.c_uint64(n), mode, _p(b, _F32P),
                           _p(out, _F64P))
    return out


def calc_bonds(a, b, box=None, impl="port", openmp=False):
    """lib/distances.py:1414-1515."""
    return _tuple_fn("bonds", (a, b), box, impl, openmp)


def calc_angles(a, b, c, box=None, impl="port", openmp=False):
    """lib/distances.py:1518-1617."""
    return _tuple_fn("angles", (a, b, c), box, impl, openmp)


def calc_dihedrals(a, b, c, d, box=None, impl="port", openmp=False):
    """lib/distances.py:1620-1735."""
    return _tuple_fn("dihedrals", (a, b, c, d), box, impl, openmp)


def _c_round(x):
    """libm round(): half away from zero, on float64 (c_distances.pyx:38 `cround`)."""
    x = np.asarray(x, dtype=np.float64)
    t = np.trunc(x)
    return t + np.where(np.abs(x - t) >= 0.5, np.sign(x), 0.0)


def minimize_vectors(vectors, box):
    """lib/distances.py:1790-1830 -> c_distances.pyx:272-414, restated with NumPy in the
    dtype the generated C evaluates each operation in (fused type = the vectors' dtype)."""
    boxtype, b = check_box(box)
    v = np.array(vectors)
    T = v.dtype.type
    if T not in (np.float32, np.float64):
        raise TypeError("minimize_vectors: float32 or float64 only")
    b = np.asarray(b).astype(T)
    out = v.copy()
    if boxtype == BOX_ORTHO:
        inv = (T(1.0) / b).astype(T)
        for k in range(3):
            if b[k] > 0:
                s = (inv[k] * out[:, k]).astype(T)
                s64 = s.astype(np.float64)
                out[:, k] = (np.float64(b[k]) * (s64 - _c_round(s64))).astype(T)
        return out
    b = b.ravel()
    inv = [T(1.0) / b[0], T(1.0) / b[4], T(1.0) / b[8]]
    dx = [out[:, 0].copy(), out[:, 1].copy(), out[:, 2].copy()]
    s = _c_round((inv[2] * dx[2]).astype(T)).astype(T)
    dx[0] = (dx[0] - (s * b[6]).astype(T)).astype(T)
    dx[1] = (dx[1] - (s * b[7]).astype(T)).astype(T)
    dx[2] = (dx[2] - (s * b[8]).astype(T)).astype(T)
    s = _c_round((inv[1] * dx[1]).astype(T)).astype(T)
    dx[0] = (dx[0] - (s * b[3]).astype(T)).astype(T)
    dx[1] = (dx[1] - (s * b[4]).astype(T)).astype(T)
    s = _c_round((inv[0] * dx[0]).astype(T)).astype(T)
    dx[0] = (dx[0] - (s * b[0]).astype(T)).astype(T)
    n = len(v)
    best = np.full(n, np.finfo(T).max, dtype=T)
    m = np.zeros((n, 3), dtype=T)
    for ix in (-1, 0, 1):
        rx = (dx[0] + T(b[0] * T(ix))).astype(T)
        for iy in (-1, 0, 1):
            ry0 = (rx + T(b[3] * T(iy))).astype(T)
            ry1 = (dx[1] + T(b[4] * T(iy))).astype(T)
            for iz in (-1, 0, 1):
                rz0 = (ry0 + T(b[6] * T(iz))).astype(T)
                rz1 = (ry1 + T(b[7] * T(iz))).astype(T)
                rz2 = (dx[2] + T(b[8] * T(iz))).astype(T)
                dsq = (((rz0 * rz0).astype(T) + (rz1 * rz1).astype(T)).astype(T) + (rz2 * rz2).astype(T)).astype(T)
                better = dsq < best
                best = np.where(better, dsq, best)
                m[better, 0], m[better, 1], m[better, 2] = rz0[better], rz1[better], rz2[better]
    return m


def coord_transform(coords, matrix, impl="port"):
    """include/calc_distances.h:533-555 (in place on a copy)."""
    xyz = _f32(coords).copy()
    m = np.ascontiguousarray(matrix, dtype=np.float64)
    L, pre = _dist_lib(impl)
    getattr(L, pre + "coord_transform")(_p(xyz, _F32P), C.c_uint64(len(xyz)), _p(m, _F64P))
    return xyz


def transform_RtoS(coords, box, impl="port"):
    """lib/distances.py:1320-1367."""
    mode, b = check_box(box)
    m = np.diag(b) if mode == BOX_ORTHO else b.reshape(3, 3)
    inv = np.array(np.linalg.inv(m.astype(np.float64)), order="C")
    return coord_transform(coords, inv, impl)


def transform_StoR(coords, box, impl="port"):
    """lib/distances.py:1370-1411."""
    mode, b = check_box(box)
    m = np.diag(b) if mode == BOX_ORTHO else b.reshape(3, 3)
    return coord_transform(coords, m.astype(np.float64), impl)


# --------------------------------------------------------------------------
# FastNS (lib/nsgrid.pyx)
# --------------------------------------------------------------------------
class FastNS:
    """Restatement of lib/nsgrid.pyx:194-644 on top of oracle.c."""

    def __init__(self, cutoff, coords, box, pbc=True):
        coords = np.asarray(coords)
        box = np.asarray(box)
        if coords.ndim != 2 or coords.shape[1] != 3:
            raise ValueError(f"coords must have a shape of (n, 3), got {coords.shape}.")
        if box.shape != (6,):
            raise ValueError(f"Box must be a numpy array of [lx, ly, lz, alpha, beta, gamma], got {box}")
        if (box[:3] == 0.0).any():
            raise ValueError("Any of the box dimensions cannot be 0")
        if cutoff < 0:
            raise ValueError("Cutoff must be positive")
        self.cutoff = float(cutoff)
        self._L = _port()
        box6 = _f32(box)
        h9 = np.ascontiguousarray(triclinic_vectors(box6).reshape(9))
        xyz = _f32(coords)
        self._g = C.c_void_p(self._L.orc_grid_create(C.c_double(self.cutoff), _p(xyz, _F32P),
                                                     C.c_uint64(len(xyz)), _p(box6, _F32P),
                                                     _p(h9, _F32P), int(bool(pbc))))
        mc = self._L.orc_grid_max_cutoff(self._g)
        if self.cutoff > mc:
            self._L.orc_grid_free(self._g)
            self._g = None
            raise ValueError(f"Cutoff {cutoff} too large for box (max {mc})")

    def __del__(self):
        if getattr(self, "_g", None):
            self._L.orc_grid_free(self._g)
            self._g = None

    def info(self):
        nc = np.zeros(3, dtype=np.int32)
        cs = np.zeros(9, dtype=np.float64)
        per = np.zeros(3, dtype=np.int32)
        tri = C.c_int(0)
        self._L.orc_grid_info(self._g, _p(nc, _I32P), _p(cs, _F64P), _p(per, _I32P), C.byref(tri))
        return nc, cs, per.astype(bool), bool(tri.value)

    def _collect(self, res):
        n = self._L.orc_pairs_count(res)
        pairs = np.empty((n, 2), dtype=np.int64)
        d2 = np.empty((n,), dtype=np.float64)
        if n:
            self._L.orc_pairs_export(res, _p(pairs, _I64P), _p(d2, _F64P))
        self._L.orc_pairs_free(res)
        return NSResults(pairs, d2)

    def search(self, search_coords):
        q = np.asarray(search_coords)
        if q.ndim != 2 or q.shape[1] != 3:
            raise ValueError(f"search_coords must have a shape of (n, 3), got {q.shape}.")
        q = _f32(q)
        res = C.c_void_p(self._L.orc_pairs_new())
        self._L.orc_grid_search(self._g, _p(q, _F32P), C.c_uint64(len(q)), res)
        return self._collect(res)

    def self_search(self):
        res = C.c_void_p(self._L.orc_pairs_new())
        self._L.orc_grid_self_search(self._g, res)
        return self._collect(res)


class NSResults:
    """lib/nsgrid.pyx:127-192."""

    def __init__(self, pairs, d2):
        self._pairs, self._d2 = pairs, d2

    def get_pairs(self):
        return self._pairs.astype(np.intp).reshape(-1, 2)

    def get_pair_distances(self):
        return np.sqrt(self._d2)

    def get_pair_distances2(self):
        return self._d2


def grid_geometry(cutoff, box):
    """Grid geometry only (nsgrid.pyx:294-387) -> ncells, cellsize, periodic, triclinic, max_cutoff."""
    L = _port()
    box6 = _f32(box)
    h9 = np.ascontiguousarray(triclinic_vectors(box6).reshape(9))
    nc = np.zeros(3, dtype=np.int32)
    cs = np.zeros(9, dtype=np.float64)
    per = np.zeros(3, dtype=np.int32)
    tri = C.c_int(0)
    mc = C.c_float(0)
    L.orc_grid_geometry(C.c_double(cutoff), _p(box6, _F32P), _p(h9, _F32P), 1, _p(nc, _I32P),
                        _p(cs, _F64P), _p(per, _I32P), C.byref(tri), C.byref(mc))
    return nc, cs, per.astype(bool), bool(tri.value), mc.value


# --------------------------------------------------------------------------
# capped distances (lib/distances.py:407-1317), methods bruteforce | nsgrid
# --------------------------------------------------------------------------
def _determine_method(reference, configuration, max_cutoff, box=None, method=None):
    """lib/distances.py:516-586 (pkdtree is never auto-selected here)."""
    if method is not None:
        return method.lower()
    if len(reference) < 10 or len(configuration) < 10:
        return "bruteforce"
    if len(reference) * len(configuration) >= 1e8:
        return "nsgrid"
    if box is None:
        lo = np.minimum(reference.min(axis=0), configuration.min(axis=0))
        hi = np.maximum(reference.max(axis=0), configuration.max(axis=0))
        size = hi - lo
    elif np.all(box[3:] == 90.0):
        size = box[:3]
    else:
        t = triclinic_vectors(box)
        size = t.max(axis=0) - t.min(axis=0)
    if np.any(max_cutoff > 0.3 * size):
        return "bruteforce"
    return "nsgrid"


def _determine_method_self(reference, max_cutoff, box=None, method=None):
    """lib/distances.py:985-1047; 'pkdtree' (auto-picked for cutoff < 3 % of the
    box) is outside the GPU path and is answered by nsgrid here."""
    if method is not None:
        return method.lower()
    if len(reference) < 100:
        return "bruteforce"
    return "nsgrid"


def capped_distance(reference, configuration, max_cutoff, min_cutoff=None, box=None,
                    method=None, return_distances=True, impl="port"):
    ref = _f32(np.atleast_2d(reference))
    conf = _f32(np.atleast_2d(configuration))
    if box is not None:
        box = np.asarray(box, dtype=np.float32)
        if box.shape[0] != 6:
            raise ValueError("Box Argument is of incompatible type.")
    pairs = np.empty((0, 2), dtype=np.intp)
    dist = np.empty((0,), dtype=np.float64)
    meth = _determine_method(ref, conf, max_cutoff, box=box, method=method)
    if len(ref) > 0 and len(conf) > 0:
        if meth == "bruteforce":
            d = distance_array(ref, conf, box=box, impl=impl)
            if min_cutoff is not None:
                mask = np.where((d <= max_cutoff) & (d > min_cutoff))
            else:
                mask = np.where(d <= max_cutoff)
            if mask[0].size > 0:
                pairs = np.c_[mask[0], mask[1]]
                dist = d[mask]
        elif meth == "nsgrid":
            if box is None:
                pseudobox = np.zeros(6, dtype=np.float32)
                allc = np.concatenate([ref, conf])
                lmax, lmin = allc.max(axis=0), allc.min(axis=0)
                boxsize = (lmax - lmin).max()
                boxsize = np.maximum(boxsize, 2 * max_cutoff)
                pseudobox[:3] = boxsize + 2.2 * max_cutoff
                pseudobox[3:] = 90.0
                sref, sconf = ref.copy(), conf.copy()
                sref -= lmin - 0.1 * max_cutoff
                sconf -= lmin - 0.1 * max_cutoff
                res = FastNS(max_cutoff, sconf, box=pseudobox, pbc=False).search(sref)
            else:
                res = FastNS(max_cutoff, conf, box=box).search(ref)
            pairs = res.get_pairs()
            dist = res.get_pair_distances()
            if min_cutoff is not None:
                idx = dist > min_cutoff
                pairs, dist = pairs[idx], dist[idx]
        else:
            raise KeyError(meth)
    return (pairs, dist) if return_distances else pairs


def self_capped_distance(reference, max_cutoff, min_cutoff=None, box=None, method=None,
                         return_distances=True, impl="port"):
    ref = _f32(np.atleast_2d(reference))
    if box is not None:
        box = np.asarray(box, dtype=np.float32)
        if box.shape[0] != 6:
            raise ValueError("Box Argument is of incompatible type.")
    pairs = np.empty((0, 2), dtype=np.intp)
    dist = np.empty((0,), dtype=np.float64)
    meth = _determine_method_self(ref, max_cutoff, box=box, method=method)
    if len(ref) > 1:
        if meth == "bruteforce":
            n = len(ref)
            d = self_distance_array(ref, box=box, impl=impl)
            if min_cutoff is not None:
                idx = np.where((d <= max_cutoff) & (d > min_cutoff))[0]
            else:
                idx = np.where(d <= max_cutoff)[0]
            if idx.size:
                iu = np.triu_indices(n, k=1)
                pairs = np.c_[iu[0][idx], iu[1][idx]].astype(np.intp)
                dist = d[idx]
        elif meth == "nsgrid":
            if box is None:
                pseudobox = np.zeros(6, dtype=np.float32)
                lmax, lmin = ref.max(axis=0), ref.min(axis=0)
                boxsize = (lmax - lmin).max()
                sizefactor = 2.2 * max_cutoff / boxsize if boxsize < 2 * max_cutoff else 1.2
                pseudobox[:3] = sizefactor * boxsize
                pseudobox[3:] = 90.0
                sref = ref.copy()
                sref -= lmin - 0.1 * boxsize
                res = FastNS(max_cutoff, sref, box=pseudobox, pbc=False).self_search()
            else:
                res = FastNS(max_cutoff, ref, box=box).self_search()
            pairs = res.get_pairs()
            dist = res.get_pair_distances()
            if min_cutoff is not None:
                idx = dist > min_cutoff
                pairs, dist = pairs[idx], dist[idx]
        else:
            raise KeyError(meth)
    return (pairs, dist) if return_distances else pairs


# --------------------------------------------------------------------------
# QCP (lib/qcprot.pyx) and its drivers
# --------------------------------------------------------------------------
def CalcRMSDRotationalMatrix(ref, conf, N, rot, weights):
    """lib/qcprot.pyx:282-320 (float32 or float64, all arrays the same dtype)."""
    L = _port()
    ref = np.ascontiguousarray(ref)
    if ref.dtype == np.float32:
        conf = np.ascontiguousarray(conf, dtype=np.float32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        r = None if rot is None else np.zeros(9, dtype=np.float32)
        val = L.orc_qcp_f32(_p(ref, _F32P), _p(conf, _F32P), int(N),
                            None if r is None else _p(r, _F32P),
                            None if w is None else _p(w, _F32P))
    else:
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        conf = np.ascontiguousarray(conf, dtype=np.float64)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        r = None if rot is None else np.zeros(9, dtype=np.float64)
        val = L.orc_qcp_f64(_p(ref, _F64P), _p(conf, _F64P), int(N),
                            None if r is None else _p(r, _F64P),
                            None if w is None else _p(w, _F64P))
    if rot is not None:
        rot[:] = r
    return float(val)


def rmsd(a, b, weights=None, center=False, superposition=False):
    """analysis/rms.py:178-283."""
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
        return CalcRMSDRotationalMatrix(a, b, N, None, weights)
    if weights is not None:
        return np.sqrt(np.sum(weights[:, np.newaxis] * ((a - b) ** 2)) / N)
    return np.sqrt(np.sum((a - b) ** 2) / N)


def distance_matrix(traj, select=None, weights=None, superposition=False, cutoff=1E0 - 5):
    """analysis/diffusionmap.py:238-268: DistanceMatrix with metric = rms.rmsd (optionally with
    superposition=True), as the reference's own double loop over frame pairs."""
    traj = np.asarray(traj, dtype=np.float32)
    if select is not None:
        traj = traj[:, np.asarray(select)]
    F = traj.shape[0]
    m = np.zeros((F, F))
    for i in range(F):
        for j in range(i, F):
            d = rmsd(traj[i], traj[j], weights=weights, superposition=superposition)
            m[i, j] = d if d > cutoff else 0
            m[j, i] = m[i, j]
    return m


def _center(xyz_f32, weights):
    """core/groups.py:1083-1106, compound='group'."""
    if weights is None:
        return xyz_f32.astype(np.float64, copy=False).mean(axis=0)
    w = np.asarray(weights).astype(np.float64, copy=False)
    return np.einsum("ij,ij->j", xyz_f32, w[:, None]) / w.sum()


def rmsd_trajectory(traj, reference, weights=None):
    """analysis/rms.py:620-727 without groupselections: per-frame best-fit RMSD.

    traj float32 [F,N,3]; reference float32 [N,3]; weights None or [N]
    (e.g. masses; normalised by their mean as in rms.py:639-643)."""
    traj = np.asarray(traj, dtype=np.float32)
    reference = np.asarray(reference, dtype=np.float32)
    w_sel = None
    if weights is not None:
        w_sel = np.asarray(weights, dtype=np.float64) / np.mean(weights)
    ref_com = _center(reference, w_sel)
    ref64 = (reference - ref_com).astype(np.float64)
    n = reference.shape[0]
    out = np.zeros(len(traj), dtype=np.float64)
    mob64 = np.empty((n, 3), dtype=np.float64)
    for f in range(len(traj)):
        com = _center(traj[f], w_sel).astype(np.float64)
        mob64[:] = traj[f]
        mob64 -= com
        out[f] = CalcRMSDRotationalMatrix(ref64, mob64, n, None, w_sel)
    return out


def rotation_matrix(a, b, weights=None):
    """analysis/align.py:224-307."""
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
    val = CalcRMSDRotationalMatrix(a, b, N, rot, weights)
    return rot.reshape(3, 3), val


def align_trajectory(traj, reference, select=None, weights=None):
    """analysis/align.py:836-854 + _fit_to :357-366 + groups.py:1478-1529.

    traj float32 [F,Nall,3] (a modified copy is returned); reference float32
    [Nsel,3] = reference positions of the selected atoms; select = index array
    into the Nall atoms (None = all).  Returns (aligned float32, rmsd[F])."""
    traj = np.array(traj, dtype=np.float32, copy=True)
    reference = np.asarray(reference, dtype=np.float32)
    sel = np.arange(traj.shape[1]) if select is None else np.asarray(select)
    w = None if weights is None else np.asarray(weights)
    ref_com = _center(reference, w)
    ref_coords = reference - ref_com            # float64
    out = np.zeros(len(traj), dtype=np.float64)
    for f in range(len(traj)):
        x = traj[f]
        mob = x[sel]
        com = _center(mob, w)
        R, out[f] = rotation_matrix(mob - com, ref_coords, weights=w)
        x += -com                               # float64 add, stored as float32
        x[:] = np.dot(x, R.T)
        x += ref_com
    return traj, out


# --------------------------------------------------------------------------
# InterRDF (analysis/rdf.py:253-316)
# --------------------------------------------------------------------------
def interrdf(pos_a, pos_b, dimensions, nbins=75, rrange=(0.0, 15.0), norm="rdf",
             exclusion_block=None, frames=None, exclude_same_ids=None):
    """pos_a [F,Na,3], pos_b [F,Nb,3] float32, dimensions [F,6] or [6]."""
    pos_a = np.asarray(pos_a, dtype=np.float32)
    pos_b = np.asarray(pos_b, dtype=np.float32)
    F = len(pos_a)
    dims = np.asarray(dimensions, dtype=np.float32)
    if dims.ndim == 1:
        dims = np.broadcast_to(dims, (F, 6))
    count, edges = np.histogram([-1], bins=nbins, range=rrange)
    count = count.astype(np.float64) * 0.0
    volume_cum = 0.0
    idx = range(F) if frames is None else frames
    nfr = 0
    for f in idx:
        pairs, dist = capped_distance(pos_a[f], pos_b[f], rrange[1], box=dims[f])
        if exclusion_block is not None:
            ia = pairs[:, 0] // exclusion_block[0]
            ib = pairs[:, 1] // exclusion_block[1]
            dist = dist[np.where(ia != ib)[0]]
        if exclude_same_ids is not None:                   # rdf.py:280-285
            ia = np.asarray(exclude_same_ids[0])[pairs[:, 0]]
            ib = np.asarray(exclude_same_ids[1])[pairs[:, 1]]
            dist = dist[np.where(ia != ib)[0]]
        c, _ = np.histogram(dist, bins=nbins, range=rrange)
        count += c
        volume_cum += box_volume(dims[f])
        nfr += 1
    res = {"count": count, "edges": edges, "bins": 0.5 * (edges[:-1] + edges[1:])}
    normv = float(nfr)
    if norm in ("rdf", "density"):
        vols = np.power(edges, 3)
        normv = normv * (4 / 3 * np.pi * np.diff(vols))
    if norm == "rdf":
        nA, nB = pos_a.shape[1], pos_b.shape[1]
        N = nA * nB
        if exclusion_block:
            xA, xB = exclusion_block
            N -= xA * xB * (nA / xA)
        normv = normv * (N / (volume_cum / nfr))
    res["rdf"] = count / normv
    res["volume_cum"] = volume_cum
    return res


# --------------------------------------------------------------------------
# InterRDF_s (analysis/rdf.py:573-640)
# --------------------------------------------------------------------------
def interrdf_s(pos, ags, dimensions, nbins=75, rrange=(0.0, 15.0), norm="rdf"):
    """pos [F, n, 3] float32, ags = [[idx1, idx2], ...]; returns count / rdf / cdf lists."""
    pos = np.asarray(pos, dtype=np.float32)
    F = len(pos)
    dims = np.asarray(dimensions, dtype=np.float32)
    if dims.ndim == 1:
        dims = np.broadcast_to(dims, (F, 6))
    c0, edges = np.histogram([-1], bins=nbins, range=rrange)
    counts = [np.zeros((len(a), len(b), len(c0)), dtype=np.float64) for a, b in ags]
    volume_cum = 0.0
    for f in range(F):
        for i, (a, b) in enumerate(ags):
            pairs, dist = capped_distance(np.ascontiguousarray(pos[f][a]), np.ascontiguousarray(pos[f][b]),
                                          rrange[1], box=dims[f])
            for j, (i1, i2) in enumerate(pairs):          # rdf.py:593-595
                c, _ = np.histogram(dist[j], bins=nbins, range=rrange)
                counts[i][i1, i2, :] += c
        if norm == "rdf":
            volume_cum += box_volume(dims[f])
    normv = float(F)
    if norm in ("rdf", "density"):
        normv = normv * (4 / 3 * np.pi * np.diff(np.power(edges, 3)))
    if norm == "rdf":
        normv = normv * (1 / (volume_cum / F))
    return {"count": counts, "edges": edges, "rdf": [c / normv for c in counts],
            "cdf": [np.cumsum(c, axis=2) / F for c in counts]}


# --------------------------------------------------------------------------
# Contacts (analysis/contacts.py:231-338, 341-499)
# --------------------------------------------------------------------------
def contacts_timeseries(pos, idx_a, idx_b, refs, dimensions=None, method="hard_cut", radius=4.5,
                        pbc=True, ref_box=None, kwargs=None):
    """pos [F, n, 3]; refs = [(refA [na,3], refB [nb,3]), ...]; returns timeseries [F, len(refs)+1]."""
    kwargs = kwargs or {}
    pos = np.asarray(pos, dtype=np.float32)
    F = len(pos)
    dims = None
    if dimensions is not None and pbc:
        dims = np.asarray(dimensions, dtype=np.float32)
        if dims.ndim == 1:
            dims = np.broadcast_to(dims, (F, 6))
    r0s, masks = [], []
    for refA, refB in refs:
        r0 = distance_array(np.asarray(refA, dtype=np.float32), np.asarray(refB, dtype=np.float32),
                            box=ref_box if pbc else None)
        r0s.append(r0)
        masks.append(r0 <= radius)
    ts = np.empty((F, len(refs) + 1))
    for f in range(F):
        ts[f, 0] = f
        d = distance_array(np.ascontiguousarray(pos[f][idx_a]), np.ascontiguousarray(pos[f][idx_b]),
                           box=None if dims is None else dims[f])
        for i, (mask, r0) in enumerate(zip(masks, r0s), 1):
            r, rr0 = d[mask], r0[mask]
            with np.errstate(invalid="ignore", divide="ignore"):
                if method == "hard_cut":
                    q = (r <= rr0).sum() / r.size
                elif method == "radius_cut":
                    q = (r <= radius).sum() / r.size
                elif method == "soft_cut":
                    beta, lam = kwargs.get("beta", 5.0), kwargs.get("lambda_constant", 1.8)
                    q = (1 / (1 + np.exp(beta * (r - lam * rr0)))).sum() / len(rr0)
                else:
                    q = method(r, rr0, **kwargs)
            ts[f, i] = q
    return ts
