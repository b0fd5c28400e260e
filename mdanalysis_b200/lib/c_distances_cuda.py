"""The ``cuda`` backend module for ``MDAnalysis.lib.distances._distances``.

Exports exactly the 18 function names, argument orders and in-place semantics
of the reference's Cython shims (``MDAnalysis/lib/c_distances.pyx:48-222``,
``c_distances.pxd:10-27``), so it can be registered without any source change:

    MDAnalysis.lib.distances._distances['cuda'] = mdanalysis_b200.lib.c_distances_cuda

(see ``mdanalysis_b200.register`` and INTEGRATION.md).  Arguments are the
arrays ``check_coords`` / ``check_box`` / ``_check_result_array`` produce:
float32 C-contiguous ``(n, 3)`` coordinates, ``float32[3]`` or ``float32[3, 3]``
boxes and a caller-owned float64 ``result`` that is filled in place (it may be
non-contiguous -- the reference does not check, distances.py:229-231).

Everything is computed by libmdagpu.so; there is no CPU path.
"""
import numpy as np

from .. import _cabi
from .._cabi import BOX_NONE, BOX_ORTHO, BOX_TRICLINIC, HOST, check, load, ptr

_UINT64_MAX = np.iinfo(np.uint64).max
OPENMP_ENABLED = False   # attribute the reference reads from the openmp shim


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _Result:
    """Gives the C side a contiguous float64 buffer and copies back if needed."""

    def __init__(self, result):
        self.user = result
        self.buf = result if result.flags["C_CONTIGUOUS"] else np.empty(result.shape, np.float64)

    def done(self):
        if self.buf is not self.user:
            self.user[...] = self.buf


def _box(mode, box):
    if mode == BOX_NONE:
        return None
    return _f32c(box).reshape(-1)


def _distance_array(ref, conf, mode, box, result):
    ref, conf, b = _f32c(ref), _f32c(conf), _box(mode, box)
    r = _Result(result)
    check(load().mdagpu_distance_array(ptr(ref), ref.shape[0], ptr(conf), conf.shape[0], mode,
                                       ptr(b), ptr(r.buf), 1, HOST, None))
    r.done()


def calc_distance_array(ref, conf, result):
    _distance_array(ref, conf, BOX_NONE, None, result)


def calc_distance_array_ortho(ref, conf, box, result):
    _distance_array(ref, conf, BOX_ORTHO, box, result)


def calc_distance_array_triclinic(ref, conf, box, result):
    _distance_array(ref, conf, BOX_TRICLINIC, box, result)


def _self_distance_array(ref, mode, box, result):
    ref, b = _f32c(ref), _box(mode, box)
    r = _Result(result)
    check(load().mdagpu_self_distance_array(ptr(ref), ref.shape[0], mode, ptr(b), ptr(r.buf), 1,
                                            HOST, None))
    r.done()


def calc_self_distance_array(ref, result):
    _self_distance_array(ref, BOX_NONE, None, result)


def calc_self_distance_array_ortho(ref, box, result):
    _self_distance_array(ref, BOX_ORTHO, box, result)


def calc_self_distance_array_triclinic(ref, box, result):
    _self_distance_array(ref, BOX_TRICLINIC, box, result)


def coord_transform(coords, box):
    """In place: coords <- coords . box (box float64 3x3)."""
    m = np.ascontiguousarray(box, dtype=np.float64)
    work = coords if (coords.flags["C_CONTIGUOUS"] and coords.dtype == np.float32) else _f32c(coords)
    check(load().mdagpu_coord_transform(ptr(work), work.shape[0], ptr(m), HOST, None))
    if work is not coords:
        coords[...] = work


def _tuples(fn, arrays, mode, box, result):
    arrs = [_f32c(a) for a in arrays]
    b = _box(mode, box)
    r = _Result(result)
    check(fn(*[ptr(a) for a in arrs], arrs[0].shape[0], mode, ptr(b), ptr(r.buf), HOST, None))
    r.done()


def calc_bond_distance(coords1, coords2, results):
    _tuples(load().mdagpu_calc_bonds, (coords1, coords2), BOX_NONE, None, results)


def calc_bond_distance_ortho(coords1, coords2, box, results):
    _tuples(load().mdagpu_calc_bonds, (coords1, coords2), BOX_ORTHO, box, results)


def calc_bond_distance_triclinic(coords1, coords2, box, results):
    _tuples(load().mdagpu_calc_bonds, (coords1, coords2), BOX_TRICLINIC, box, results)


def calc_angle(coords1, coords2, coords3, results):
    _tuples(load().mdagpu_calc_angles, (coords1, coords2, coords3), BOX_NONE, None, results)


def calc_angle_ortho(coords1, coords2, coords3, box, results):
    _tuples(load().mdagpu_calc_angles, (coords1, coords2, coords3), BOX_ORTHO, box, results)


def calc_angle_triclinic(coords1, coords2, coords3, box, results):
    _tuples(load().mdagpu_calc_angles, (coords1, coords2, coords3), BOX_TRICLINIC, box, results)


def calc_dihedral(coords1, coords2, coords3, coords4, results):
    _tuples(load().mdagpu_calc_dihedrals, (coords1, coords2, coords3, coords4), BOX_NONE, None, results)


def calc_dihedral_ortho(coords1, coords2, coords3, coords4, box, results):
    _tuples(load().mdagpu_calc_dihedrals, (coords1, coords2, coords3, coords4), BOX_ORTHO, box, results)


def calc_dihedral_triclinic(coords1, coords2, coords3, coords4, box, results):
    _tuples(load().mdagpu_calc_dihedrals, (coords1, coords2, coords3, coords4), BOX_TRICLINIC, box, results)


def _wrap(coords, mode, box):
    b = _box(mode, box)
    work = coords if (coords.flags["C_CONTIGUOUS"] and coords.dtype == np.float32) else _f32c(coords)
    check(load().mdagpu_wrap(ptr(work), work.shape[0], mode, ptr(b), HOST, None))
    if work is not coords:
        coords[...] = work


def ortho_pbc(coords, box):
    _wrap(coords, BOX_ORTHO, box)


def triclinic_pbc(coords, box):
    _wrap(coords, BOX_TRICLINIC, box)


FUNCTION_NAMES = (
    "calc_distance_array", "calc_distance_array_ortho", "calc_distance_array_triclinic",
    "calc_self_distance_array", "calc_self_distance_array_ortho",
    "calc_self_distance_array_triclinic", "coord_transform", "calc_bond_distance",
    "calc_bond_distance_ortho", "calc_bond_distance_triclinic", "calc_angle", "calc_angle_ortho",
    "calc_angle_triclinic", "calc_dihedral", "calc_dihedral_ortho", "calc_dihedral_triclinic",
    "ortho_pbc", "triclinic_pbc",
)
