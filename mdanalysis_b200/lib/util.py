"""Argument checking of the geometry front end.

Mirrors the two helpers of ``MDAnalysis/lib/util.py`` that define the boundary
contract of the hot path: ``check_coords`` (:1965-2207) and ``check_box``
(:2471-2524).  ``check_coords`` is offered as a plain function
(:func:`prepare_coords`) instead of a decorator; the semantics (float32
C-contiguous copies, ``(3,)`` -> ``(1, 3)``, ``.positions`` of AtomGroup-likes,
length checks, error types and messages) are the reference's.
"""
import numpy as np

from .mdamath import triclinic_vectors


def check_box(box):
    """``box`` -> ``('ortho', float32[3])`` or ``('tri_vecs', float32[3,3])``."""
    if box is None:
        raise ValueError("Box is None")
    box = np.asarray(box, dtype=np.float32, order="C")
    if box.shape != (6,):
        raise ValueError("Invalid box information. Must be of the form "
                         "[lx, ly, lz, alpha, beta, gamma].")
    if np.all(box[3:] == 90.0):
        return "ortho", box[:3]
    return "tri_vecs", triclinic_vectors(box)


def prepare_coords(fname, named_coords, *, enforce_copy=True, enforce_dtype=True,
                   allow_single=True, convert_single=True, reduce_result_if_single=True,
                   check_lengths_match=None, allow_atomgroup=False):
    """Validate coordinate arguments the way ``@check_coords`` does.

    ``named_coords`` is a list of ``(argname, value)``.  Returns
    ``(arrays, reduce)`` where ``reduce`` tells the caller to return
    ``result[0]`` (all inputs were single coordinates).
    """
    if check_lengths_match is None:
        check_lengths_match = len(named_coords) > 1
    out, ncoords = [], []
    all_single = allow_single
    for argname, coords in named_coords:
        is_single = False
        if isinstance(coords, np.ndarray):
            if allow_single:
                if coords.ndim not in (1, 2) or coords.shape[-1] != 3:
                    raise ValueError(f"{fname}(): {argname}.shape must be (3,) or "
                                     f"(n, 3), got {coords.shape}")
                if coords.ndim == 1:
                    is_single = True
                    if convert_single:
                        coords = coords[None, :]
            elif coords.ndim != 2 or coords.shape[1] != 3:
                raise ValueError(f"{fname}(): {argname}.shape must be (n, 3) got {coords.shape}")
            if enforce_dtype:
                try:
                    coords = coords.astype(np.float32, order="C", copy=enforce_copy)
                except ValueError:
                    raise TypeError(f"{fname}(): {argname}.dtype must beconvertible to "
                                    f"float32, got {coords.dtype}.") from None
        else:
            try:
                coords = coords.positions
                _ = coords.shape[0]
            except AttributeError:
                raise TypeError(f"{fname}(): Parameter '{argname}' must be a numpy.ndarray "
                                f"or an AtomGroup, got {type(coords)}.") from None
            if not allow_atomgroup:
                raise TypeError("AtomGroup or other class with a`.positions` method supplied "
                                "as anargument, but allow_atomgroup is False")
        all_single &= is_single
        ncoords.append(coords.shape[0])
        out.append(coords)
    if check_lengths_match and ncoords and ncoords.count(ncoords[0]) != len(ncoords):
        names = ", ".join(n for n, _ in named_coords)
        raise ValueError(f"{fname}(): {names} must contain the same number of coordinates, "
                         f"got {ncoords}.")
    return out, bool(all_single and reduce_result_if_single)
