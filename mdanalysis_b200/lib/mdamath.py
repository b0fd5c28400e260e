"""Unit-cell helpers needed by the hot path (host-side NumPy, float64 maths).

Mirrors ``MDAnalysis/lib/mdamath.py``: ``triclinic_vectors`` (:309-405),
``box_volume`` (:408-444) and ``triclinic_box`` (:246-306).  These feed every
triclinic kernel, so their arithmetic (NumPy float64, cast at the end) is kept
identical to the reference -- the float32 cell matrix is part of the parity
contract.
"""
import numpy as np


def triclinic_vectors(dimensions, dtype=np.float32):
    """``[lx, ly, lz, alpha, beta, gamma]`` -> lower-triangular 3x3 cell matrix.

    Invalid cells (non-positive lengths/angles, angles >= 180, or angle triples
    that do not close) give an all-zero matrix, as in the reference.
    """
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim
    valid = np.all(dim > 0.0) and alpha < 180.0 and beta < 180.0 and gamma < 180.0
    if not valid:
        return np.zeros((3, 3), dtype=dtype)
    if alpha == beta == gamma == 90.0:
        return np.diag(dim[:3].astype(dtype, copy=False))
    cell = np.zeros((3, 3), dtype=np.float64)
    cell[0, 0] = lx
    # exact trigonometric values for right angles
    cos_alpha = 0.0 if alpha == 90.0 else np.cos(np.deg2rad(alpha))
    cos_beta = 0.0 if beta == 90.0 else np.cos(np.deg2rad(beta))
    if gamma == 90.0:
        cos_gamma, sin_gamma = 0.0, 1.0
    else:
        gamma = np.deg2rad(gamma)
        cos_gamma, sin_gamma = np.cos(gamma), np.sin(gamma)
    cell[1, 0] = ly * cos_gamma
    cell[1, 1] = ly * sin_gamma
    cell[2, 0] = lz * cos_beta
    cell[2, 1] = lz * (cos_alpha - cos_beta * cos_gamma) / sin_gamma
    cell[2, 2] = np.sqrt(lz * lz - cell[2, 0] ** 2 - cell[2, 1] ** 2)
    if cell[2, 2] > 0.0:
        return cell.astype(dtype, copy=False)
    return np.zeros((3, 3), dtype=dtype)


def box_volume(dimensions):
    """Volume of the unit cell; zero for invalid cells."""
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim
    if alpha == beta == gamma == 90.0 and lx > 0 and ly > 0 and lz > 0:
        return lx * ly * lz
    cell = triclinic_vectors(dim, dtype=np.float64)
    return cell[0, 0] * cell[1, 1] * cell[2, 2]


def triclinic_box(x, y, z, dtype=np.float32):
    """Three cell vectors -> ``[lx, ly, lz, alpha, beta, gamma]`` (mdamath.py:246-306)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    lx, ly, lz = np.linalg.norm(x), np.linalg.norm(y), np.linalg.norm(z)
    if lx == 0 or ly == 0 or lz == 0:
        return np.zeros(6, dtype=dtype)
    alpha = np.rad2deg(np.arccos(np.dot(y, z) / (ly * lz)))
    beta = np.rad2deg(np.arccos(np.dot(x, z) / (lx * lz)))
    gamma = np.rad2deg(np.arccos(np.dot(x, y) / (lx * ly)))
    box = np.array([lx, ly, lz, alpha, beta, gamma], dtype=dtype)
    if np.all(box > 0.0) and alpha < 180.0 and beta < 180.0 and gamma < 180.0:
        return box
    return np.zeros(6, dtype=dtype)
