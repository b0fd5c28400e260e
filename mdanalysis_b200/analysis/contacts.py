"""Contacts on the GPU -- the pieces of ``MDAnalysis/analysis/contacts.py`` that sit
on the hot path: ``contact_matrix`` (:315-338), ``hard_cut_q`` (:268-297) and a
fused ``self_distance_array -> d <= radius`` reduction that never materialises
the n(n-1)/2 distances (config "contacts on a 100k-atom
selection").
"""
import numpy as np

from .._cabi import (BOX_NONE, DEVICE, HOST, check, current_stream, is_tensor, load, ptr,
                     tensor_ptr)
from ..frames import as_block
from ..lib.distances import _box_args
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
