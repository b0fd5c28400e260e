"""``FastNS`` / ``NSResults`` on the GPU -- mirror of ``MDAnalysis/lib/nsgrid.pyx``.

Same constructor checks and error messages (nsgrid.pyx:278-291), same methods:
``FastNS(cutoff, coords, box, pbc=True)``, ``.search(coords)``, ``.self_search()``
returning an ``NSResults`` with ``get_pairs()`` / ``get_pair_distances()``
(nsgrid.pyx:127-192).  The grid geometry is the reference's, so the pair sets
are identical to the reference's for every box, including the triclinic boxes
where its 27-cell stencil is not complete.  Pair order is unspecified.
"""
import ctypes as C

import numpy as np

from .._cabi import DEVICE, HOST, check, current_stream, is_tensor, load, ptr, tensor_ptr
from .mdamath import triclinic_vectors


class NSResults:
    """Pairs and distances of a search; buffers stay on the GPU until asked for."""

    def __init__(self, handle):
        self._h = handle
        self._pairs = None
        self._dist = None

    def __del__(self):
        self._release()

    def _release(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            load().mdagpu_pairs_free(h)

    def __len__(self):
        if self._pairs is not None:
            return len(self._pairs)
        return int(load().mdagpu_pairs_count(self._h)) if self._h else 0

    def _fetch(self):
        if self._pairs is None:
            n = len(self)
            pairs = np.empty((n, 2), dtype=np.int64)
            dist = np.empty((n,), dtype=np.float64)
            if n:
                check(load().mdagpu_pairs_export(self._h, ptr(pairs), ptr(dist), HOST, None))
            self._pairs, self._dist = pairs, dist
            self._release()

    def get_pairs(self):
        """intp (n, 2): ``[query index, grid index]`` (search) or an unordered pair (self_search)."""
        self._fetch()
        return self._pairs.astype(np.intp, copy=False).reshape(-1, 2)

    def get_pair_distances(self):
        """float64 (n,): distance of each pair (sqrt of the float64 squared distance)."""
        self._fetch()
        return self._dist

    def to_tensors(self, device=None):
        """Device-resident copy: (pairs int64 [n,2], distances float64 [n]) torch tensors."""
        import torch
        if self._pairs is not None:
            return torch.as_tensor(self._pairs, device=device or "cuda"), \
                torch.as_tensor(self._dist, device=device or "cuda")
        n = len(self)
        pairs = torch.empty((n, 2), dtype=torch.int64, device=device or "cuda")
        dist = torch.empty((n,), dtype=torch.float64, device=device or "cuda")
        if n:
            check(load().mdagpu_pairs_export(self._h, tensor_ptr(pairs), tensor_ptr(dist), DEVICE,
                                             current_stream(pairs)))
            torch.cuda.current_stream(pairs.device).synchronize()
        return pairs, dist


class FastNS:
    """Grid based neighbour search (minimum image if ``pbc``)."""

    def __init__(self, cutoff, coords, box, pbc=True):
        dev = is_tensor(coords)
        shape = tuple(coords.shape)
        if len(shape) != 2 or shape[1] != 3:
            raise ValueError(f"coords must have a shape of (n, 3), got {shape}.")
        box = np.asarray(box)
        if box.shape != (6,):
            raise ValueError("Box must be a numpy array of [lx, ly, lz, alpha, beta, gamma], "
                             f"got {box}")
        if (box[:3] == 0.0).any():
            raise ValueError("Any of the box dimensions cannot be 0")
        if cutoff < 0:
            raise ValueError("Cutoff must be positive")
        self.cutoff = float(cutoff)
        box6 = np.ascontiguousarray(box, dtype=np.float32)
        h9 = np.ascontiguousarray(triclinic_vectors(box6).reshape(9), dtype=np.float32)
        self._h = C.c_void_p()
        if dev:
            import torch
            xyz = coords.to(torch.float32).contiguous()
            rc = load().mdagpu_grid_create(self.cutoff, tensor_ptr(xyz), xyz.shape[0], ptr(box6),
                                           ptr(h9), int(bool(pbc)), DEVICE, current_stream(xyz),
                                           C.byref(self._h))
        else:
            xyz = np.ascontiguousarray(coords, dtype=np.float32)
            rc = load().mdagpu_grid_create(self.cutoff, ptr(xyz), xyz.shape[0], ptr(box6), ptr(h9),
                                           int(bool(pbc)), HOST, None, C.byref(self._h))
        check(rc)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            load().mdagpu_grid_free(h)

    def grid_info(self):
        """(ncells int[3], cellsize float64[9], periodic bool[3], triclinic, max_cutoff)."""
        nc = np.zeros(3, dtype=np.int32)
        cs = np.zeros(9, dtype=np.float64)
        per = np.zeros(3, dtype=np.int32)
        tri, mc = C.c_int(0), C.c_float(0)
        check(load().mdagpu_grid_info(self._h, ptr(nc), ptr(cs), ptr(per), C.byref(tri), C.byref(mc)))
        return nc, cs, per.astype(bool), bool(tri.value), mc.value

    def search(self, search_coords):
        """Neighbours of every query coordinate among the grid coordinates."""
        shape = tuple(search_coords.shape)
        if len(shape) != 2 or shape[1] != 3:
            raise ValueError(f"search_coords must have a shape of (n, 3), got {shape}.")
        out = C.c_void_p()
        if is_tensor(search_coords):
            import torch
            q = search_coords.to(torch.float32).contiguous()
            rc = load().mdagpu_grid_search(self._h, tensor_ptr(q), q.shape[0], DEVICE,
                                           current_stream(q), C.byref(out))
        else:
            if search_coords.dtype != np.float32:
                # the reference's typed memoryview accepts float32 only
                raise ValueError(f"Buffer dtype mismatch, expected 'float' but got "
                                 f"'{search_coords.dtype}'")
            q = np.ascontiguousarray(search_coords, dtype=np.float32)
            rc = load().mdagpu_grid_search(self._h, ptr(q), q.shape[0], HOST, None, C.byref(out))
        check(rc)
        return NSResults(out)

    def self_search(self):
        """All pairs within the grid coordinates, each once."""
        out = C.c_void_p()
        check(load().mdagpu_grid_self_search(self._h, None, C.byref(out)))
        return NSResults(out)
