"""``contact_matrix`` of ``MDAnalysis/analysis/distances.py`` (:61-120) on the GPU."""
import numpy as np

from ..lib.distances import capped_distance


def contact_matrix(coord, cutoff=15.0, returntype="numpy", box=None):
    """Boolean n x n matrix of contacts within ``cutoff`` (diagonal included).

    ``returntype='sparse'`` returns a ``scipy.sparse.lil_matrix`` built from the
    same GPU pair list (the reference uses two dedicated CPU loops for it,
    c_distances.pyx:226-270; the resulting matrix is the same).
    """
    pairs = capped_distance(coord, coord, max_cutoff=cutoff, box=box, return_distances=False)
    n = len(coord)
    if returntype == "numpy":
        adj = np.full((n, n), False, dtype=bool)
        if len(pairs):
            idx, idy = np.transpose(pairs)
            adj[idx, idy] = True
        return adj
    if returntype == "sparse":
        import scipy.sparse
        m = scipy.sparse.lil_matrix((n, n), dtype="bool")
        if len(pairs):
            idx, idy = np.transpose(pairs)
            m[idx, idy] = True
        return m
    raise ValueError(f"unknown returntype {returntype!r}")
