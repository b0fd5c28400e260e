"""Host-side mirror of ``MDAnalysis.lib`` for the geometry hot path (CUDA backend)."""
from . import mdamath, util, distances, nsgrid, qcprot  # noqa: F401
