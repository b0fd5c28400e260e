"""GPU drivers for the analyses that sit on the geometry hot path."""
from . import base, rms, align, rdf, contacts, distances  # noqa: F401
