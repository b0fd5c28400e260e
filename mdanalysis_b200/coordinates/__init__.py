"""Bulk trajectory ingest for the GPU analyses."""
from .dcd import DCDFile, read_dcd  # noqa: F401
