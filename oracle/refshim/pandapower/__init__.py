"""Stub: only imported, never called on the Laurent power-flow path."""
from . import topology  # noqa: F401
