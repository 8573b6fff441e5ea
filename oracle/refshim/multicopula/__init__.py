"""Stub for the un-vendored `multicopula` dependency (only imported)."""
class EllipticalCopula:
    pass
from . import multicopula  # noqa: E402,F401
