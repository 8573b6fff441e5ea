"""Stub: the reference imports matplotlib at module import time only."""
def use(*a, **k):
    pass
