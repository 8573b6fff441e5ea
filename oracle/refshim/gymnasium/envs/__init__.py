from . import registration  # noqa: F401


def register(*a, **k):
    pass
