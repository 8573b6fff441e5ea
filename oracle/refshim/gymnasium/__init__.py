"""Minimal stand-in for `gymnasium` so the read-only Python reference under
/root/reference can be imported in a container that lacks it.
TEST INFRASTRUCTURE ONLY: used by tools/make_golden.py to generate fixtures."""
from . import spaces, envs, utils, core  # noqa: F401
from .core import Env, Wrapper, ActionWrapper, ObservationWrapper  # noqa: F401


def make(*a, **k):
    raise NotImplementedError("refshim gymnasium has no registry")
