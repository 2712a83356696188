"""Throw-away shim of gymnasium (test infrastructure; see oracle/ref_shims/omegaconf)."""
from . import spaces, logger, error, wrappers  # noqa: F401


class Env:
    observation_space = None
    action_space = None

    def reset(self, **kw):
        raise NotImplementedError

    def step(self, a):
        raise NotImplementedError


def make(*a, **k):
    raise error.DependencyNotInstalled("gymnasium shim has no environments")
