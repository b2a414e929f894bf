"""Test-only stand-in for `gym` (not installed here). See oracle/shims/README.md."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def reset(self, **kwargs):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode="human"):
        return None

    def close(self):
        return None


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)
