"""Test-only stand-in for `gymnasium` (not installed here). See oracle/shims/README.md."""
import sys as _sys
from gym import Env, Wrapper  # noqa: F401
from gym import spaces  # noqa: F401

_sys.modules[__name__ + ".spaces"] = spaces
