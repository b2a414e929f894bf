"""Observation / action spaces of the continuous env (cyberbattle_env_compressed.py:112-142).  Uses
gymnasium when it is installed (what Stable-Baselines3 expects); otherwise a minimal stand-in with the
same attributes, so that the package imports on machines without gymnasium."""
from __future__ import annotations

import numpy as np

from . import constants as C

try:  # pragma: no cover - depends on the environment
    from gymnasium import spaces as _spaces
    Box, Dict = _spaces.Box, _spaces.Dict
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape, dtype=np.float32):
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}


def action_space():
    """Box(-4, 4, (905,), float32) — compressed:112-114"""
    return Box(low=-4.0, high=4.0, shape=(C.ACTION_DIM,), dtype=np.float32)


def observation_space(graph_dim: int = C.OBS_DIM):
    """Dict(graph_embeddings Box(-16,16,(192,)) [(256,) for *_node goals], discrete_features Box(0,300,(2,))) —
    compressed:119-142"""
    return Dict({"graph_embeddings": Box(low=-16.0, high=16.0, shape=(graph_dim,), dtype=np.float64),
                 "discrete_features": Box(low=0.0, high=300.0, shape=(2,), dtype=np.float64)})
