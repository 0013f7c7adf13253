"""Descriptors with the constructor surface of ``gym/optimized_walker/core.py``.

``Point(m, pos, v, r, color, e)`` (core.py:35-58) and ``DingPoint`` (:259-275)
carry what the user passes in and mirror what the device computes (``pos``,
``v``, ``old_a``).  The physics never runs here: ``Environment.update_physics``
(``env.py``) runs it in the CUDA library.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np


class Config:
    """core.py:5-15."""
    precision = np.float32
    r = 16e-36
    e = 16e-20
    k = 8.99e9
    g = 9.8
    batch_size = 100


def to_data(data) -> np.ndarray:
    """core.py:18-25: tuples/lists/arrays become float32 arrays, anything else is a TypeError."""
    if isinstance(data, (tuple, list)):
        return np.array(data, dtype=Config.precision)
    if isinstance(data, np.ndarray):
        return data.astype(Config.precision)
    raise TypeError(f"Data must be a numpy array, tuple, list (not {type(data).__name__})")


class Point:
    points: List["Point"] = []      # process-wide registry (core.py:31), part of the snapshot schema
    r_points: Dict = {}
    fps = 0

    fixed = False

    def __init__(self, m, pos, v, r=None, color="black", e=Config.e):
        self.m = m
        self.pos = to_data(pos)
        self.v = to_data(v)
        if self.pos.shape != (3,) or self.v.shape != (3,):
            raise TypeError("pos and v must be 3-vectors")
        self.a = np.zeros(3, dtype=Config.precision)
        self.r = m ** 0.3 if r is None else r
        self.old_a = self.a.copy()
        self.color = color
        self.e = e
        Point.points.append(self)

    def __repr__(self):
        return f"Point(m={self.m}, pos={self.pos}, v={self.v}, a={self.old_a})"

    def params(self) -> dict:
        return {"m": self.m, "v": self.v.tolist(), "a": self.a.tolist(), "pos": self.pos.tolist(),
                "r": self.r, "e": self.e, "color": self.color, "old_a": self.old_a.tolist()}

    def zero(self) -> None:
        self.a[:] = 0.0

    def forced(self, f) -> None:
        """core.py:81-83 on the descriptor's pending acceleration.  Only bookkeeping: ``update_physics``
        zeroes ``a`` before it applies any force (env.py:141-142), so nothing added here reaches the dynamics
        -- in the reference either."""
        self.a += np.asarray(f) / self.m

    @classmethod
    def clear(cls) -> None:
        Point.points = []
        Point.r_points = {}
        Point.fps = 0

    @classmethod
    def snapshot(cls, path: str = "state.pkl") -> None:
        """core.py:236-246: ``{"points", "r_points", "fps"}``, pickle protocol 4."""
        from ..state_io import save_points
        save_points(path, Point.points, Point.r_points, module="optimized_walker.core", extra={"fps": Point.fps})

    @classmethod
    def load_snapshot(cls, path: str = "state.pkl") -> None:
        """core.py:248-256."""
        from ..state_io import load_state_dict
        state = load_state_dict(path, point_cls=Point, ding_cls=DingPoint)
        Point.points = list(state["points"])
        Point.r_points = dict(state.get("r_points", {}))
        Point.fps = state.get("fps", 0)


class DingPoint(Point):
    """A pinned point (core.py:259-275): ``forced`` and ``zero`` are no-ops; ``run1`` still integrates
    its velocity, so a DingPoint created with v != 0 drifts -- the device reproduces that."""
    fixed = True

    def __init__(self, m, pos, v, r=None, color="black"):
        super().__init__(m, pos, v, r, color)

    def forced(self, f) -> None:
        pass

    def zero(self) -> None:
        pass
