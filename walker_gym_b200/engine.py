"""Host-side descriptors with the constructor surface of ``gym/optimized_engine.py``.

``Point`` / ``DingPoint`` here are *descriptors*: they carry the numbers a user
passes to the reference's constructors (``Point(m, pos, v, r, color, e)``,
gym/optimized_engine.py:42-68) and mirror the state the device computes, so
code written against the reference keeps reading ``p.pos``, ``p.v``,
``p.old_a``, ``p.r`` and ``p.color``.  The physics itself never runs here --
it runs in the CUDA library (see ``env.py``).
"""
from __future__ import annotations

import pickle
from typing import Dict, List

import numpy as np


class Config:
    """Engine constants (gym/optimized_engine.py:5-10)."""
    precision = np.float32
    r = 16e-36
    e = 16e-20
    k = 8.99e9
    g = 9.8


class Point:
    """A point mass.  Same attributes as the reference's ``Point`` objects
    (``m, pos, v, a, r, old_a, color, e``), which is also the schema of ``state.pkl``."""

    points: List["Point"] = []      # process-wide registry, as in the reference (:16)
    r_points: Dict = {}             # part of the snapshot schema (:17)

    fixed = False                   # True for DingPoint

    def __init__(self, m, pos, v, r=None, color="black", e=Config.e):
        self.m = m
        self.pos = np.array(pos, dtype=Config.precision)
        self.v = np.array(v, dtype=Config.precision)
        self.a = np.zeros(3, dtype=Config.precision)
        self.r = m ** 0.3 if r is None else r
        self.old_a = np.zeros(3, dtype=Config.precision)
        self.color = color
        self.e = e
        if self.pos.shape != (3,) or self.v.shape != (3,):
            raise TypeError("pos and v must be 3-vectors")
        Point.points.append(self)

    def __repr__(self):
        return f"Point(m={self.m}, pos={self.pos}, v={self.v}, a={self.old_a})"

    def params(self):
        return {"m": self.m, "v": self.v.tolist(), "a": self.a.tolist(), "pos": self.pos.tolist(),
                "r": self.r, "e": self.e, "color": self.color, "old_a": self.old_a.tolist()}

    def zero(self) -> None:
        """Clear the pending acceleration (gym/optimized_engine.py:100-102)."""
        self.a = np.zeros(3, dtype=Config.precision)

    @classmethod
    def clear(cls) -> None:
        """Forget every registered point (gym/optimized_engine.py:28-40)."""
        Point.points = []
        Point.r_points = {}

    @classmethod
    def snapshot(cls, path="state.pkl") -> None:
        """Write ``{"points": [...], "r_points": {...}}`` with pickle protocol 4
        (gym/optimized_engine.py:319-324, gym/engine.py:199-204)."""
        from .state_io import save_points
        save_points(path, Point.points, Point.r_points)

    @classmethod
    def backup(cls, path="state.pkl") -> None:
        """Restore the registry from a snapshot (gym/optimized_engine.py:326-336).
        Reads files written by the reference (``gym.engine.Point``, ``optimized_engine.Point``,
        ``optimized_walker.core.Point``) through an allow-listed unpickler."""
        from .state_io import load_points
        pts, rp = load_points(path)
        Point.points = pts
        Point.r_points = rp


class DingPoint(Point):
    """A pinned point: forces on it are ignored (gym/optimized_engine.py:404-416).
    Note the reference still integrates its velocity, so reset jitter makes it drift;
    the device reproduces that."""

    fixed = True

    def __init__(self, m, p, v=None, r=None, color="black"):
        super().__init__(m, p, [0, 0, 0] if v is None else v, r, color)
        self.original_pos = self.pos.copy()
