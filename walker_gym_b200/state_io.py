"""``state.pkl`` compatibility (gym/engine.py:199-212, gym/optimized_engine.py:319-336).

The file is a protocol-4 pickle ``{"points": [obj, ...], "r_points": {...}}``
whose objects are instances of the *writer's* Point class with ``__dict__``
keys ``m, pos, v, a, r, old_a, color, e`` and float32[3] arrays.  Loading uses
an allow-listed ``Unpickler`` (no arbitrary code execution) that maps every
known writer class onto this package's ``Point``; saving emits objects under
the module path ``gym.engine`` / class ``Point`` so the reference's
``Point.backup`` can read them back.
"""
from __future__ import annotations

import io
import pickle
import sys
import types

import numpy as np

_POINT_CLASSES = {
    ("gym.engine", "Point"), ("engine", "Point"),
    ("optimized_engine", "Point"), ("gym.optimized_engine", "Point"),
    ("optimized_walker.core", "Point"), ("gym.optimized_walker.core", "Point"),
    ("walker_gym_b200.engine", "Point"),
}
_DING_CLASSES = {
    ("gym.engine", "DingPoint"), ("engine", "DingPoint"),
    ("optimized_engine", "DingPoint"), ("gym.optimized_engine", "DingPoint"),
    ("optimized_walker.core", "DingPoint"), ("walker_gym_b200.engine", "DingPoint"),
}
_NUMPY_OK = {
    ("numpy._core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "_reconstruct"),
    ("numpy._core.multiarray", "scalar"), ("numpy.core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
}


class _SafeUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        from .engine import DingPoint, Point
        if (module, name) in _POINT_CLASSES:
            return Point
        if (module, name) in _DING_CLASSES:
            return DingPoint
        if (module, name) in _NUMPY_OK:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"state.pkl: class {module}.{name} is not allowed")


def load_points(path):
    """Return ``(points, r_points)`` from a reference-style snapshot."""
    with open(path, "rb") as f:
        state = _SafeUnpickler(f).load()
    if not isinstance(state, dict) or "points" not in state:
        raise ValueError("not a walker-gym snapshot: missing 'points'")
    pts = list(state["points"])
    for p in pts:
        for key in ("pos", "v", "a", "old_a"):
            setattr(p, key, np.asarray(getattr(p, key), dtype=np.float32).copy())
    return pts, dict(state.get("r_points", {}))


def save_points(path, points, r_points=None, module="gym.engine"):
    """Write a snapshot the reference can ``Point.backup``: objects are pickled
    as ``<module>.Point`` (default ``gym.engine.Point``, the class the shipped
    ``state.pkl`` was written from)."""
    shim_mod = types.ModuleType(module)
    fields = ("m", "pos", "v", "a", "r", "old_a", "color", "e")

    class Point:          # noqa: D401 - pickled by reference (module, qualname)
        pass

    class DingPoint(Point):
        pass

    Point.__module__ = DingPoint.__module__ = module
    Point.__qualname__, DingPoint.__qualname__ = "Point", "DingPoint"
    shim_mod.Point, shim_mod.DingPoint = Point, DingPoint
    objs = []
    for p in points:
        o = DingPoint() if getattr(p, "fixed", False) else Point()
        for k in fields:
            v = getattr(p, k)
            o.__dict__[k] = np.array(v, dtype=np.float32) if isinstance(v, np.ndarray) else v
        if getattr(p, "fixed", False):
            o.__dict__["original_pos"] = np.array(p.original_pos, dtype=np.float32)
        objs.append(o)
    saved = {}
    parts = module.split(".")
    try:
        for i in range(1, len(parts) + 1):      # make `import gym.engine` resolvable for pickle's lookup
            name = ".".join(parts[:i])
            saved[name] = sys.modules.get(name)
            if i < len(parts):
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
            else:
                sys.modules[name] = shim_mod
        buf = io.BytesIO()
        pickle.dump({"points": objs, "r_points": dict(r_points or {})}, buf, protocol=4)
    finally:
        for name, mod in saved.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod
    with open(path, "wb") as f:
        f.write(buf.getvalue())
