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
    ("optimized_walker.core", "DingPoint"), ("gym.optimized_walker.core", "DingPoint"),
    ("walker_gym_b200.engine", "DingPoint"), ("walker_gym_b200.optimized_walker.core", "DingPoint"),
}
_POINT_CLASSES.add(("walker_gym_b200.optimized_walker.core", "Point"))
_NUMPY_OK = {
    ("numpy._core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "_reconstruct"),
    ("numpy._core.multiarray", "scalar"), ("numpy.core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
}


class _SafeUnpickler(pickle.Unpickler):
    point_cls = ding_cls = None          # default: this package's flat-lineage descriptors

    def find_class(self, module, name):
        from .engine import DingPoint, Point
        if (module, name) in _POINT_CLASSES:
            return self.point_cls or Point
        if (module, name) in _DING_CLASSES:
            return self.ding_cls or DingPoint
        if (module, name) in _NUMPY_OK:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"state.pkl: class {module}.{name} is not allowed")


def load_points(path):
    """Return ``(points, r_points)`` from a reference-style snapshot."""
    state = _safe_load(path)
    if not isinstance(state, dict) or "points" not in state:
        raise ValueError("not a walker-gym snapshot: missing 'points'")
    pts = list(state["points"])
    for p in pts:
        for key in ("pos", "v", "a", "old_a"):
            setattr(p, key, np.asarray(getattr(p, key), dtype=np.float32).copy())
    return pts, dict(state.get("r_points", {}))


def _safe_load(path, point_cls=None, ding_cls=None):
    with open(path, "rb") as f:
        up = _SafeUnpickler(f)
        up.point_cls, up.ding_cls = point_cls, ding_cls
        return up.load()


def load_state_dict(path, point_cls=None, ding_cls=None):
    """The whole snapshot dict (the package lineage adds ``"fps"``, gym/optimized_walker/core.py:236-256)."""
    state = _safe_load(path, point_cls, ding_cls)
    if not isinstance(state, dict) or "points" not in state:
        raise ValueError("not a walker-gym snapshot: missing 'points'")
    for p in state["points"]:
        for key in ("pos", "v", "a", "old_a"):
            setattr(p, key, np.asarray(getattr(p, key), dtype=np.float32).copy())
    return state


_ENV_STATE_KEYS = ("points", "ding_points", "springs", "gravity", "damping", "ground", "ground_level",
                   "ground_restitution", "air_resistance", "friction", "time_step")


def load_env_state(path, point_cls=None, ding_cls=None):
    """``env_state.pkl`` of the package lineage's ``Environment.save_state``
    (gym/optimized_walker/env.py:262-281): points, ding_points, springs as ``(point1, point2, x, k,
    string)`` tuples that reference those point objects, and the constructor arguments."""
    state = _safe_load(path, point_cls, ding_cls)
    if not isinstance(state, dict) or any(k not in state for k in _ENV_STATE_KEYS):
        raise ValueError("not an env_state.pkl: missing keys")
    for p in list(state["points"]) + list(state["ding_points"]):
        for key in ("pos", "v", "a", "old_a"):
            setattr(p, key, np.asarray(getattr(p, key), dtype=np.float32).copy())
    return state


def _shim_classes(module):
    shim_mod = types.ModuleType(module)

    class Point:          # noqa: D401 - pickled by reference (module, qualname)
        pass

    class DingPoint(Point):
        pass

    Point.__module__ = DingPoint.__module__ = module
    Point.__qualname__, DingPoint.__qualname__ = "Point", "DingPoint"
    shim_mod.Point, shim_mod.DingPoint = Point, DingPoint
    return shim_mod


def _shim_object(shim_mod, p):
    o = shim_mod.DingPoint() if getattr(p, "fixed", False) else shim_mod.Point()
    for k in ("m", "pos", "v", "a", "r", "old_a", "color", "e"):
        v = getattr(p, k)
        o.__dict__[k] = np.array(v, dtype=np.float32) if isinstance(v, np.ndarray) else v
    if hasattr(p, "original_pos"):
        o.__dict__["original_pos"] = np.array(p.original_pos, dtype=np.float32)
    return o


def _dump_as(module, shim_mod, obj) -> bytes:
    """pickle ``obj`` while ``import <module>`` resolves to the shim (pickle looks classes up by path)."""
    saved = {}
    parts = module.split(".")
    try:
        for i in range(1, len(parts) + 1):
            name = ".".join(parts[:i])
            saved[name] = sys.modules.get(name)
            if i < len(parts):
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
            else:
                sys.modules[name] = shim_mod
        buf = io.BytesIO()
        pickle.dump(obj, buf, protocol=4)
    finally:
        for name, mod in saved.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod
    return buf.getvalue()


def save_env_state(path, env, module="optimized_walker.core"):
    """Write ``env`` (walker_gym_b200.optimized_walker.Environment) as the reference's ``env_state.pkl``:
    objects are pickled as ``<module>.Point`` / ``DingPoint`` so the reference's ``load_state`` reads them."""
    shim_mod = _shim_classes(module)
    objs = {id(p): _shim_object(shim_mod, p) for p in list(env.points) + list(env.ding_points)}
    state = {"points": [objs[id(p)] for p in env.points], "ding_points": [objs[id(p)] for p in env.ding_points],
             "springs": [(objs[id(a)], objs[id(b)], x, k, string) for a, b, x, k, string in env.springs],
             "gravity": np.array(env.gravity, dtype=np.float32)}
    for k in _ENV_STATE_KEYS[4:]:
        state[k] = getattr(env, k)
    with open(path, "wb") as f:
        f.write(_dump_as(module, shim_mod, state))


def save_points(path, points, r_points=None, module="gym.engine", extra=None):
    """Write a snapshot the reference can ``Point.backup``: objects are pickled
    as ``<module>.Point`` (default ``gym.engine.Point``, the class the shipped
    ``state.pkl`` was written from)."""
    shim_mod = _shim_classes(module)
    state = {"points": [_shim_object(shim_mod, p) for p in points], "r_points": dict(r_points or {})}
    state.update(extra or {})
    with open(path, "wb") as f:
        f.write(_dump_as(module, shim_mod, state))
