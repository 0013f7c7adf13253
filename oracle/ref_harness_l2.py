"""TEST INFRASTRUCTURE ONLY -- executes the reference's *package* lineage ("L2",
``gym/optimized_walker/{core,env}.py``) unmodified, for golden vectors of its
``Environment.update_physics`` (gym/optimized_walker/env.py:135-184).

Only ``pygame`` is stubbed (the package's renderer imports it).  The package is
imported under a private name so it cannot collide with the flat module
``optimized_walker.py`` that ``ref_harness.py`` loads under the same name.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from unittest import mock

import numpy as np

DEFAULT_REF = os.environ.get("WALKER_GYM_REFERENCE", "/root/reference")
_PKG = "_wg_ref_l2pkg"
_loaded = {}


def available(ref_root: str = DEFAULT_REF) -> bool:
    return os.path.isfile(os.path.join(ref_root, "gym", "optimized_walker", "env.py"))


def load(ref_root: str = DEFAULT_REF):
    """Return (core, env) modules of the reference package."""
    if ref_root in _loaded:
        return _loaded[ref_root]
    pkg_dir = os.path.join(ref_root, "gym", "optimized_walker")
    if not available(ref_root):
        raise FileNotFoundError(pkg_dir)
    if "pygame" not in sys.modules:
        sys.modules["pygame"] = mock.MagicMock(name="pygame")
    spec = importlib.util.spec_from_file_location(_PKG, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules[_PKG] = pkg
    spec.loader.exec_module(pkg)
    core, env = sys.modules[_PKG + ".core"], sys.modules[_PKG + ".env"]
    _loaded[ref_root] = (core, env)
    return core, env


def rollout(system, steps, *, env_kwargs=None, ref_root: str = DEFAULT_REF):
    """Build the reference Environment from ``system`` and call update_physics ``steps`` times.

    system: {"points": [(m, pos, vel, ding)], "springs": [(i, j, x_or_None, k, string)]}
    Returns pos / vel / old_a [steps+1, P, 3] (points in the order given) and the spring rest lengths."""
    core, envmod = load(ref_root)
    core.Point.points = []
    core.Point.r_points = {}
    env = envmod.Environment(**dict(env_kwargs or {}))
    pts = []
    for m, pos, vel, ding in system["points"]:
        pts.append(env.add_ding_point(m, list(pos), list(vel)) if ding else env.add_point(m, list(pos), list(vel)))
    for i, j, x, k, string in system["springs"]:
        env.add_spring(pts[i], pts[j], x, k, string)
    rest = np.array([float(s[2]) for s in env.springs], dtype=np.float64)

    def snap():
        return (np.array([p.pos for p in pts], np.float32), np.array([p.v for p in pts], np.float32),
                np.array([p.old_a for p in pts], np.float32))

    frames = [snap()]
    for _ in range(steps):
        env.update_physics()
        frames.append(snap())
    core.Point.points = []
    core.Point.r_points = {}
    return dict(pos=np.stack([f[0] for f in frames]), vel=np.stack([f[1] for f in frames]),
                old_a=np.stack([f[2] for f in frames]), rest=rest)


def body_rollout(name, steps, *, env_kwargs=None, ref_root: str = DEFAULT_REF):
    """Build body ``name`` with the reference's own builder (gym/optimized_walker/walker.py:356-639) in a default
    Environment, record its tables and ``steps`` update_physics calls (points in env.points + env.ding_points
    creation order as registered in Point.points)."""
    core, envmod = load(ref_root)
    walker = sys.modules[_PKG + ".walker"]
    core.Point.points = []
    core.Point.r_points = {}
    env = envmod.Environment(**dict(env_kwargs or {}))
    creature = getattr(walker, name)(env)
    pts = list(core.Point.points)                       # creation order
    index = {id(p): n for n, p in enumerate(pts)}
    rec = dict(mass=np.array([float(p.m) for p in pts]), ding=np.array([isinstance(p, core.DingPoint) for p in pts]),
               pos0=np.array([p.pos for p in pts], np.float32),
               si=np.array([index[id(s[0])] for s in env.springs], np.int32),
               sj=np.array([index[id(s[1])] for s in env.springs], np.int32),
               sx=np.array([np.float32(s[2]) for s in env.springs], np.float32),
               sk=np.array([float(s[3]) for s in env.springs]), sstring=np.array([bool(s[4]) for s in env.springs]),
               muscle_i=np.array([index[id(m.point1)] for m in creature.skeleton.muscles], np.int32),
               muscle_j=np.array([index[id(m.point2)] for m in creature.skeleton.muscles], np.int32),
               muscle_x=np.array([np.float32(m.x) for m in creature.skeleton.muscles], np.float32),
               muscle_par=np.array([[m.amp, m.freq, m.phase, m.power] for m in creature.skeleton.muscles], np.float64).reshape(-1, 4))
    frames_p, frames_v = [rec["pos0"].copy()], [np.array([p.v for p in pts], np.float32)]
    states = []
    for _ in range(steps):
        creature.act(env.time_step)                      # muscles push; update_physics wipes it (env.py:141-142)
        states.append([float(m.state) for m in creature.skeleton.muscles])
        env.update_physics()
        frames_p.append(np.array([p.pos for p in pts], np.float32))
        frames_v.append(np.array([p.v for p in pts], np.float32))
    rec["pos"], rec["vel"] = np.stack(frames_p), np.stack(frames_v)
    rec["muscle_state"] = np.array(states, np.float64).reshape(steps, -1)
    core.Point.points = []
    core.Point.r_points = {}
    return rec
