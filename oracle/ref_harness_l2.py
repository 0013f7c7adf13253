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
