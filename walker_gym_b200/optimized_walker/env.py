"""``Environment`` of the reference's package lineage (gym/optimized_walker/env.py:8-305) on the B200
library: same constructor, ``add_point / add_ding_point / add_spring / batch_add_*``,
``update_physics``, ``run``, ``get_statistics``, ``save_state / load_state``.

``update_physics`` (env.py:135-184) runs in ``wg_pkg_update_physics`` (CUDA); there is no CPU path.
New: ``num_envs`` steps that many independent copies of the system per call, and
``update_physics(steps=n)`` / ``run(steps=n)`` keep the state on chip for all n updates of one launch.
With ``num_envs == 1`` the ``Point`` objects are the state, as in the reference: they are uploaded
before and refreshed after every call, so user code that reads or edits ``p.pos`` / ``p.v`` keeps working.
With ``num_envs > 1`` the device tensors ``pos`` / ``vel`` / ``old_a`` ([3*P, E], row n*3+c) are the state.
Rendering is out of scope: ``renderer`` is accepted and ignored.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import List

import numpy as np

from .. import _lib
from .core import Config, DingPoint, Point, to_data


class Environment:
    def __init__(self, gravity=(0, -9.8, 0), damping=0.99, ground=True, ground_level=-50, ground_restitution=0.8,
                 air_resistance=0.01, friction=0.5, time_step=0.01, renderer=None, *, num_envs=1, device="cuda:0"):
        self.gravity = to_data(gravity)
        self.damping = damping
        self.ground = ground
        self.ground_level = ground_level
        self.ground_restitution = ground_restitution
        self.air_resistance = air_resistance
        self.friction = friction
        self.time_step = time_step
        self.points: List[Point] = []
        self.ding_points: List[Point] = []
        self.springs = []            # (point1, point2, x, k, string)
        self.renderer = renderer
        self.running = False
        self.paused = False
        self.frame_count = 0
        self.start_time = 0
        self.last_time = 0
        if int(num_envs) < 1:
            raise ValueError("num_envs must be >= 1")
        self.num_envs = int(num_envs)
        self.device = device
        self._order: List[Point] = []       # creation order = row order of the device state
        self._sys = None
        self._state = None                  # dict(pos, vel, old_a) of torch tensors

    # ---- construction (env.py:56-133) -------------------------------------------------------------
    def add_point(self, m, pos, v=(0, 0, 0), r=None, color="black") -> Point:
        p = Point(m, pos, v, r, color)
        self.points.append(p)
        self._order.append(p)
        self._invalidate()
        return p

    def add_ding_point(self, m, pos, v=(0, 0, 0), r=None, color="red") -> DingPoint:
        p = DingPoint(m, pos, v, r, color)
        self.ding_points.append(p)
        self._order.append(p)
        self._invalidate()
        return p

    def add_spring(self, point1, point2, x=None, k=100, string=False) -> None:
        if x is None:
            x = np.linalg.norm(point1.pos - point2.pos).astype(Config.precision)
        self.springs.append((point1, point2, x, k, string))
        self._invalidate()

    def batch_add_points(self, points_data) -> List[Point]:
        return [self.add_point(**d) for d in points_data]

    def batch_add_springs(self, springs_data) -> None:
        for d in springs_data:
            self.add_spring(**d)

    # ---- device plumbing --------------------------------------------------------------------------
    def _invalidate(self):
        if self._state is not None and self.num_envs > 1:
            raise RuntimeError("the system of a batched Environment cannot change after the first update")
        self._sys = None
        self._state = None

    def system(self) -> "_lib.WgPkgSystem":
        """The ``wg_pkg_system`` of the current points and springs."""
        if self._sys is not None:
            return self._sys
        pts = self._order
        if not 1 <= len(pts) <= _lib.MAX_MASS:
            raise ValueError(f"an Environment needs 1..{_lib.MAX_MASS} points, got {len(pts)}")
        if len(self.springs) > _lib.MAX_SPRING:
            raise ValueError(f"at most {_lib.MAX_SPRING} springs are supported, got {len(self.springs)}")
        index = {id(p): n for n, p in enumerate(pts)}
        s = _lib.WgPkgSystem()
        s.n_point, s.n_spring = len(pts), len(self.springs)
        for n, p in enumerate(pts):
            s.mass[n] = float(p.m)
            s.fixed[n] = 1 if p.fixed else 0
        for q, (p1, p2, x, k, string) in enumerate(self.springs):
            try:
                s.si[q], s.sj[q] = index[id(p1)], index[id(p2)]
            except KeyError:
                raise ValueError("a spring references a point that is not in this Environment") from None
            s.srest[q], s.sk[q], s.sstring[q] = np.float32(x), np.float32(k), 1 if string else 0
        self._sys = s
        return s

    def params(self) -> "_lib.WgPkgParams":
        """The ``wg_pkg_params`` of the current attribute values (float32 at the point of use)."""
        p = _lib.WgPkgParams()
        g = to_data(self.gravity)
        for c in range(3):
            p.gravity[c] = g[c]
        p.damping = np.float32(self.damping)
        p.drag_c = np.float32(-0.5 * self.air_resistance)
        p.ground_level = np.float32(self.ground_level)
        p.restitution = np.float32(self.ground_restitution)
        p.friction = np.float32(self.friction)
        p.dt = np.float32(self.time_step)
        p.min_dist = np.float32(Config.r)
        p.ground = 1 if self.ground else 0
        return p

    def _host_rows(self):
        pos = np.array([p.pos for p in self._order], np.float32).reshape(-1, 1)
        vel = np.array([p.v for p in self._order], np.float32).reshape(-1, 1)
        return pos, vel

    def _ensure_state(self):
        import torch
        if self._state is None:
            pos, vel = self._host_rows()
            E = self.num_envs
            dev = torch.device(self.device)
            self._state = dict(pos=torch.from_numpy(np.repeat(pos, E, 1)).to(dev),
                               vel=torch.from_numpy(np.repeat(vel, E, 1)).to(dev),
                               old_a=torch.zeros((pos.shape[0], E), dtype=torch.float32, device=dev))
        elif self.num_envs == 1:            # the Point objects are the state: pick up user edits
            pos, vel = self._host_rows()
            self._state["pos"].copy_(torch.from_numpy(pos))
            self._state["vel"].copy_(torch.from_numpy(vel))
        return self._state

    @property
    def pos(self):
        """Device positions [3*P, num_envs] (row n*3+c, points in creation order)."""
        return self._ensure_state()["pos"]

    @property
    def vel(self):
        return self._ensure_state()["vel"]

    @property
    def old_a(self):
        return self._ensure_state()["old_a"]

    def sync_points(self, env_index: int = 0) -> None:
        """Copy env ``env_index`` of the device state into the ``Point`` objects."""
        st = self._ensure_state() if self._state is None else self._state
        pos = st["pos"][:, env_index].cpu().numpy().reshape(-1, 3)
        vel = st["vel"][:, env_index].cpu().numpy().reshape(-1, 3)
        oa = st["old_a"][:, env_index].cpu().numpy().reshape(-1, 3)
        for n, p in enumerate(self._order):
            p.pos[:], p.v[:] = pos[n], vel[n]
            p.old_a = oa[n].copy()
            p.a[:] = 0.0

    # ---- the hot path (env.py:135-184) -------------------------------------------------------------
    def update_physics(self, steps: int = 1) -> None:
        if not self.points:                 # env.py:137-138
            return
        import torch
        lib = _lib.load()
        st = self._ensure_state()
        sysm, prm = self.system(), self.params()
        stream = C.c_void_p(torch.cuda.current_stream(torch.device(self.device)).cuda_stream)
        with torch.cuda.device(torch.device(self.device)):
            rc = lib.wg_pkg_update_physics(C.byref(sysm), C.byref(prm), st["pos"].data_ptr(), st["vel"].data_ptr(),
                                           st["old_a"].data_ptr(), self.num_envs, int(steps), stream)
        _lib.check(rc, "wg_pkg_update_physics")
        self.frame_count += int(steps)
        Point.fps += int(steps)             # Point.run1 counts frames (core.py:200)
        if self.num_envs == 1:
            self.sync_points(0)

    def update(self) -> None:
        if self.running and not self.paused:
            self.update_physics()

    def run(self, steps: int = None, real_time: bool = True) -> None:
        """env.py:198-225 without the window: ``steps`` updates in ONE launch (``steps=None`` waited for the
        user to close the window in the reference; without a renderer that has no meaning here)."""
        if steps is None:
            raise ValueError("run(steps=None) needs the reference's interactive renderer, which is out of scope")
        self.running = True
        self.start_time = time.time()
        self.last_time = self.start_time
        if not self.paused:
            self.update_physics(int(steps))

    def pause(self) -> None:
        self.paused = True

    def resume(self) -> None:
        self.paused = False

    def stop(self) -> None:
        self.running = False

    def get_statistics(self) -> dict:
        elapsed = time.time() - self.start_time
        return {"frame_count": self.frame_count, "elapsed_time": elapsed,
                "avg_fps": self.frame_count / elapsed if elapsed > 0 else 0,
                "point_count": len(self.points) + len(self.ding_points), "spring_count": len(self.springs),
                "time_step": self.time_step}

    # ---- env_state.pkl (env.py:262-305) --------------------------------------------------------------
    def save_state(self, path: str = "env_state.pkl", env_index: int = 0) -> None:
        """Write env ``env_index`` in the reference's ``env_state.pkl`` schema (protocol 4)."""
        from ..state_io import save_env_state
        if self._state is not None:
            self.sync_points(env_index)
        save_env_state(path, self)

    def load_state(self, path: str = "env_state.pkl") -> None:
        """Read an ``env_state.pkl`` (written by the reference or by ``save_state``) through the allow-listed
        unpickler; every env of a batched Environment starts from the loaded state."""
        from ..state_io import load_env_state
        state = load_env_state(path, point_cls=Point, ding_cls=DingPoint)
        self._state = None
        self._sys = None
        self.points = list(state["points"])
        self.ding_points = list(state["ding_points"])
        self.springs = [tuple(s) for s in state["springs"]]
        self._order = self.points + self.ding_points
        self.gravity = to_data(np.asarray(state["gravity"]))
        for k in ("damping", "ground", "ground_level", "ground_restitution", "air_resistance", "friction", "time_step"):
            setattr(self, k, state[k])


class OptimizedEnvironment(Environment):
    """env.py:307-425.  The reference's subclass adds a spatial hash that it rebuilds every update and that
    nothing on the physics path reads (``detect_collisions`` has no caller): the physics is identical, so
    this is the same Environment with the reference's attributes."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.spatial_partition_size = 50
        self.spatial_partitions = {}
        self.enable_spatial_partitioning = True
        self.collision_margin = 1.0
        self.enable_parallel = True
