"""``Muscle`` / ``Skeleton`` / ``Creature`` / ``Brain`` and the eight bodies of
gym/optimized_walker/walker.py, as builders on top of the device-backed ``Environment``.

In the reference a muscle is a sinusoidal pattern generator that pushes its two points
(walker.py:56-90), but ``Environment.update_physics`` zeroes every acceleration before it applies
its own forces (env.py:141-142), so muscle forces never reach the dynamics.  This module keeps that
behaviour: ``Muscle.act`` advances ``t`` / ``state`` and books the force on the descriptors' pending
``a`` (visible until the next update, exactly as in the reference); the dynamics are the springs.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np

from .core import Config, Point
from .env import Environment


class Muscle:
    def __init__(self, point1, point2, amp=1.0, freq=1.0, phase=0.0, power=100.0, x=None):
        self.point1, self.point2 = point1, point2
        self.amp, self.freq, self.phase, self.power = amp, freq, phase, power
        self.x = np.linalg.norm(point1.pos - point2.pos).astype(Config.precision) if x is None else x
        self.t = 0
        self.state = 0
        self.active = True

    def __repr__(self):
        return f"Muscle(amp={self.amp}, freq={self.freq}, phase={self.phase}, power={self.power})"

    def params(self) -> dict:
        return {"amp": self.amp, "freq": self.freq, "phase": self.phase, "power": self.power, "x": self.x,
                "t": self.t, "state": self.state, "active": self.active}

    def _push(self) -> None:
        """walker.py:71-88: spring-like pull towards the target length, booked on the pending accelerations."""
        target = self.x * (1 - self.amp * self.state)
        current = np.linalg.norm(self.point1.pos - self.point2.pos).astype(Config.precision)
        magnitude = (target - current) * self.power
        direction = self.point2.pos - self.point1.pos
        length = np.linalg.norm(direction).astype(Config.precision)
        if length > Config.r:
            force = magnitude * (direction / length)
            self.point1.forced(force)
            self.point2.forced(-force)

    def act(self, dt) -> float:
        if not self.active:
            return self.state
        self.t += dt
        self.state = (np.sin(2 * np.pi * self.freq * self.t + self.phase) + 1) / 2
        self._push()
        return self.state

    def actdisp(self, dt, disp) -> float:
        if not self.active:
            return self.state
        self.t += dt
        self.state = np.clip(disp, 0, 1)
        self._push()
        return self.state

    def run(self, dt) -> None:
        self.act(dt)

    def toggle(self) -> None:
        self.active = not self.active

    def set_params(self, **kwargs) -> None:
        for key, value in kwargs.items():
            if hasattr(self, key):
                setattr(self, key, value)


class Skeleton:
    """Builder that registers points and springs with the Environment (walker.py:144-219)."""

    def __init__(self, env: Environment):
        self.env = env
        self.points: List[Point] = []
        self.springs = []
        self.muscles: List[Muscle] = []

    def add_point(self, m, pos, v=(0, 0, 0), r=None, color="black", is_ding=False) -> Point:
        p = self.env.add_ding_point(m, pos, v, r, color) if is_ding else self.env.add_point(m, pos, v, r, color)
        self.points.append(p)
        return p

    def add_spring(self, point1, point2, k=100, x=None, string=False) -> None:
        self.env.add_spring(point1, point2, x, k, string)
        self.springs.append((point1, point2))

    def add_muscle(self, point1, point2, amp=1.0, freq=1.0, phase=0.0, power=100.0, x=None) -> Muscle:
        m = Muscle(point1, point2, amp, freq, phase, power, x)
        self.muscles.append(m)
        return m

    def update(self, dt) -> None:
        for m in self.muscles:
            m.run(dt)


class Creature:
    def __init__(self, env: Environment, skeleton: Optional[Skeleton] = None):
        self.env = env
        self.skeleton = Skeleton(env) if skeleton is None else skeleton
        self.brain = None
        self.fitness = 0.0
        self.age = 0

    def __repr__(self):
        return f"Creature(fitness={self.fitness}, age={self.age})"

    def act(self, dt) -> None:
        self.skeleton.update(dt)
        if self.brain is not None:
            self.brain.control(self.skeleton.muscles, dt)
        self.age += 1

    def actdisp(self, dt, disp) -> None:
        n = len(self.skeleton.muscles)
        disp = list(disp)[:n] + [0.0] * max(0, n - len(disp))
        for muscle, d in zip(self.skeleton.muscles, disp):
            muscle.actdisp(dt, d)
        self.age += 1

    def evaluate_fitness(self) -> float:
        """x of the centre of mass (walker.py:275-300), from the Point descriptors (env 0)."""
        pts = self.skeleton.points
        if not pts:
            return 0.0
        total = 0.0
        com = np.zeros(3, dtype=Config.precision)
        for p in pts:
            total += p.m
            com += p.pos * p.m
        if total > 0:
            com /= total
        self.fitness = com[0]
        return self.fitness

    def evaluate_fitness_batched(self):
        """Same quantity for every env of a batched Environment, as a device tensor [num_envs]."""
        import torch
        env = self.env
        rows = [env._order.index(p) * 3 for p in self.skeleton.points]
        m = torch.tensor([float(p.m) for p in self.skeleton.points], dtype=torch.float32, device=env.pos.device)
        x = env.pos[rows, :]
        total = float(m.sum())
        return (x * m[:, None]).sum(0) / total if total > 0 else (x * m[:, None]).sum(0)

    def set_brain(self, brain: Callable) -> None:
        self.brain = brain


class Brain:
    def __init__(self, pattern: List[Dict] = None):
        self.pattern = pattern or []
        self.t = 0

    def control(self, muscles, dt) -> None:
        self.t += dt
        if self.pattern and len(self.pattern) >= len(muscles):
            for muscle, pat in zip(muscles, self.pattern):
                for key in ("amp", "freq", "phase", "power"):
                    if key in pat:
                        setattr(muscle, key, pat[key])


# ---- bodies (walker.py:356-639), as data ------------------------------------------------------------
def _build(env, points, springs, muscles=()) -> Creature:
    sk = Skeleton(env)
    pts = [sk.add_point(m, pos, **kw) for m, pos, kw in points]
    for i, j, k in springs:
        sk.add_spring(pts[i], pts[j], k=k)
    for i, j, kw in muscles:
        sk.add_muscle(pts[i], pts[j], **kw)
    return Creature(env, sk)


def test(env) -> Creature:
    c = _build(env, [(1, (0, 0, 0), {}), (1, (10, 0, 0), {})], [])
    sk = c.skeleton
    sk.add_spring(sk.points[0], sk.points[1])
    sk.add_muscle(sk.points[0], sk.points[1], amp=0.1, freq=1)
    return c


def leg2(env) -> Creature:
    pts = [(5, (0, 10, 0), dict(r=3)), (1, (-5, 5, 0), {}), (1, (-5, -5, 0), {}), (2, (-5, -15, 0), dict(r=2)),
           (1, (5, 5, 0), {}), (1, (5, -5, 0), {}), (2, (5, -15, 0), dict(r=2))]
    springs = [(0, 1, 500), (1, 2, 300), (2, 3, 300), (0, 4, 500), (4, 5, 300), (5, 6, 300)]
    mus = [(1, 2, dict(amp=0.1, freq=0.5, phase=0, power=200)), (2, 3, dict(amp=0.1, freq=0.5, phase=0.5, power=200)),
           (4, 5, dict(amp=0.1, freq=0.5, phase=0.5, power=200)), (5, 6, dict(amp=0.1, freq=0.5, phase=0, power=200))]
    return _build(env, pts, springs, mus)


def box(env, size: float = 10, mass: float = 1) -> Creature:
    h = size / 2
    corners = [(-h, h, -h), (h, h, -h), (h, -h, -h), (-h, -h, -h), (-h, h, h), (h, h, h), (h, -h, h), (-h, -h, h)]
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    return _build(env, [(mass, c, {}) for c in corners], [(i, j, 500) for i, j in edges])


def _pendulum(env, bobs) -> Creature:
    pts = [(0, (0, 20, 0), dict(is_ding=True, color="red"))] + bobs
    return _build(env, pts, [(n, n + 1, 200) for n in range(len(bobs))])


def balance1(env) -> Creature:
    return _pendulum(env, [(5, (0, 0, 0), dict(r=3))])


def balance2(env) -> Creature:
    return _pendulum(env, [(2, (0, 10, 0), {}), (2, (0, 0, 0), dict(r=2))])


def balance3(env) -> Creature:
    return _pendulum(env, [(1.5, (0, 15, 0), {}), (1.5, (0, 10, 0), {}), (1.5, (0, 0, 0), dict(r=2))])


def humanb(env) -> Creature:
    pts = [(3, (0, 30, 0), dict(r=3, color="blue")), (10, (0, 20, 0), dict(r=4)),
           (2, (-8, 25, 0), {}), (1, (-15, 20, 0), {}), (1, (-20, 20, 0), {}),
           (2, (8, 25, 0), {}), (1, (15, 20, 0), {}), (1, (20, 20, 0), {}),
           (2, (-5, 10, 0), {}), (1, (-5, 0, 0), {}), (2, (-5, -10, 0), dict(r=2)),
           (2, (5, 10, 0), {}), (1, (5, 0, 0), {}), (2, (5, -10, 0), dict(r=2))]
    springs = [(0, 1, 500), (1, 2, 400), (2, 3, 300), (3, 4, 200), (1, 5, 400), (5, 6, 300), (6, 7, 200),
               (1, 8, 500), (8, 9, 400), (9, 10, 400), (1, 11, 500), (11, 12, 400), (12, 13, 400)]
    mk = lambda f, ph, pw: dict(amp=0.1, freq=f, phase=ph, power=pw)     # noqa: E731
    mus = [(1, 3, mk(0.3, 0, 150)), (2, 4, mk(0.3, 0.5, 100)), (1, 6, mk(0.3, 0.5, 150)), (5, 7, mk(0.3, 0, 100)),
           (1, 9, mk(0.5, 0, 200)), (8, 10, mk(0.5, 0.5, 150)), (1, 12, mk(0.5, 0.5, 200)), (11, 13, mk(0.5, 0, 150))]
    return _build(env, pts, springs, mus)


def insect(env, legs: int = 6) -> Creature:
    sk = Skeleton(env)
    half = legs // 2
    length = legs * 5
    spine = []
    for i in range(half):
        x = -length / 2 + i * (length / (half - 1)) if legs > 2 else 0
        spine.append(sk.add_point(2, (x, 5, 0), r=2))
    for a, b in zip(spine, spine[1:]):
        sk.add_spring(a, b, k=400)
    for i, bp in enumerate(spine):
        bx = bp.pos[0]
        sides = []
        for sgn in (-1, +1):        # left leg, then right leg; per leg: upper, lower, foot
            sides.append([sk.add_point(1, (bx + sgn * 5, 0, 0)), sk.add_point(1, (bx + sgn * 10, -5, 0)),
                          sk.add_point(1, (bx + sgn * 15, -10, 0), r=1.5)])
        for up, lo, ft in sides:
            sk.add_spring(bp, up, k=300)
            sk.add_spring(up, lo, k=200)
            sk.add_spring(lo, ft, k=200)
        ph = i * (np.pi / half)
        (lu, ll, lf), (ru, rl, rf) = sides
        sk.add_muscle(bp, ll, amp=0.1, freq=0.8, phase=ph, power=100)
        sk.add_muscle(lu, lf, amp=0.1, freq=0.8, phase=ph + 0.5, power=80)
        sk.add_muscle(bp, rl, amp=0.1, freq=0.8, phase=ph + np.pi, power=100)
        sk.add_muscle(ru, rf, amp=0.1, freq=0.8, phase=ph + np.pi + 0.5, power=80)
    return Creature(env, sk)


export = {"Muscle": Muscle, "Skeleton": Skeleton, "Creature": Creature, "Brain": Brain, "test": test, "leg2": leg2,
          "box": box, "balance1": balance1, "balance2": balance2, "balance3": balance3, "humanb": humanb,
          "insect": insect}
