"""Morphology descriptors with the surface of ``gym/optimized_walker.py`` and the
body tables of ``gym/walker.py``.

``Muscle`` / ``Skeleton`` / ``Creature`` describe the spring-mass body; the
forces they stand for are evaluated on the device by the fused step kernel
(``csrc/wg_physics.cuh``), in the reference's order: all muscles in list
order, then all skeletons (gym/optimized_walker.py:117-127).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .engine import DingPoint, Point


def _distance(p1: Point, p2: Point):
    # np.linalg.norm of a float32 difference -> np.float32, as in Muscle.distant (:23-25)
    return np.linalg.norm(p1.pos - p2.pos)


class Muscle:
    """An actuated spring: its rest length ``x`` is the control variable
    (gym/optimized_walker.py:7-43)."""

    def __init__(self, p1: Point, p2: Point, x=None, k=1000, maxl=1.5, minl=0.1, stride=2, dampk=20, string=False):
        self.p1, self.p2 = p1, p2
        self.x = _distance(p1, p2) if x is None else x
        self.originx = self.x
        self.k, self.dampk = k, dampk
        self.minl, self.maxl = minl, maxl
        self.stride = stride
        # extension (keyword only in spirit; the reference's constructor ends at dampk): rope-type spring, no elastic
        # force while shorter than its rest length -- Point.resilience's `string` (gym/optimized_engine.py:116-140)
        self.string = bool(string)

    def distant(self, p1: Point, p2: Point):
        return _distance(p1, p2)

    def regulation(self) -> None:
        """Clamp ``x`` to ``[originx*minl, originx*maxl]`` with python max/min semantics (:27-30)."""
        self.x = max(self.x, self.originx * self.minl)
        self.x = min(self.x, self.originx * self.maxl)

    def act(self, a) -> None:
        """Continuous control: ``x += a`` then clamp (:32-35)."""
        self.x += a
        self.regulation()

    def actdisp(self, a) -> None:
        """Discrete control: lengthen or shorten by ``stride`` (:37-43)."""
        self.x = self.x + self.stride if a else self.x - self.stride
        self.regulation()

    def run(self) -> None:
        raise NotImplementedError("forces are evaluated on the device; call PhysicsEnv.step / Environment.step")


class Skeleton:
    """A passive spring (gym/optimized_walker.py:69-82)."""

    def __init__(self, p1: Point, p2: Point, x=None, k=1000, dampk=20, string=False):
        self.p1, self.p2 = p1, p2
        self.x = _distance(p1, p2) if x is None else x
        self.k, self.dampk = k, dampk
        self.string = bool(string)      # extension: rope-type (gym/optimized_engine.py:134-138), see Muscle

    def distant(self, p1: Point, p2: Point):
        return _distance(p1, p2)

    def run(self) -> None:
        raise NotImplementedError("forces are evaluated on the device; call PhysicsEnv.step / Environment.step")


class Creature:
    """Points + muscles + skeletons (gym/optimized_walker.py:108-172)."""

    def __init__(self, phylist: List[Point], musclelist: List[Muscle], skeletonlist: List[Skeleton]):
        self.phys = phylist
        self.muscles = musclelist
        self.skeletons = skeletonlist

    def run(self) -> None:
        raise NotImplementedError("forces are evaluated on the device; call PhysicsEnv.step / Environment.step")

    def getstat(self, in3d=True, pk=1, vk=1, ak=1, mk=1, midform=True, conmid=False) -> List[float]:
        """Flatten the mirrored state into the reference's observation list (:129-162).

        This only *formats* numbers the device already produced (positions
        relative to the centroid, velocities, last accelerations, muscle
        lengths); ``PhysicsEnv`` returns the kernel's own observation and uses
        this accessor only for non-default scale factors."""
        d = 3 if in3d else 2
        out: List[float] = []
        mid = np.zeros(3, dtype=np.float32)
        if midform:
            for p in self.phys:
                mid += p.pos
            mid /= len(self.phys)
        for p in self.phys:
            rel = (p.pos[:d] - mid[:d]) * pk if midform else p.pos[:d] * pk
            out.extend(rel.tolist())
            out.extend((p.v[:d] * vk).tolist())
            out.extend((p.old_a[:d] * ak).tolist())
        if conmid:
            out.extend(mid.tolist())
        out.extend(m.x * mk for m in self.muscles)
        return out

    def act(self, a: Sequence[float]) -> None:
        for m, ai in zip(self.muscles, a):
            m.act(ai)

    def actdisp(self, a: Sequence[bool]) -> None:
        for m, ai in zip(self.muscles, a):
            m.actdisp(ai)


# ---------------------------------------------------------------------------------
# Morphology tables.  (mass, position) per point; muscles / skeletons as (i, j, kwargs).
# ---------------------------------------------------------------------------------
def _build(points, muscles, skeletons, ding=()):
    pts = [DingPoint(m, list(p)) if n in ding else Point(m, list(p), [0, 0, 0]) for n, (m, p) in enumerate(points)]
    mus = [Muscle(pts[i], pts[j], **kw) for i, j, kw in muscles]
    sks = [Skeleton(pts[i], pts[j], **kw) for i, j, kw in skeletons]
    return Creature(pts, mus, sks)


def _pairs(*ij):
    return [(i, j, {}) for i, j in ij]


BODIES = {
    # gym/optimized_walker.py:176-199
    "balance_v0": dict(points=[(5, (-50, 100, 0)), (5, (50, 100, 0)), (1, (0, 0, 0)), (3, (0, 100, 0))],
                       muscles=_pairs((0, 2), (1, 2)), skeletons=_pairs((0, 1), (0, 3), (1, 3))),
    # gym/optimized_walker.py:201-224
    "box_v0": dict(points=[(1, (-50, 0, 0)), (1, (-50, 100, 0)), (1, (50, 100, 0)), (1, (50, 0, 0))],
                   muscles=_pairs((0, 1), (0, 2), (3, 1), (3, 2)), skeletons=_pairs((1, 2))),
    # ---- gym/walker.py:112-353 (legacy constructor order is Phy(m, v, p)) ----
    "test": dict(points=[(1, (-100, 100, 0)), (1, (100, 100, 0)), (1, (100, -100, 0)), (1, (-100, -100, 0))],
                 muscles=[(1, 2, {"stride": 3})],
                 skeletons=[(0, 1, {}), (0, 3, {}), (2, 3, {}), (0, 2, {"k": 100}), (1, 3, {"k": 100})]),
    "leg2": dict(points=[(1, (0, 100, 0)), (1, (100, 100, 0)), (1, (50, 50, 0)), (1, (100, 0, 0)),
                         (1, (-100, 100, 0)), (1, (-150, 50, 0)), (1, (-100, 0, 0))],
                 muscles=_pairs((1, 3), (4, 6), (0, 2), (0, 5)),
                 skeletons=_pairs((0, 1), (0, 4), (1, 4), (1, 2), (2, 3), (4, 5), (5, 6))),
    "box": dict(points=[(1, (-50, 0, 0)), (1, (-50, 100, 0)), (1, (50, 0, 0)), (1, (50, 100, 0))],
                muscles=_pairs((0, 2), (1, 3)), skeletons=_pairs((0, 1), (1, 2), (2, 3))),
    "box2": dict(points=[(1, (-50, 0, 0)), (1, (-50, 100, 0)), (1, (50, 100, 0)), (1, (50, 0, 0))],
                 muscles=_pairs((0, 1), (0, 2), (3, 1), (3, 2)), skeletons=_pairs((1, 2))),
    "balance": dict(points=[(1, (-50, 100, 0)), (1, (50, 100, 0)), (1, (0, 0, 0)), (1, (0, 100, 0))],
                    muscles=_pairs((0, 2), (1, 2)), skeletons=_pairs((0, 1), (0, 3), (1, 3))),
    "balance2": dict(points=[(5, (-50, 100, 0)), (5, (50, 100, 0)), (1, (0, 0, 0)), (0.1, (0, 100, 0))],
                     muscles=_pairs((0, 2), (1, 2)),
                     skeletons=[(0, 1, {}), (0, 3, {"k": 10000}), (1, 3, {"k": 10000})]),
    "balance3": dict(points=[(1, (-50, 100, 0)), (1, (50, 100, 0)), (1, (0, 0, 0)), (0.1, (0, 100, 0))],
                     muscles=_pairs((0, 2), (1, 2)),
                     skeletons=[(0, 1, {}), (0, 3, {"k": 20000}), (1, 3, {"k": 20000})], ding=(2,)),
    "intrian": dict(points=[(1, (-50, 100, 0)), (1, (50, 100, 0)), (1, (0, 0, 0))],
                    muscles=_pairs((0, 2), (1, 2), (0, 1)), skeletons=[]),
    "humanb": dict(points=[(1, (25, 250, 0)), (1, (-25, 200, 0)), (1, (25, 150, 0)), (1, (-25, 100, 0)),
                           (1, (25, 0, 0)), (1, (-25, 0, 0))],
                   muscles=_pairs((2, 4), (2, 5), (3, 4), (3, 5)),
                   skeletons=_pairs((0, 1), (0, 2), (1, 2), (1, 3), (2, 3))),
    "insect": dict(points=[(1, (-75, 100, 0)), (1, (-25, 100, 0)), (1, (25, 100, 0)), (1, (75, 100, 0)),
                           (1, (-100, 50, 0)), (1, (-50, 50, 0)), (1, (0, 50, 0)), (1, (50, 50, 0)), (1, (100, 50, 0)),
                           (1, (-75, 0, 0)), (1, (-25, 0, 0)), (1, (25, 0, 0)), (1, (75, 0, 0))],
                   muscles=_pairs((9, 4), (9, 5), (10, 5), (10, 6), (11, 6), (11, 7), (12, 7), (12, 8)),
                   skeletons=_pairs((0, 1), (0, 4), (0, 5), (1, 2), (1, 5), (1, 6), (2, 3), (2, 6), (2, 7),
                                    (3, 7), (3, 8), (4, 5), (5, 6), (6, 7), (7, 8))),
    "box4": dict(points=[(1, (-50, 100, 0)), (1, (50, 100, 0)), (1, (50, 0, 0)), (1, (17, 0, 0)),
                         (1, (-17, 0, 0)), (1, (-50, 0, 0))],
                 muscles=_pairs((0, 2), (0, 3), (0, 4), (0, 5), (1, 2), (1, 3), (1, 4), (1, 5)),
                 skeletons=_pairs((0, 1))),
    "leg": dict(points=[(1, (-50, 200, 0)), (1, (50, 200, 0)), (1, (-50, 140, 0)), (1, (50, 140, 0)),
                        (1, (-50, 70, 0)), (1, (50, 70, 0)), (1, (-50, 0, 0)), (1, (50, 0, 0))],
                muscles=_pairs((1, 3), (2, 4), (5, 7)),
                skeletons=_pairs((0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (3, 5), (4, 5), (4, 6), (5, 6), (6, 7))),
    "hat": dict(points=[(1, (0, 150, 0)), (1, (-50, 30, 0)), (1, (50, 30, 0)), (1, (-50, 0, 0)), (1, (50, 0, 0))],
                muscles=_pairs((1, 3), (1, 4), (2, 3), (2, 4)), skeletons=_pairs((0, 1), (0, 2), (1, 2))),
}


def _quad_balance():
    """Four Balance-v0 units side by side in one env: 4x the masses and springs
    (N=16, S=20, M=8) -- the enlarged morphology of BASELINE.json config 4."""
    base = BODIES["balance_v0"]
    pts, mus, sks = [], [], []
    for u in range(4):
        off = 150.0 * (u - 1.5)
        pts += [(m, (p[0] + off, p[1], p[2])) for m, p in base["points"]]
        mus += [(4 * u + i, 4 * u + j, {}) for i, j, _ in base["muscles"]]
    for u in range(4):
        sks += [(4 * u + i, 4 * u + j, {}) for i, j, _ in base["skeletons"]]
    return dict(points=pts, muscles=mus, skeletons=sks)


BODIES["quad_balance"] = _quad_balance()


def _quad_balance_chain():
    """The same four Balance-v0 units CONNECTED into one body: a bone from unit u's right shoulder (its point 1) to
    unit u+1's left shoulder (its point 0), rest length = the 50-unit gap.  N=16, S=23, M=8 -- BASELINE.json config 4's
    "enlarged walker morphology (4x masses/springs)" as one connected creature (the size of gym/walker.py's insect,
    :255-293, N=13, S=23, M=8).  The link bones come after every unit's own bones in the skeleton list."""
    spec = _quad_balance()
    spec["skeletons"] = spec["skeletons"] + [(4 * u + 1, 4 * (u + 1), {}) for u in range(3)]
    return spec


BODIES["quad_balance_chain"] = _quad_balance_chain()


def make_creature(name: str) -> Creature:
    """Build one of the in-tree morphologies by name (keys of ``BODIES``)."""
    spec = BODIES[name]
    return _build(spec["points"], spec["muscles"], spec["skeletons"], spec.get("ding", ()))


def create_balance_creature() -> Creature:
    """gym/optimized_walker.py:176-199"""
    return make_creature("balance_v0")


def create_box_creature() -> Creature:
    """gym/optimized_walker.py:201-224"""
    return make_creature("box_v0")


def _legacy(name):
    def ctor() -> Creature:
        return make_creature(name)
    ctor.__name__ = name
    ctor.__doc__ = f"gym/walker.py `{name}` morphology (data transcribed; legacy Phy(m, v, p) order)."
    return ctor


test, leg2, box, box2, balance, balance2, balance3, intrian, humanb, insect, box4, leg, hat = (
    _legacy(n) for n in ("test", "leg2", "box", "box2", "balance", "balance2", "balance3", "intrian",
                         "humanb", "insect", "box4", "leg", "hat"))
quad_balance = _legacy("quad_balance")
quad_balance_chain = _legacy("quad_balance_chain")
