"""Creature -> ``wg_topology``: evaluate the morphology the way the reference's
constructors and NumPy would, once, on the host."""
from __future__ import annotations

import numpy as np

from ._lib import MAX_MASS, MAX_SPRING, WgTopology, WgX64
from .walker import Creature


def topology_from_creature(creature: Creature) -> WgTopology:
    """Springs are emitted muscles-first then skeletons -- the order ``Creature.run``
    applies them (gym/optimized_walker.py:117-127), which fixes the float32
    accumulation order on the device."""
    phys = list(creature.phys)
    index = {id(p): n for n, p in enumerate(phys)}
    springs = list(creature.muscles) + list(creature.skeletons)
    if not 1 <= len(phys) <= MAX_MASS:
        raise ValueError(f"a creature needs 1..{MAX_MASS} points, got {len(phys)}")
    if len(springs) > MAX_SPRING:
        raise ValueError(f"at most {MAX_SPRING} springs are supported, got {len(springs)}")
    t = WgTopology()
    t.n_mass, t.n_muscle, t.n_spring = len(phys), len(creature.muscles), len(springs)
    for n, p in enumerate(phys):
        t.mass[n] = float(p.m)
        t.fixed[n] = 1 if getattr(p, "fixed", False) else 0
        tp = getattr(p, "original_pos", None) if getattr(p, "fixed", False) else None
        src = p.pos if tp is None else tp
        for c in range(3):
            t.tmpl_pos[n * 3 + c] = np.float32(src[c])
    for s, sp in enumerate(springs):
        try:
            t.si[s], t.sj[s] = index[id(sp.p1)], index[id(sp.p2)]
        except KeyError:
            raise ValueError("a spring references a point that is not in creature.phys") from None
        if t.si[s] == t.sj[s]:
            raise ValueError("a spring must join two different points")
        t.sk[s] = np.float32(sp.k)
        t.sdamp[s] = np.float32(sp.dampk)
        t.sstring[s] = 1 if getattr(sp, "string", False) else 0
        if s < t.n_muscle:
            t.srest[s] = np.float32(sp.originx)
            t.mlo[s] = np.float32(sp.originx * sp.minl)     # Muscle.regulation operands, python semantics
            t.mhi[s] = np.float32(sp.originx * sp.maxl)
        else:
            t.srest[s] = np.float32(sp.x)
    return t


def x64_from_creature(creature: Creature) -> WgX64:
    """The double-typed objects the reference's ``Muscle`` keeps (x64 mode, ``wg_step_x64``): ``float(k)``; ``originx``
    (the np.float32 distance, or the python float the user passed); ``originx * minl`` / ``originx * maxl`` exactly as
    ``regulation`` forms them (np.float32 * python float -> np.float32; python * python -> double)."""
    x = WgX64()
    for s, sp in enumerate(list(creature.muscles) + list(creature.skeletons)):
        x.sk_d[s] = float(sp.k)
        if s < len(creature.muscles):
            x.x0_d[s] = float(sp.originx)
            x.mlo_d[s] = float(sp.originx * sp.minl)
            x.mhi_d[s] = float(sp.originx * sp.maxl)
        else:
            x.x0_d[s] = float(sp.x)
    return x


def topology_from_spec(spec) -> WgTopology:
    """Same, from a plain dict spec {"points": [(m, pos, fixed)], "muscles": [(i, j, kw)], "skeletons": [...]}."""
    from .engine import DingPoint, Point
    from .walker import Muscle, Skeleton
    saved = list(Point.points)
    try:
        pts = [DingPoint(m, list(pos)) if fixed else Point(m, list(pos), [0, 0, 0]) for m, pos, fixed in spec["points"]]
        mus = [Muscle(pts[i], pts[j], **kw) for i, j, kw in spec["muscles"]]
        sks = [Skeleton(pts[i], pts[j], **kw) for i, j, kw in spec["skeletons"]]
        return topology_from_creature(Creature(pts, mus, sks))
    finally:
        Point.points = saved
