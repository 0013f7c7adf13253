"""In-kernel action sources for ``BatchedPhysicsEnv.step_many`` (``wg_action_gen``): open-loop controllers whose
action is a function of the env's own step counter, evaluated inside the T-steps-per-launch kernel so that a launch
reads no action memory at all.

* ``ScriptedActions``: the phase-table gait sketched at the end of gym/walker.py (:356-366) --
  ``tt = (t // 50) % 3; c.act([... row tt ...])``.
* ``CPGActions``: a sinusoidal central pattern generator after the package lineage's ``Muscle.act``
  (gym/optimized_walker/walker.py:56-90: ``t += dt; sin(2 pi freq t + phase)``), used as an action source.
"""
from __future__ import annotations

import math
from typing import Sequence

from ._lib import GEN_MAX_MUSCLE, GEN_MAX_ROWS, WgActionGen


class ActionSource:
    """Base: anything with ``.struct(n_muscle, time_step) -> WgActionGen``."""

    def struct(self, n_muscle: int, time_step: float) -> WgActionGen:          # pragma: no cover
        raise NotImplementedError


class ScriptedActions(ActionSource):
    """``action[m] = table[(steps // hold) % len(table)][m]`` with ``steps`` the env's step counter before the step
    (an auto-reset restarts the gait).  ``table``: up to 32 rows of one value per muscle."""

    def __init__(self, table: Sequence[Sequence[float]], hold: int = 50):
        self.table = [[float(v) for v in row] for row in table]
        self.hold = int(hold)
        if not 1 <= len(self.table) <= GEN_MAX_ROWS:
            raise ValueError(f"a scripted table has 1..{GEN_MAX_ROWS} rows")
        if self.hold < 1:
            raise ValueError("hold must be >= 1")
        if len({len(r) for r in self.table}) != 1:
            raise ValueError("every row needs one value per muscle")

    def struct(self, n_muscle: int, time_step: float) -> WgActionGen:
        if len(self.table[0]) != n_muscle:
            raise ValueError(f"table rows have {len(self.table[0])} values, the body has {n_muscle} muscles")
        if n_muscle > GEN_MAX_MUSCLE:
            raise ValueError(f"in-kernel action sources drive at most {GEN_MAX_MUSCLE} muscles")
        g = WgActionGen()
        g.mode, g.n_rows, g.hold = 1, len(self.table), self.hold
        for r, row in enumerate(self.table):
            for m, v in enumerate(row):
                g.table[r * GEN_MAX_MUSCLE + m] = v
        return g

    def describe(self) -> dict:
        return {"mode": 1, "table": self.table, "hold": self.hold}


class CPGActions(ActionSource):
    """``action[m] = amp[m] * sin(2 pi (freq[m] * t + phase[m] / (2 pi)))`` at ``t = (steps + 1) * time_step``.
    The phase is kept as a 24-bit fraction of a turn (``dphase = round(freq * time_step * 2^24)`` per step), so the
    generator is exactly periodic and bit-reproducible on the CPU oracle."""

    def __init__(self, amp: Sequence[float], freq: Sequence[float], phase: Sequence[float] = None):
        self.amp = [float(a) for a in amp]
        self.freq = [float(f) for f in freq]
        self.phase = [0.0] * len(self.amp) if phase is None else [float(p) for p in phase]
        if not (len(self.amp) == len(self.freq) == len(self.phase)):
            raise ValueError("amp, freq and phase need one value per muscle")

    def turns(self, time_step: float):
        p0 = [int(round((p / (2 * math.pi)) * (1 << 24))) & 0xFFFFFF for p in self.phase]
        dp = [int(round(f * time_step * (1 << 24))) & 0xFFFFFF for f in self.freq]
        return p0, dp

    def struct(self, n_muscle: int, time_step: float) -> WgActionGen:
        if len(self.amp) != n_muscle:
            raise ValueError(f"the generator has {len(self.amp)} channels, the body has {n_muscle} muscles")
        if n_muscle > GEN_MAX_MUSCLE:
            raise ValueError(f"in-kernel action sources drive at most {GEN_MAX_MUSCLE} muscles")
        g = WgActionGen()
        g.mode, g.n_rows, g.hold = 2, 1, 1
        p0, dp = self.turns(time_step)
        for m in range(n_muscle):
            g.amp[m], g.phase0[m], g.dphase[m] = self.amp[m], p0[m], dp[m]
        return g

    def describe(self, time_step: float) -> dict:
        p0, dp = self.turns(time_step)
        return {"mode": 2, "amp": self.amp, "phase0": p0, "dphase": dp}
