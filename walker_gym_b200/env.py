"""Drop-in single-env surface of ``gym/optimized_env.py``: ``PhysicsEnv``,
``make_env`` and the compat ``Environment``.

Same constructor arguments, attributes, return values and errors as the
reference; the work of ``step``/``reset`` is one call into the CUDA library
with ``n_env = 1``.  The ``Point``/``Muscle`` objects the caller passed in stay
the user-visible state (the reference mutates them in place, and callers read
``p.pos`` / ``m.x`` back): they are uploaded before and refreshed after every
call.  For throughput use ``BatchedPhysicsEnv``; this class exists so code
written against the reference runs unchanged.
"""
from __future__ import annotations

import warnings
import weakref
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from .batched import BatchedPhysicsEnv, creature_from_id, make_params
from .engine import Point
from .walker import Creature

_f32 = np.float32


class _DeviceBody:
    """One creature mirrored on the device (E = 1)."""

    def __init__(self, creature: Creature, env_kwargs: dict, device):
        self.creature = creature
        self.core = BatchedPhysicsEnv(creature, 1, device, auto_reset=None, keep_old_a=True, track_info=True,
                                      track_contacts=True, track_stats=False, initial_reset=False, state_layout="soa",
                                      **env_kwargs)
        c = self.core
        self.noise = torch.zeros(3 * c.N, 1, dtype=torch.float32, device=c.device)
        self._env_kwargs, self._device = dict(env_kwargs), device
        self.core64 = None                  # x64 twin (float64 actions), built on first use

    def x64_core(self) -> BatchedPhysicsEnv:
        if self.core64 is None:
            self.core64 = BatchedPhysicsEnv(self.creature, 1, self._device, auto_reset=None, keep_old_a=True, track_info=True,
                                            track_contacts=True, track_stats=False, initial_reset=False, x64=True,
                                            **self._env_kwargs)
        return self.core64

    def upload(self, steps: int, x64: bool = False) -> None:
        c, cr = (self.x64_core() if x64 else self.core), self.creature
        host = np.concatenate([np.asarray(p.pos, _f32) for p in cr.phys] + [np.asarray(p.v, _f32) for p in cr.phys]
                              + [np.asarray(p.old_a, _f32) for p in cr.phys] + [_f32([m.x for m in cr.muscles])])
        dev = torch.from_numpy(host).to(c.device)
        n3 = 3 * c.N
        c.pos[:, 0], c.vel[:, 0], c.old_a[:, 0] = dev[:n3], dev[n3:2 * n3], dev[2 * n3:3 * n3]
        if c.M:
            c.mx[:, 0] = dev[3 * n3:]
            if x64:     # the Muscle objects are the state: their python type is the "weak" bit (np.float64 = strong)
                c.mx64[:, 0] = torch.tensor([float(m.x) for m in cr.muscles], dtype=torch.float64)
                c.mx_weak[:, 0] = torch.tensor([0 if isinstance(m.x, np.float64) else 1 for m in cr.muscles], dtype=torch.uint8)
        c.steps.fill_(int(steps))

    def download(self, refresh_contact: bool, x64: bool = False, muscles: bool = True) -> None:
        c, cr = (self.core64 if x64 else self.core), self.creature
        n3 = 3 * c.N
        host = torch.cat([c.pos[:, 0], c.vel[:, 0], c.old_a[:, 0], c.mx[:, 0]]).cpu().numpy()
        cpre = int(c.contact_pre.item()) if refresh_contact else 0
        for n, p in enumerate(cr.phys):
            p.pos[:] = host[3 * n:3 * n + 3]
            p.v[:] = host[n3 + 3 * n:n3 + 3 * n + 3]
            p.old_a = host[2 * n3 + 3 * n:2 * n3 + 3 * n + 3].copy()
            p.zero()
            if refresh_contact:             # colour / radius side effects (gym/optimized_env.py:155-156,174-175)
                hit = (cpre >> n) & 1
                p.color, p.r = ("red", 3) if hit else ("black", 1)
        if not muscles:                     # reset() does not touch Muscle.x (value or type)
            return
        if x64 and c.M:
            x64v, weak = c.mx64[:, 0].cpu().numpy(), c.mx_weak[:, 0].cpu().numpy()
            for i, m in enumerate(cr.muscles):
                v = float(x64v[i])
                # strong = np.float64 (after `x += np.float64`); weak = the limit / constructor object: np.float32
                # when representable, else the python float the user passed
                m.x = np.float64(v) if not weak[i] else (_f32(v) if float(_f32(v)) == v else v)
            return
        for i, m in enumerate(cr.muscles):
            m.x = _f32(host[3 * n3 + i])


class PhysicsEnv:
    """Gym-style environment around one creature (gym/optimized_env.py:8-269)."""

    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 60}

    def __init__(self, creature: Creature, in3d: bool = False, g: float = 100, dampk: float = 0,
                 ground_high: float = 0, ground_k: float = 1000, ground_damp: float = 100,
                 friction: float = 100, rand_sigma: float = 0.1, device: Union[str, torch.device] = "cuda"):
        self.creature = creature
        self.in3d = in3d
        self.g, self.dampk, self.ground = g, dampk, ground_high
        self.ground_k, self.ground_damp, self.friction = ground_k, ground_damp, friction
        self.sigma = rand_sigma
        self.time_step = 0.01
        self.steps = 0
        self.max_steps = 1000
        self.renderer = None
        self.render_mode = None
        self._body = _DeviceBody(creature, dict(in3d=in3d), device)
        self.reset()

    # the reference reads its attributes on every call, so they can be changed between steps
    def _refresh_params(self, time_step=None):
        p = make_params(in3d=self.in3d, g=self.g, dampk=self.dampk, ground_high=self.ground, ground_k=self.ground_k,
                        ground_damp=self.ground_damp, friction=self.friction, rand_sigma=self.sigma,
                        time_step=self.time_step if time_step is None else time_step, max_steps=self.max_steps,
                        k_sub=1, auto_reset=0)
        self._body.core.params = p

    def _obs_array(self, x64: bool = False) -> np.ndarray:
        core = self._body.core64 if x64 else self._body.core
        obs = core.obs[0].cpu().numpy().astype(np.float64)
        if core.M:                          # the reference's observation carries Muscle.x at its own precision
            obs[-core.M:] = [float(m.x) for m in self.creature.muscles]
        return obs

    def reset(self) -> np.ndarray:
        """Jitter-only reset, exactly the reference's (gym/optimized_env.py:53-68):
        accelerations cleared, N(0, sigma) added to each velocity component, steps = 0.
        The draws come from ``np.random.normal`` in the reference's call order, so
        ``seed()`` reproduces the reference's jitter."""
        body, core = self._body, self._body.core
        d = 3 if self.in3d else 2
        nz = np.zeros((core.N, 3), _f32)
        for n in range(core.N):
            for c in range(d):
                nz[n, c] = np.random.normal(0, self.sigma)
        self._refresh_params()
        body.upload(0)
        body.noise.copy_(torch.from_numpy(nz.reshape(-1, 1)))
        core.reset(noise=body.noise, mode="jitter")
        body.download(refresh_contact=False, muscles=False)
        self.steps = 0
        return self._obs_array()

    def step(self, action) -> Tuple[np.ndarray, float, bool, Dict[str, Any]]:
        """One environment step (gym/optimized_env.py:70-92).  ``action`` drives the first
        min(len(action), M) muscles.  NumPy's promotion rules decide the arithmetic exactly as in the
        reference: python floats and float32 arrays keep ``Muscle.x`` in float32; a float64 ndarray (what
        ``np.random.uniform`` returns, gym/performance_demo.py:241-262) turns it into an np.float64 and the
        muscle's spring term into double (x64 mode, ``wg_step_x64``) -- both are bit-identical to the reference."""
        body = self._body
        x64 = ((isinstance(action, np.ndarray) and action.dtype == np.float64)
               or any(isinstance(v, np.float64) for v in np.asarray(action, dtype=object).reshape(-1))
               or any(isinstance(m.x, np.float64) for m in self.creature.muscles))
        core = body.x64_core() if x64 else body.core
        self._refresh_params()
        if x64:
            core.params = body.core.params
        body.upload(self.steps, x64)
        act = np.asarray(action, dtype=np.float64 if x64 else _f32).reshape(1, -1)
        core.step(torch.from_numpy(act).to(core.device))
        body.download(refresh_contact=True, x64=x64)
        self.steps += 1
        reward = _f32(core.reward.item())
        done = bool(core.done.item())
        if self.renderer is not None and not self.renderer.is_running():
            done = True
        info = {"steps": self.steps,
                "centroid_position": core.centroid[:, 0].cpu().numpy().tolist(),
                "total_energy": _f32(core.energy.item())}
        return self._obs_array(x64), reward, done, info

    def render(self, mode: str = "human") -> Optional[np.ndarray]:
        """Rendering (pygame) is outside the accelerated path; this is a no-op."""
        self.render_mode = mode
        warnings.warn("walker_gym_b200 does not render; render() is a no-op", RuntimeWarning, stacklevel=2)
        return None

    def close(self) -> None:
        self.renderer = None

    def seed(self, seed: Optional[int] = None) -> List[int]:
        np.random.seed(seed)
        return [seed] if seed is not None else []

    def get_action_space(self) -> Dict[str, Any]:
        return {"shape": (len(self.creature.muscles),), "type": "continuous", "low": -1.0, "high": 1.0}

    def get_observation_space(self) -> Dict[str, Any]:
        return {"shape": (self._body.core.obs_dim,), "type": "continuous", "low": -np.inf, "high": np.inf}

    # same private helpers as the reference, for code that reaches into them
    def _get_observation(self) -> np.ndarray:
        return np.array(self.creature.getstat(self.in3d))


def make_env(env_id: str, **kwargs) -> PhysicsEnv:
    """``make_env('Balance-v0' | 'Box-v0', **kw)`` (gym/optimized_env.py:273-294);
    unknown ids raise ``ValueError`` like the reference."""
    key = env_id.lower()
    if key not in ("balance-v0", "box-v0"):
        raise ValueError(f"Unknown environment ID: {key}")
    return PhysicsEnv(creature_from_id(key), **kwargs)


_live_environments: "weakref.WeakSet[Environment]" = weakref.WeakSet()


class Environment(PhysicsEnv):
    """Legacy-signature environment over a list of creatures
    (gym/optimized_env.py:298-334; legacy gym/env.py:9-50).  ``step(t)`` applies
    every creature's springs and the environment forces, then integrates with the
    caller's ``t``.  The legacy two-call form ``env.run(); Point.run1(t)`` is
    supported: ``run()`` stages the force pass and ``Point.run1`` integrates it."""

    def __init__(self, creaturelist, in3d=False, g=100, dampk=0, groundhigh=0, groundk=1000, grounddamp=100,
                 friction=100, randsigma=0.1, device: Union[str, torch.device] = "cuda"):
        creature = creaturelist[0] if creaturelist else None
        if creature is None:
            raise ValueError("Environment needs at least one creature")
        super().__init__(creature, in3d, g, dampk, groundhigh, groundk, grounddamp, friction, randsigma, device=device)
        self.creatures = creaturelist
        self._bodies = [self._body] + [_DeviceBody(c, dict(in3d=in3d), device) for c in creaturelist[1:]]
        self._staged = False
        _live_environments.add(self)

    def run(self) -> None:
        self._staged = True

    def _integrate(self, t) -> None:
        for body in self._bodies:
            self._body = body
            self._refresh_params(time_step=t)
            body.upload(self.steps)
            body.core.step(None)
            body.download(refresh_contact=True)
        self._body = self._bodies[0]
        self._staged = False

    def step(self, t):  # noqa: D401 - legacy signature: step(dt), returns None
        self.run()
        self._integrate(t)


def _run1(cls, t: float) -> None:
    """``Point.run1(t)`` (gym/optimized_engine.py:258-272): integrate what ``Environment.run`` staged."""
    for env in list(_live_environments):
        if env._staged:
            env._integrate(t)


Point.run1 = classmethod(_run1)
