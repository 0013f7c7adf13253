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


class _Arena:
    """Every per-step tensor of an E = 1 core re-homed into ONE device buffer with a pinned host mirror, so that a facade
    step is one host-to-device copy (state + action), one kernel and one device-to-host copy (state + results) instead
    of a dozen small torch operations and as many synchronisations.  Layout: [inputs / state | results], 8-byte slots."""

    def __init__(self, core: BatchedPhysicsEnv, max_act: int):
        f32, n3, M = torch.float32, 3 * core.N, core.M
        fields = [("pos", n3, f32), ("vel", n3, f32), ("old_a", n3, f32), ("mx", M, f32), ("steps", 1, torch.int32),
                  ("action", max_act, f32), ("noise", n3, f32)]
        if core.x64:
            fields += [("mx64", M, torch.float64), ("mx_weak", M, torch.uint8), ("action64", max_act, torch.float64)]
        n_in = len(fields)
        fields += [("obs", core.obs_dim, f32), ("reward", 1, f32), ("done", 1, torch.uint8), ("energy", 1, f32),
                   ("centroid", 3, f32), ("contact_pre", 1, torch.int32), ("contact_post", 1, torch.int32)]
        off, self._slots = 0, {}
        for i, (name, n, dt) in enumerate(fields):
            if i == n_in:
                self.in_bytes = off
            size = n * torch.empty(0, dtype=dt).element_size()
            self._slots[name] = (off, size, dt)
            off += (max(size, 1) + 7) // 8 * 8
        self.dev = torch.zeros(off, dtype=torch.uint8, device=core.device)
        self.host = torch.zeros(off, dtype=torch.uint8).pin_memory()
        self.h = {k: self.host[o:o + sz].view(dt).numpy() for k, (o, sz, dt) in self._slots.items()}    # numpy views of the mirror
        self.stream_device = core.device

    def view(self, name: str, shape) -> torch.Tensor:
        o, sz, dt = self._slots[name]
        return self.dev[o:o + sz].view(dt).view(shape)

    def push(self) -> None:
        self.dev[:self.in_bytes].copy_(self.host[:self.in_bytes], non_blocking=True)

    def pull(self) -> None:
        self.host.copy_(self.dev, non_blocking=True)
        torch.cuda.current_stream(self.stream_device).synchronize()


def _rehome(core: BatchedPhysicsEnv, max_act: int) -> _Arena:
    """Move ``core``'s (SoA, E = 1) tensors into an arena, keeping their current values."""
    ar = _Arena(core, max_act)
    n3, M = 3 * core.N, core.M

    def move(attr, name, shape):
        old = getattr(core, attr)
        new = ar.view(name, shape)
        new.copy_(old.reshape(shape))
        setattr(core, attr, new)

    move("_pos", "pos", (n3, 1)); move("_vel", "vel", (n3, 1)); move("old_a", "old_a", (n3, 1)); move("_mx", "mx", (M, 1))
    move("_steps", "steps", (1,))
    move("obs", "obs", tuple(core.obs.shape)); move("reward", "reward", (1,)); move("_done_u8", "done", (1,))
    core.done = core._done_u8.view(torch.bool)
    move("energy", "energy", (1,)); move("centroid", "centroid", (3, 1))
    move("contact_pre", "contact_pre", (1,)); move("contact_post", "contact_post", (1,))
    if core.x64:
        move("mx64", "mx64", (M, 1)); move("mx_weak", "mx_weak", (M, 1))
    core._bind()
    return ar


class _DeviceBody:
    """One creature mirrored on the device (E = 1)."""

    def __init__(self, creature: Creature, env_kwargs: dict, device):
        self.creature = creature
        self.core = BatchedPhysicsEnv(creature, 1, device, auto_reset=None, keep_old_a=True, track_info=True,
                                      track_contacts=True, track_stats=False, initial_reset=False, state_layout="soa",
                                      **env_kwargs)
        self._max_act = max(self.core.M, 1) + 8               # the reference accepts more actions than muscles
        self.arena = _rehome(self.core, self._max_act)
        self.noise = self.arena.view("noise", (3 * self.core.N, 1))
        self._env_kwargs, self._device = dict(env_kwargs), device
        self.core64, self.arena64 = None, None               # x64 twin (float64 actions), built on first use

    def x64_core(self) -> BatchedPhysicsEnv:
        if self.core64 is None:
            self.core64 = BatchedPhysicsEnv(self.creature, 1, self._device, auto_reset=None, keep_old_a=True, track_info=True,
                                            track_contacts=True, track_stats=False, initial_reset=False, x64=True,
                                            **self._env_kwargs)
            self.arena64 = _rehome(self.core64, self._max_act)
        return self.core64

    def upload(self, steps: int, x64: bool = False, action=None, noise=None):
        """Python objects -> pinned arena -> ONE host-to-device copy.  Returns the device view of the action (or None)."""
        c, cr = (self.x64_core() if x64 else self.core), self.creature
        ar = self.arena64 if x64 else self.arena
        h = ar.h
        n3 = 3 * c.N
        h["pos"][:] = np.concatenate([np.asarray(p.pos, _f32) for p in cr.phys])
        h["vel"][:] = np.concatenate([np.asarray(p.v, _f32) for p in cr.phys])
        h["old_a"][:] = np.concatenate([np.asarray(p.old_a, _f32) for p in cr.phys])
        if c.M:
            h["mx"][:] = [m.x for m in cr.muscles]
            if x64:     # the Muscle objects are the state: their python type is the "weak" bit (np.float64 = strong)
                h["mx64"][:] = [float(m.x) for m in cr.muscles]
                h["mx_weak"][:] = [0 if isinstance(m.x, np.float64) else 1 for m in cr.muscles]
        h["steps"][0] = int(steps)
        if noise is not None:
            h["noise"][:] = noise
        act_view = None
        if action is not None:
            key = "action64" if x64 else "action"
            a = np.asarray(action, dtype=np.float64 if x64 else _f32).reshape(-1)
            if a.size > self._max_act:
                a = a[:self._max_act]                         # only the first min(len, M) drive muscles anyway
            h[key][:a.size] = a
            act_view = ar.view(key, (self._max_act,))[:a.size].view(1, a.size)
        ar.push()
        return act_view

    def download(self, refresh_contact: bool, x64: bool = False, muscles: bool = True) -> dict:
        """ONE device-to-host copy of state + results, then the Point / Muscle objects are refreshed from it.
        Returns the host views (obs, reward, done, energy, centroid) of this step."""
        c, cr = (self.core64 if x64 else self.core), self.creature
        ar = self.arena64 if x64 else self.arena
        ar.pull()
        h = ar.h
        pos, vel, old_a = h["pos"], h["vel"], h["old_a"]
        cpre = int(h["contact_pre"][0]) if refresh_contact else 0
        for n, p in enumerate(cr.phys):
            p.pos[:] = pos[3 * n:3 * n + 3]
            p.v[:] = vel[3 * n:3 * n + 3]
            p.old_a = old_a[3 * n:3 * n + 3].copy()
            p.zero()
            if refresh_contact:             # colour / radius side effects (gym/optimized_env.py:155-156,174-175)
                hit = (cpre >> n) & 1
                p.color, p.r = ("red", 3) if hit else ("black", 1)
        if not muscles:                     # reset() does not touch Muscle.x (value or type)
            return h
        if x64 and c.M:
            x64v, weak = h["mx64"], h["mx_weak"]
            for i, m in enumerate(cr.muscles):
                v = float(x64v[i])
                # strong = np.float64 (after `x += np.float64`); weak = the limit / constructor object: np.float32
                # when representable, else the python float the user passed
                m.x = np.float64(v) if not weak[i] else (_f32(v) if float(_f32(v)) == v else v)
            return h
        mx = h["mx"]
        for i, m in enumerate(cr.muscles):
            m.x = _f32(mx[i])
        return h


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
        key = (self.in3d, self.g, self.dampk, self.ground, self.ground_k, self.ground_damp, self.friction, self.sigma,
               self.time_step if time_step is None else time_step, self.max_steps)
        cache = self.__dict__.setdefault("_params_cache", {})
        if cache.get("key") != key:          # rebuilt only when the caller changed an attribute between steps
            cache["key"] = key
            cache["params"] = make_params(in3d=key[0], g=key[1], dampk=key[2], ground_high=key[3], ground_k=key[4],
                                          ground_damp=key[5], friction=key[6], rand_sigma=key[7], time_step=key[8],
                                          max_steps=key[9], k_sub=1, auto_reset=0)
        self._body.core.params = cache["params"]

    def _obs_array(self, x64: bool = False) -> np.ndarray:
        ar = self._body.arena64 if x64 else self._body.arena
        core = self._body.core64 if x64 else self._body.core
        obs = ar.h["obs"].astype(np.float64)                 # the host mirror the last download() filled
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
        body.upload(0, noise=nz.reshape(-1))
        core.reset(noise=body.noise, mode="jitter")
        body.download(refresh_contact=False, muscles=False)
        self.steps = 0
        return self._obs_array()

    def step(self, action) -> Tuple[np.ndarray, float, bool, Dict[str, Any]]:
        """One environment step (gym/optimized_env.py:70-92).  ``action`` drives the first
        min(len(action), M) muscles.  NumPy's promotion rules decide the arithmetic exactly as in the
        reference: python floats and float32 arrays keep ``Muscle.x`` in float32; a float64 ndarray (what
        ``np.random.uniform`` returns, gym/performance_demo.py:241-262) turns it into an np.float64 and the
        muscle's spring term into double (x64 mode, ``wg_step_x64``) -- both are bit-identical to the reference.
        One host-to-device copy, one kernel, one device-to-host copy per call (``_Arena``)."""
        body = self._body
        if isinstance(action, np.ndarray) and action.dtype != object:
            x64 = action.dtype == np.float64
        else:
            x64 = any(isinstance(v, np.float64) for v in np.asarray(action, dtype=object).reshape(-1))
        x64 = x64 or any(isinstance(m.x, np.float64) for m in self.creature.muscles)
        core = body.x64_core() if x64 else body.core
        self._refresh_params()
        if x64:
            core.params = body.core.params
        act = body.upload(self.steps, x64, action=action)
        core.step(act)
        h = body.download(refresh_contact=True, x64=x64)
        self.steps += 1
        reward = _f32(h["reward"][0])
        done = bool(h["done"][0])
        if self.renderer is not None and not self.renderer.is_running():
            done = True
        info = {"steps": self.steps,
                "centroid_position": h["centroid"].tolist(),
                "total_energy": _f32(h["energy"][0])}
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
