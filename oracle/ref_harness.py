"""TEST INFRASTRUCTURE ONLY -- executes the *unmodified* reference files.

This module imports walker-gym's own L1 "optimized flat" modules from a
read-only checkout (default ``/root/reference/gym``) so that golden vectors
can be generated from, and the C restatement (``walker_oracle.c``) validated
against, the reference itself.  It only works where the checkout exists (the
authoring container); it never travels to the GPU box and nothing in the
product path, ``bench.py`` or the ``-m gpu`` tests may import it.

The reference does not run as shipped.  The harness
applies exactly three things, none of which touches the arithmetic:

1. ``pygame`` and ``turtle`` are replaced by inert stubs in ``sys.modules``
   (imported at ``gym/optimized_env.py:4,339`` and
   ``gym/optimized_walker.py:2`` only for rendering).
2. ``gym/optimized_walker.py`` is loaded *by path* under the module name
   ``optimized_walker`` because the package directory of the same name
   shadows it (``gym/optimized_env.py:5,281``).
3. ``Point.forced`` (``gym/optimized_engine.py:104-106``) receives Python
   lists from ``PhysicsEnv._run_physics`` (``gym/optimized_env.py:148-172``)
   and would raise ``TypeError``; the shim converts with ``np.asarray`` and
   zero-pads a 2-vector to the 3-vector ``a`` (2-D mode), then calls the
   original.  The resulting dtype promotion (int64/float64 list -> float64
   divide -> rounded into the float32 accumulator) is the reference's own.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np

DEFAULT_REF = os.environ.get("WALKER_GYM_REFERENCE", "/root/reference")

_loaded = {}


# the byte-compiled copy made by oracle/make_ref.py (bytecode of the unmodified modules; travels to the GPU box)
COMPILED_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _module_file(gym_dir: str, stem: str):
    for ext in (".py", ".refbc"):
        p = os.path.join(gym_dir, stem + ext)
        if os.path.isfile(p):
            return p
    return None


def available(ref_root: str = DEFAULT_REF) -> bool:
    return _module_file(os.path.join(ref_root, "gym"), "optimized_env") is not None


def load(ref_root: str = DEFAULT_REF):
    """Return (engine, walker, env) reference modules, harnessed as above."""
    if ref_root in _loaded:
        return _loaded[ref_root]
    gym_dir = os.path.join(ref_root, "gym")
    if not available(ref_root):
        raise FileNotFoundError(f"reference checkout not found under {ref_root}")
    for name in ("pygame", "turtle"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in
                  ("optimized_engine", "optimized_renderer", "optimized_walker", "optimized_env")}
    sys.path.insert(0, gym_dir)
    try:
        def by_path(modname, fname):
            path = _module_file(gym_dir, fname[:-3])
            if path.endswith(".refbc"):        # bytecode made by oracle/make_ref.py: a .pyc under a neutral extension
                from importlib.machinery import SourcelessFileLoader
                loader = SourcelessFileLoader(modname, path)
                spec = importlib.util.spec_from_loader(modname, loader, origin=path)
            else:
                spec = importlib.util.spec_from_file_location(modname, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[modname] = mod
            spec.loader.exec_module(mod)
            return mod

        engine = by_path("optimized_engine", "optimized_engine.py")
        by_path("optimized_renderer", "optimized_renderer.py")
        walker = by_path("optimized_walker", "optimized_walker.py")
        env = by_path("optimized_env", "optimized_env.py")
    finally:
        sys.path[:] = saved_path

    orig_forced = engine.Point.forced

    def forced(self, f):
        f = np.asarray(f)
        if f.shape[0] < self.a.shape[0]:
            f = np.concatenate([f, np.zeros(self.a.shape[0] - f.shape[0], dtype=f.dtype)])
        return orig_forced(self, f)

    engine.Point.forced = forced
    _loaded[ref_root] = (engine, walker, env)
    return _loaded[ref_root]


def build_creature(spec, ref_root: str = DEFAULT_REF):
    """Build a reference ``Creature`` from a body spec (dict, see bodies.py).

    spec: {"points": [(m, pos, fixed)], "muscles": [(i, j, kwargs)],
           "skeletons": [(i, j, kwargs)]}
    """
    engine, walker, _ = load(ref_root)
    engine.Point.clear()
    pts = []
    for m, pos, fixed in spec["points"]:
        if fixed:
            pts.append(engine.DingPoint(m, list(pos)))
        else:
            pts.append(engine.Point(m, list(pos), [0, 0, 0]))
    mus = [walker.Muscle(pts[i], pts[j], **kw) for i, j, kw in spec["muscles"]]
    sks = [walker.Skeleton(pts[i], pts[j], **kw) for i, j, kw in spec["skeletons"]]
    return walker.Creature(pts, mus, sks)


def snapshot(env):
    """Copy the full observable state of a reference env into numpy arrays."""
    c = env.creature
    return dict(
        pos=np.array([p.pos for p in c.phys], dtype=np.float32),
        vel=np.array([p.v for p in c.phys], dtype=np.float32),
        old_a=np.array([p.old_a for p in c.phys], dtype=np.float32),
        x=np.array([float(m.x) for m in c.muscles], dtype=np.float64),
        contact_pre=np.array([p.r == 3 for p in c.phys], dtype=np.bool_),
    )


def rollout(creature_or_id, actions, *, env_kwargs=None, seed=0, noise=None,
            max_steps=None, k_sub=1, reset_on_done=False, init_state=None, integrator="run1",
            ref_root: str = DEFAULT_REF):
    """Run the reference ``PhysicsEnv`` for ``len(actions)`` steps.

    ``actions``: float32 [T, M] (float32 keeps ``Muscle.x`` in float32,
    SURVEY.md section 7.7).  ``noise``: optional flat array that replaces the
    ``np.random.normal`` draws of ``reset()`` in call order (per point: x, y,
    then z if in3d) so a kernel can be fed the identical jitter.
    ``k_sub``: physics substeps per env step: ``act`` once,
    then ``k_sub`` x ``_run_physics``; k_sub == 1 calls ``env.step`` itself.
    ``reset_on_done``: call ``env.reset()`` after a done step (the batched
    library's "jitter" auto-reset); the recorded state/obs of that step are
    then the post-reset ones, reward/done the pre-reset ones.
    ``init_state``: optional dict(pos=[N,3], vel=[N,3]) written into the
    creature's points after construction.
    Returns a dict of per-step arrays with a leading T+1 (state) or T axis.
    """
    engine, walker, envmod = load(ref_root)
    env_kwargs = dict(env_kwargs or {})
    engine.Point.clear()
    np.random.seed(seed)
    draws = []
    # integrator="run2": PhysicsEnv._run_physics calls Point.run1 (gym/optimized_env.py:178); route that
    # call to the reference's own Point.run2 (gym/optimized_engine.py:274-288) for the duration of the rollout
    run1_saved = engine.Point.__dict__["run1"]
    if integrator == "run2":
        engine.Point.run1 = engine.Point.__dict__["run2"]

    real_normal = np.random.normal

    def fake_normal(loc=0.0, scale=1.0, size=None):
        if noise is not None:
            v = float(np.asarray(noise).reshape(-1)[len(draws)])
        else:
            v = real_normal(loc, scale, size)
        draws.append(v)
        return v

    with mock.patch.object(np.random, "normal", fake_normal):
        if isinstance(creature_or_id, str):
            env = envmod.make_env(creature_or_id, **env_kwargs)
        elif isinstance(creature_or_id, dict):
            env = envmod.PhysicsEnv(build_creature(creature_or_id, ref_root), **env_kwargs)
        else:
            env = envmod.PhysicsEnv(creature_or_id, **env_kwargs)
        if init_state is not None:
            for n, p in enumerate(env.creature.phys):
                p.pos[:] = np.asarray(init_state["pos"][n], dtype=np.float32)
                p.v[:] = np.asarray(init_state["vel"][n], dtype=np.float32)
        obs0 = np.asarray(env._get_observation(), dtype=np.float64)
        if max_steps is not None:
            env.max_steps = max_steps
        T = len(actions)
        snaps = [snapshot(env)]
        obs = [obs0]
        rew, done, energy, centroid, steps = [], [], [], [], []
        for t in range(T):
            if k_sub == 1:
                o, r, d, info = env.step(actions[t])
            else:
                env.creature.act(actions[t])
                for _ in range(k_sub):
                    env._run_physics()
                env.steps += 1
                o, r, d, info = (env._get_observation(), env._get_reward(), env._is_done(),
                                 env._get_info())
            cpre = np.array([p.r == 3 for p in env.creature.phys], dtype=np.bool_)
            if d and reset_on_done:
                o = env.reset()
            sn = snapshot(env)
            sn["contact_pre"] = cpre
            snaps.append(sn)
            obs.append(np.asarray(o, dtype=np.float64))
            rew.append(r)
            done.append(bool(d))
            energy.append(info["total_energy"])
            centroid.append(info["centroid_position"])
            steps.append(env.steps)
    engine.Point.run1 = run1_saved
    out = {k: np.stack([s[k] for s in snaps]) for k in snaps[0]}
    out["contact_pre"] = out["contact_pre"][1:]
    out.update(
        obs=np.stack(obs),
        reward=np.asarray(rew, dtype=np.float64),
        done=np.asarray(done, dtype=np.bool_),
        energy=np.asarray(energy, dtype=np.float64),
        centroid=np.asarray(centroid, dtype=np.float64).reshape(T, 3),
        steps=np.asarray(steps, dtype=np.int64),
        reset_noise=np.asarray(draws, dtype=np.float64),
        masses=np.array([float(p.m) for p in env.creature.phys], dtype=np.float64),
    )
    engine.Point.clear()
    return out


def rollout_compat(spec_or_id, actions, t_step, *, env_kwargs=None, noise=None, ref_root: str = DEFAULT_REF):
    """Legacy-signature ``Environment`` (gym/optimized_env.py:298-334): per step the caller does
    ``creature.act(a)`` then ``env.step(t)`` (= run() + Point.run1(t)); returns the point states."""
    engine, walker, envmod = load(ref_root)
    engine.Point.clear()
    draws = []

    def fake_normal(loc=0.0, scale=1.0, size=None):
        v = float(np.asarray(noise).reshape(-1)[len(draws)])
        draws.append(v)
        return v

    with mock.patch.object(np.random, "normal", fake_normal):
        if isinstance(spec_or_id, dict):
            creature = build_creature(spec_or_id, ref_root)
        else:
            creature = {"balance-v0": walker.create_balance_creature, "box-v0": walker.create_box_creature}[spec_or_id.lower()]()
        env = envmod.Environment([creature], **dict(env_kwargs or {}))
        snaps = [snapshot(env)]
        for a in actions:
            creature.act(a)
            env.step(t_step)
            snaps.append(snapshot(env))
    out = {k: np.stack([s[k] for s in snaps]) for k in snaps[0]}
    out["contact_pre"] = out["contact_pre"][1:]
    out["reset_noise"] = np.asarray(draws, dtype=np.float64)
    engine.Point.clear()
    return out
