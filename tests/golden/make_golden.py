"""Generate the golden fixtures in this directory from the reference itself.

Run in the authoring container (needs the read-only reference checkout at
/root/reference):

    python tests/golden/make_golden.py

Every ``*.npz`` here is the output of walker-gym's own ``PhysicsEnv``
(``gym/optimized_env.py``), executed through ``oracle/ref_harness.py`` on
seeded float32 actions.  They pin the C oracle (and through it the CUDA
library) to the reference bit for bit.  Each file carries its inputs
(actions, reset noise, body spec as JSON, env kwargs as JSON) so the tests
need nothing else.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_harness as rh  # noqa: E402
import walker_oracle as wo  # noqa: E402  (only for the BALANCE / BOX specs)

# gym/walker.py:255-293 `insect` as data (legacy constructor order is Phy(m, v, p)).
INSECT = {
    "points": [(1, p, False) for p in [
        (-75, 100, 0), (-25, 100, 0), (25, 100, 0), (75, 100, 0),
        (-100, 50, 0), (-50, 50, 0), (0, 50, 0), (50, 50, 0), (100, 50, 0),
        (-75, 0, 0), (-25, 0, 0), (25, 0, 0), (75, 0, 0)]],
    "muscles": [(9, 4, {}), (9, 5, {}), (10, 5, {}), (10, 6, {}), (11, 6, {}), (11, 7, {}), (12, 7, {}), (12, 8, {})],
    "skeletons": [(0, 1, {}), (0, 4, {}), (0, 5, {}), (1, 2, {}), (1, 5, {}), (1, 6, {}), (2, 3, {}), (2, 6, {}),
                  (2, 7, {}), (3, 7, {}), (3, 8, {}), (4, 5, {}), (5, 6, {}), (6, 7, {}), (7, 8, {})],
}
# An irregular body: float masses, a DingPoint, per-spring k/dampk, an explicit
# python-float rest length, custom muscle limits, non-zero z.
CUSTOM = {
    "points": [(2.5, (-30, 40, 5), False), (0.1, (35, 60, -3), False), (7, (0, 5, 0), False),
               (1, (10, 90, 2), True), (3, (-60, 20, 0), False)],
    "muscles": [(0, 2, {"k": 800, "dampk": 15, "minl": 0.3, "maxl": 1.2}),
                (1, 2, {"x": 70.0, "k": 1200, "dampk": 25}),
                (4, 0, {"k": -300})],
    "skeletons": [(0, 1, {"k": 500}), (1, 3, {"k": 2000, "dampk": 5}), (4, 2, {"x": 61.5})],
}
# Two free unit masses joined by a pure damper (k=0): settles on the ground.
RESTING = {
    "points": [(1, (-10, 0.5, 0), False), (1, (10, 0.25, 0), False)],
    "muscles": [(0, 1, {"k": 0})],
    "skeletons": [],
}


def n_muscles(body):
    return len(body["muscles"]) if isinstance(body, dict) else {"balance-v0": 2, "box-v0": 4}[body.lower()]


def save(name, out, body, env_kwargs, actions, extra=None):
    spec = body if isinstance(body, dict) else {"balance-v0": wo.BALANCE, "box-v0": wo.BOX}[body.lower()]
    arrs = dict(
        pos=out["pos"].astype(np.float32), vel=out["vel"].astype(np.float32),
        old_a=out["old_a"].astype(np.float32), x=out["x"].astype(np.float32),
        obs=out["obs"].astype(np.float32), reward=out["reward"].astype(np.float32),
        done=out["done"], contact_pre=out["contact_pre"],
        energy=out["energy"].astype(np.float32), centroid=out["centroid"].astype(np.float32),
        steps=out["steps"].astype(np.int32), reset_noise=out["reset_noise"].astype(np.float32),
        actions=np.asarray(actions, np.float32),
        spec=np.array(json.dumps(spec)), env_kwargs=np.array(json.dumps(env_kwargs)),
    )
    # every value the reference produced must be float32-representable (or non-finite)
    for k in ("x", "obs", "reward", "energy", "centroid", "reset_noise"):
        a64, a32 = np.asarray(out[k], np.float64), arrs[k].astype(np.float64)
        assert np.array_equal(a64, a32, equal_nan=True), f"{name}:{k} not float32-representable"
    if extra:
        arrs.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(f"{name}: T={len(actions)} done={int(out['done'].sum())} "
          f"contact={int(out['contact_pre'].sum())} nonfinite={int((~np.isfinite(out['pos'])).sum())}")


def case(name, body, T, seed, env_kwargs=None, **kw):
    env_kwargs = dict(env_kwargs or {})
    rng = np.random.default_rng(1000 + seed)
    actions = rng.uniform(-1, 1, (T, n_muscles(body))).astype(np.float32)
    extra = {}
    if "noise" not in kw:   # draw the reset jitter ourselves so it is float32-exact
        nrng = np.random.default_rng(5000 + seed)
        sigma = env_kwargs.get("rand_sigma", 0.1)
        kw["noise"] = (nrng.standard_normal(4096) * sigma).astype(np.float32)
    for k in ("k_sub", "max_steps", "reset_on_done"):
        if k in kw:
            extra[k] = np.array(int(kw[k]))
    if "integrator" in kw:
        extra["integrator"] = np.array(1 if kw["integrator"] == "run2" else 0)
    if "init_state" in kw:
        extra["init_pos"] = np.asarray(kw["init_state"]["pos"], np.float32)
        extra["init_vel"] = np.asarray(kw["init_state"]["vel"], np.float32)
    out = rh.rollout(body, actions, env_kwargs=env_kwargs, seed=seed, **kw)
    save(name, out, body, env_kwargs, actions, extra)


def batch_case(name, body, E, seed, env_kwargs):
    """E independent envs, one step each, from perturbed states (config-2 style)."""
    spec = body if isinstance(body, dict) else {"balance-v0": wo.BALANCE, "box-v0": wo.BOX}[body.lower()]
    N, M = len(spec["points"]), len(spec["muscles"])
    rng = np.random.default_rng(9000 + seed)
    tmpl = np.array([p[1] for p in spec["points"]], np.float32)
    pos0 = (tmpl[None] + rng.normal(0, 8.0, (E, N, 3))).astype(np.float32)
    pos0[:, :, 1] -= rng.uniform(0, 30, (E, 1)).astype(np.float32)      # push some masses under ground
    vel0 = rng.normal(0, 20.0, (E, N, 3)).astype(np.float32)
    if not env_kwargs.get("in3d", False):
        pos0[:, :, 2] = 0
        vel0[:, :, 2] = 0
    actions = rng.uniform(-1, 1, (E, M)).astype(np.float32)
    keys = ("pos", "vel", "old_a", "x", "obs", "reward", "done", "contact_pre", "energy", "centroid")
    acc = {k: [] for k in keys}
    for e in range(E):
        out = rh.rollout(body, actions[e:e + 1], env_kwargs=env_kwargs, seed=seed,
                         noise=np.zeros(64, np.float32), init_state=dict(pos=pos0[e], vel=vel0[e]))
        for k in keys:
            acc[k].append(out[k][-1])
    arrs = {k: np.stack(v) for k, v in acc.items()}
    for k in ("x", "obs", "reward", "energy", "centroid"):
        a32 = arrs[k].astype(np.float32)
        assert np.array_equal(arrs[k].astype(np.float64), a32.astype(np.float64), equal_nan=True)
        arrs[k] = a32
    np.savez_compressed(os.path.join(HERE, name + ".npz"), init_pos=pos0, init_vel=vel0, actions=actions,
                        spec=np.array(json.dumps(spec)), env_kwargs=np.array(json.dumps(env_kwargs)), **arrs)
    print(f"{name}: E={E} done={int(arrs['done'].sum())} contact={int(arrs['contact_pre'].sum())}")


def main():
    assert rh.available(), "reference checkout missing"
    # the two in-tree morphologies, 3-D and 2-D, two seeds
    case("balance3d_s0", "Balance-v0", 120, 0, dict(in3d=True))
    case("balance2d_s1", "Balance-v0", 120, 1, dict(in3d=False))
    case("box3d_s0", "Box-v0", 120, 0, dict(in3d=True))
    case("box2d_s1", "Box-v0", 120, 1, dict(in3d=False))
    # long run through overflow to inf/NaN (as-written dynamics are unstable)
    case("box3d_overflow", "Box-v0", 400, 2, dict(in3d=True))
    # irregular body and non-default env parameters
    case("custom3d", CUSTOM, 150, 3, dict(in3d=True, g=9.8, dampk=0.5, ground_high=-10, ground_k=500,
                                          ground_damp=30, friction=20, rand_sigma=0.3))
    case("custom2d", CUSTOM, 60, 4, dict(in3d=False, g=30, dampk=0.0, ground_high=2.5))
    # N=13 body: exercises NumPy's 8-accumulator pairwise path
    case("insect3d", INSECT, 100, 5, dict(in3d=True))
    # termination branches
    case("done_fall", "Balance-v0", 80, 6, dict(in3d=True, ground_k=0, ground_damp=0, friction=0))
    case("done_stopped", RESTING, 130, 7, dict(in3d=True, rand_sigma=0.01))
    case("done_maxsteps", "Box-v0", 30, 8, dict(in3d=True), max_steps=20)
    # jitter auto-reset (env.reset() after done) and physics substeps
    case("autoreset_jitter", "Balance-v0", 40, 9, dict(in3d=True), max_steps=7, reset_on_done=True)
    case("autoreset_jitter2d", "Box-v0", 40, 10, dict(in3d=False), max_steps=5, reset_on_done=True)
    case("substeps4_box", "Box-v0", 40, 11, dict(in3d=True), k_sub=4)
    case("substeps8_insect", INSECT, 20, 12, dict(in3d=True), k_sub=8)
    # the reference's second integrator, Point.run2, in place of run1
    case("run2_balance3d", "Balance-v0", 60, 14, dict(in3d=True), integrator="run2")
    # physically-signed springs == negative k: stable for 300 steps
    phys_box = json.loads(json.dumps(wo.BOX))
    for grp in ("muscles", "skeletons"):
        phys_box[grp] = [(i, j, {"k": -1000}) for i, j, _ in phys_box[grp]]
    case("box3d_physical_sign", phys_box, 300, 13, dict(in3d=True))
    # config-2 style batches: independent envs, one step from perturbed states
    batch_case("batch_balance3d", "Balance-v0", 256, 20, dict(in3d=True))
    batch_case("batch_box3d", "Box-v0", 256, 21, dict(in3d=True))
    batch_case("batch_box2d", "Box-v0", 128, 22, dict(in3d=False))
    # legacy-signature Environment: creature.act(a); env.step(t) with a caller-chosen t
    rng = np.random.default_rng(31)
    acts = rng.uniform(-1, 1, (50, 4)).astype(np.float32)
    nz = (rng.standard_normal(64) * 0.1).astype(np.float32)
    kw = dict(in3d=True, g=50, groundk=800)
    out = rh.rollout_compat("Box-v0", acts, 0.005, env_kwargs=kw, noise=nz)
    np.savez_compressed(os.path.join(HERE, "compat_environment.npz"), pos=out["pos"], vel=out["vel"],
                        old_a=out["old_a"], x=out["x"].astype(np.float32), contact_pre=out["contact_pre"],
                        reset_noise=out["reset_noise"].astype(np.float32), actions=acts, t_step=np.array(0.005),
                        spec=np.array(json.dumps(wo.BOX)), env_kwargs=np.array(json.dumps(kw)))
    print("compat_environment: T=50")
    # a snapshot written by the reference's own Point.snapshot (gym/optimized_engine.py:319-324)
    engine, walker, _ = rh.load()
    engine.Point.clear()
    c = walker.create_box_creature()
    c.phys[0].v[:] = [1.5, -2.0, 0.25]
    c.phys[2].pos[:] = [51.0, 99.5, -0.5]
    engine.Point.snapshot(os.path.join(HERE, "state_ref_box.pkl"))
    engine.Point.clear()


def f64_case(name, body, T, seed, env_kwargs, scale=1.0, n_act=None, **kw):
    """The reference driven the way its own demo loop drives it (gym/performance_demo.py:241-262): float64 ndarray
    actions.  Muscle.x then silently becomes float64 and the muscle spring term is evaluated in double:
    this is NOT reproduced bit for bit (the device computes in float32 on float32 actions); the fixture pins the
    tolerance protocol of SURVEY 7.2 instead -- teacher-forced single steps within 1e-5, flags exact."""
    spec = body if isinstance(body, dict) else {"balance-v0": wo.BALANCE, "box-v0": wo.BOX}[body.lower()]
    rng = np.random.default_rng(7000 + seed)
    actions = rng.uniform(-1, 1, (T, n_act or n_muscles(body))) * scale      # float64
    noise = (np.random.default_rng(8000 + seed).standard_normal(4096) * env_kwargs.get("rand_sigma", 0.1)).astype(np.float32)
    out = rh.rollout(body, actions, env_kwargs=env_kwargs, seed=seed, noise=noise, **kw)
    assert out["x"].dtype == np.float64 and not np.array_equal(out["x"], out["x"].astype(np.float32).astype(np.float64))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pos=out["pos"], vel=out["vel"], old_a=out["old_a"],
                        x=out["x"], obs=out["obs"], reward=out["reward"], done=out["done"], contact_pre=out["contact_pre"],
                        steps=out["steps"].astype(np.int32), reset_noise=out["reset_noise"].astype(np.float32),
                        actions=actions, spec=np.array(json.dumps(spec)), env_kwargs=np.array(json.dumps(env_kwargs)),
                        **{k: np.array(int(v)) for k, v in kw.items() if k in ("max_steps", "reset_on_done", "k_sub")})
    print(f"{name}: T={T} float64 actions, done={int(out['done'].sum())} contact={int(out['contact_pre'].sum())} "
          f"nonfinite={int((~np.isfinite(out['pos'])).sum())}")


def main_f64():
    f64_case("f64act_balance3d", "Balance-v0", 100, 0, dict(in3d=True))
    f64_case("f64act_box2d", "Box-v0", 100, 1, dict(in3d=False))
    phys_box = json.loads(json.dumps(wo.BOX))
    for grp in ("muscles", "skeletons"):
        phys_box[grp] = [(i, j, {"k": -1000}) for i, j, _ in phys_box[grp]]
    f64_case("f64act_box3d_physical_sign", phys_box, 100, 2, dict(in3d=True))
    # clamps on most steps (narrow limits, large actions), a python-float rest length, float masses, a DingPoint,
    # fewer action columns than muscles, physically signed springs so the run stays finite
    clampy = {"points": [(2.5, (-30, 40, 5), False), (0.7, (35, 60, -3), False), (7, (0, 5, 0), False),
                         (1, (10, 90, 2), True), (3, (-60, 20, 0), False)],
              "muscles": [(0, 2, {"k": -800, "dampk": 15, "minl": 0.9, "maxl": 1.05}),
                          (1, 2, {"x": 70.3, "k": -1200.5, "dampk": 25, "minl": 0.95, "maxl": 1.1}),
                          (4, 0, {"k": -300, "minl": 0.5, "maxl": 2.0}), (3, 1, {"k": -150})],
              "skeletons": [(0, 1, {"k": -500}), (1, 3, {"k": -2000, "dampk": 5}), (4, 2, {"x": 61.5, "k": -900})]}
    f64_case("f64act_clamps_custom3d", clampy, 150, 3, dict(in3d=True, g=30, ground_high=-10), scale=6.0, n_act=3)
    # substeps and jitter auto-reset under float64 actions (as-written springs)
    f64_case("f64act_autoreset_box3d", "Box-v0", 40, 4, dict(in3d=True), max_steps=6, reset_on_done=True, k_sub=1)
    f64_case("f64act_substeps_balance2d", "Balance-v0", 30, 5, dict(in3d=False), k_sub=3)


# ---- L2: the package lineage's Environment.update_physics (gym/optimized_walker/env.py:135-184) ------
L2_CHAIN = {   # a pinned rope-and-rod chain swinging into the ground: DingPoint, string springs, explicit rest length
    "points": [(1.0, (0, 50, 0), (0, 0, 0), True), (2.0, (30, 40, 0), (1, 0, 0.5), False),
               (0.5, (60, 10, 5), (0, -3, 0), False), (3, (10, -45, 0), (5, -20, 1), False)],
    "springs": [(0, 1, None, 100, False), (1, 2, None, 250, True), (0, 2, 70.0, 50, True), (2, 3, None, 1000, False)],
}
L2_BOX = {     # a free box dropped on the ground: restitution + friction branch
    "points": [(1, (-20, 0, 0), (3, 0, 0), False), (1, (20, 0, 0), (3, 0, 0), False),
               (1, (20, 40, 0), (3, 1, 0), False), (1, (-20, 40, 0), (3, 1, 0), False)],
    "springs": [(0, 1, None, 500, False), (1, 2, None, 500, False), (2, 3, None, 500, False), (3, 0, None, 500, False),
                (0, 2, None, 300, False), (1, 3, None, 300, False)],
}


def l2_case(name, system, steps, env_kwargs):
    import ref_harness_l2 as r2
    out = r2.rollout(system, steps, env_kwargs=env_kwargs)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pos=out["pos"], vel=out["vel"], old_a=out["old_a"],
                        system=np.array(json.dumps(system)), env_kwargs=np.array(json.dumps(env_kwargs)))
    print(f"{name}: steps={steps} nonfinite={int((~np.isfinite(out['pos'])).sum())} "
          f"ground_hits={int((out['pos'][:, :, 1] <= env_kwargs.get('ground_level', -50)).sum())}")


def main_l2():
    l2_case("l2_chain_default", L2_CHAIN, 300, dict())
    l2_case("l2_chain_custom", L2_CHAIN, 300, dict(ground_level=-20, gravity=(0.5, -30, 0.1), damping=0.95, air_resistance=0.2,
                                                   friction=0.3, ground_restitution=0.6, time_step=0.02))
    l2_case("l2_box_bounce", L2_BOX, 400, dict(ground_level=-10, gravity=(0, -98, 0)))
    l2_case("l2_chain_noground", L2_CHAIN, 150, dict(ground=False))
    l2_bodies()
    l2_env_state()


def l2_bodies():
    """Every body builder of gym/optimized_walker/walker.py:356-639, built by the reference: the point and spring
    tables (pins the transcription in walker_gym_b200/optimized_walker/walker.py) and a 120-step trajectory."""
    import ref_harness_l2 as r2
    out = {}
    for name in ("test", "leg2", "box", "balance1", "balance2", "balance3", "humanb", "insect"):
        rec = r2.body_rollout(name, 120)
        for k, v in rec.items():
            out[f"{name}__{k}"] = v
        print(f"l2 body {name}: points={len(rec['mass'])} springs={len(rec['si'])} muscles={len(rec['muscle_x'])}")
    np.savez_compressed(os.path.join(HERE, "l2_bodies.npz"), **out)


def l2_env_state():
    """An env_state.pkl written by the reference's own Environment.save_state (env.py:262-281), in a
    subprocess that imports the package under its real name so the pickled class path is the reference's."""
    import subprocess
    import sys
    dst = os.path.join(HERE, "env_state_ref.pkl")
    code = (
        "import sys, types\n"
        "from unittest import mock\n"
        "sys.modules['pygame'] = mock.MagicMock(name='pygame')\n"
        "sys.path.insert(0, '/root/reference/gym')\n"
        "import optimized_walker as ow\n"
        "from optimized_walker.env import Environment\n"
        "env = Environment(gravity=(0.5, -30, 0.1), damping=0.95, ground_level=-20, air_resistance=0.2,\n"
        "                  friction=0.3, ground_restitution=0.6, time_step=0.02)\n"
        "a = env.add_ding_point(1.0, [0, 50, 0], [0, 0, 0])\n"
        "b = env.add_point(2.0, [30, 40, 0], [1, 0, 0.5])\n"
        "c = env.add_point(0.5, [60, 10, 5], [0, -3, 0])\n"
        "env.add_spring(a, b, None, 100, False)\n"
        "env.add_spring(b, c, None, 250, True)\n"
        "env.add_spring(a, c, 70.0, 50, True)\n"
        "for _ in range(25): env.update_physics()\n"
        "env.renderer = None; env.scene = None\n"
        f"env.save_state({dst!r})\n"
        "import numpy as np\n"
        "for _ in range(40): env.update_physics()\n"
        f"np.savez({dst[:-4] + '_after40.npz'!r}, pos=np.array([p.pos for p in (a, b, c)]), vel=np.array([p.v for p in (a, b, c)]))\n")
    subprocess.run([sys.executable, "-c", code], check=True)
    print("env_state_ref.pkl written by the reference:", os.path.getsize(dst), "bytes")


GETSTAT_VARIANTS = [
    dict(in3d=True, pk=0.5, vk=2.0, ak=0.25, mk=3.0, midform=True, conmid=True),
    dict(in3d=False, pk=1, vk=1, ak=1, mk=1, midform=False, conmid=False),
    dict(in3d=True, pk=0.1, vk=0.01, ak=0.001, mk=0.7, midform=False, conmid=True),
    dict(in3d=True),
]


def main_getstat():
    """Creature.actdisp (discrete stride actions, gym/optimized_walker.py:37-43,169-172) driving PhysicsEnv.step, and
    Creature.getstat with non-default options (:129-162) recorded after every step, from the reference itself."""
    from unittest import mock
    engine, walker, envmod = rh.load()
    spec = json.loads(json.dumps(CUSTOM))
    for (i, j, kw), stride in zip(spec["muscles"], (2, 3, 0.7)):
        kw["stride"] = stride
        kw.pop("x", None)       # np.float32 rest lengths: a python-float Muscle.x would make `x * mk` a float64 product
    spec["points"] = [(m, tuple(p), bool(f)) for m, p, f in spec["points"]]
    T = 40
    rng = np.random.default_rng(77)
    disp = rng.integers(0, 2, (T, len(spec["muscles"]))).astype(np.uint8)
    noise = (np.random.default_rng(78).standard_normal(64) * 0.1).astype(np.float32)
    draws = []

    def fake_normal(loc=0.0, scale=1.0, size=None):
        v = float(noise[len(draws)])
        draws.append(v)
        return v

    engine.Point.clear()
    with mock.patch.object(np.random, "normal", fake_normal):
        env = envmod.PhysicsEnv(rh.build_creature(spec), in3d=True)
        cr = env.creature
        rec = {f"stat{v}": [] for v in range(len(GETSTAT_VARIANTS))}
        rec.update(pos=[], vel=[], old_a=[], x=[], reward=[], done=[])
        for t in range(T):
            cr.actdisp([bool(b) for b in disp[t]])
            _, r, d, _ = env.step([])                   # an empty action list: Creature.act touches no muscle
            for v, kw in enumerate(GETSTAT_VARIANTS):
                rec[f"stat{v}"].append(np.asarray(cr.getstat(**kw), np.float64))
            sn = rh.snapshot(env)
            for k in ("pos", "vel", "old_a", "x"):
                rec[k].append(sn[k])
            rec["reward"].append(r)
            rec["done"].append(bool(d))
    engine.Point.clear()
    arrs = {}
    for k, v in rec.items():
        a = np.stack([np.asarray(x) for x in v])
        if a.dtype == np.float64:
            a32 = a.astype(np.float32)
            assert np.array_equal(a, a32.astype(np.float64), equal_nan=True), f"getstat fixture: {k} not float32-representable"
            a = a32
        arrs[k] = a
    np.savez_compressed(os.path.join(HERE, "getstat_actdisp_custom3d.npz"), disp=disp, reset_noise=np.asarray(draws, np.float32),
                        spec=np.array(json.dumps(spec)), variants=np.array(json.dumps(GETSTAT_VARIANTS)), **arrs)
    print(f"getstat_actdisp_custom3d: T={T} draws={len(draws)} done={int(arrs['done'].sum())}")


if __name__ == "__main__":
    import warnings
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    if len(sys.argv) > 1 and sys.argv[1] == "getstat":
        main_getstat()
        sys.exit(0)
    main()
    main_f64()
    main_l2()
    main_getstat()
