"""CPU: the C oracle must reproduce the reference's recorded outputs bit for bit."""
import numpy as np
import pytest

import golden_util as gu
import walker_oracle as wo


def replay_trajectory(g, stepper):
    """Replay a golden trajectory; returns the name of the first mismatching field or None."""
    spec, kw = g["spec"], g["env_kwargs"]
    body = stepper.make_body(spec)
    N, M = body.n_mass, body.n_muscle
    in3d = bool(kw.get("in3d", False))
    d = 3 if in3d else 2
    prm_kw = dict(kw)
    if g["max_steps"] is not None:
        prm_kw["max_steps"] = g["max_steps"]
    if g["k_sub"] is not None:
        prm_kw["k_sub"] = g["k_sub"]
    if g.get("integrator"):
        prm_kw["integrator"] = g["integrator"]
    auto = 1 if g["reset_on_done"] else 0
    prm = stepper.make_params(auto_reset=auto, **prm_kw)
    st = stepper.init_state(body, 1)
    draws = g["reset_noise"].astype(np.float32)
    cursor = 0

    def next_noise():
        nonlocal cursor
        nz = np.zeros((N, 3), np.float32)
        nz[:, :d] = draws[cursor:cursor + N * d].reshape(N, d)
        cursor += N * d
        return gu.soa(nz)

    obs0 = stepper.reset(body, prm, st, mode=1, noise=next_noise())
    if "init_pos" in g:
        st["pos"][:] = gu.soa(g["init_pos"])
        st["vel"][:] = gu.soa(g["init_vel"])
    else:
        if not gu.same(obs0[0], g["obs"][0]):
            return "obs0"
    T = len(g["actions"])
    for t in range(T):
        nz = next_noise() if (auto and g["done"][t]) else None
        if nz is None and auto:
            nz = np.zeros((N * 3, 1), np.float32)
        out = stepper.step(body, prm, st, g["actions"][t:t + 1], noise=nz)
        checks = {
            "pos": gu.same(gu.aos(st["pos"], N)[0], g["pos"][t + 1]),
            "vel": gu.same(gu.aos(st["vel"], N)[0], g["vel"][t + 1]),
            "old_a": gu.same(gu.aos(st["old_a"], N)[0], g["old_a"][t + 1]),
            "x": gu.same(st["mx"][:, 0], g["x"][t + 1]),
            "obs": gu.same(out["obs"][0], g["obs"][t + 1]),
            "reward": gu.same(out["reward"][0], g["reward"][t]),
            "done": bool(out["done"][0]) == bool(g["done"][t]),
            "contact_pre": gu.same(gu.mask_bits(out["contact_pre"], N)[0], g["contact_pre"][t]),
            "energy": gu.same(out["energy"][0], g["energy"][t]),
            "centroid": gu.same(out["centroid"][:, 0], g["centroid"][t]),
            "steps": int(st["steps"][0]) == int(g["steps"][t]),
        }
        bad = [k for k, ok in checks.items() if not ok]
        if bad:
            return f"step {t}: {bad}"
    return None


@pytest.mark.parametrize("name", gu.trajectory_names())
def test_oracle_matches_reference_trajectory(name):
    g = gu.load(name)
    assert replay_trajectory(g, wo) is None


def replay_batch(g, stepper):
    spec, kw = g["spec"], g["env_kwargs"]
    body = stepper.make_body(spec)
    N = body.n_mass
    prm = stepper.make_params(**kw)
    E = g["init_pos"].shape[0]
    st = stepper.init_state(body, E)
    st["pos"][:] = gu.soa(g["init_pos"])
    st["vel"][:] = gu.soa(g["init_vel"])
    out = stepper.step(body, prm, st, g["actions"])
    checks = {
        "pos": gu.same(gu.aos(st["pos"], N), g["pos"]),
        "vel": gu.same(gu.aos(st["vel"], N), g["vel"]),
        "old_a": gu.same(gu.aos(st["old_a"], N), g["old_a"]),
        "x": gu.same(st["mx"].T, g["x"]),
        "obs": gu.same(out["obs"], g["obs"]),
        "reward": gu.same(out["reward"], g["reward"]),
        "done": gu.same(out["done"].astype(bool), g["done"]),
        "contact_pre": gu.same(gu.mask_bits(out["contact_pre"], N), g["contact_pre"]),
        "energy": gu.same(out["energy"], g["energy"]),
        "centroid": gu.same(out["centroid"].T, g["centroid"]),
    }
    return [k for k, ok in checks.items() if not ok]


@pytest.mark.parametrize("name", gu.batch_names())
def test_oracle_matches_reference_batch(name):
    assert replay_batch(gu.load(name), wo) == []


def test_golden_cover_the_branches():
    """The fixtures must actually exercise contact, every done clause, NaN and resets."""
    assert gu.load("done_fall")["done"].any()
    assert gu.load("done_stopped")["done"].any()
    assert gu.load("done_maxsteps")["done"].any()
    assert not np.isfinite(gu.load("box3d_overflow")["pos"]).all()
    assert gu.load("balance3d_s0")["contact_pre"].any()
    assert gu.load("autoreset_jitter")["done"].sum() >= 3


def replay_l2(g, stepper, chunk=1):
    """Replay an L2 (package lineage) trajectory; `chunk` steps per call."""
    sysm = stepper.make_l2_system(g["system"])
    prm = stepper.make_l2_params(**g["env_kwargs"])
    st = stepper.l2_init_state(g["system"], 1)
    P = len(g["system"]["points"])
    T = g["pos"].shape[0] - 1
    t = 0
    while t < T:
        n = min(chunk, T - t)
        stepper.l2_step(sysm, prm, st, n)
        t += n
        ok = (gu.same(st["pos"].reshape(P, 3), g["pos"][t]) and gu.same(st["vel"].reshape(P, 3), g["vel"][t])
              and gu.same(st["old_a"].reshape(P, 3), g["old_a"][t]))
        if not ok:
            return f"step {t}"
    return None


@pytest.mark.parametrize("chunk", [1, 7])
@pytest.mark.parametrize("name", gu.l2_names())
def test_oracle_matches_reference_package_physics(name, chunk):
    """L2: Environment.update_physics of gym/optimized_walker/env.py, bit for bit."""
    assert replay_l2(gu.load_l2(name), wo, chunk) is None
