"""GPU: the drop-in single-env surface (make_env / PhysicsEnv / Environment) replayed
against the reference's recorded trajectories.  Written the way a reference test
would read: build the env, seed, step, compare what step() returns and what the
creature's Point/Muscle objects show afterwards."""
from unittest import mock

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def feed_normal(draws):
    it = iter(np.asarray(draws, np.float64).tolist())
    return mock.patch.object(np.random, "normal", lambda loc=0.0, scale=1.0, size=None: next(it))


@pytest.mark.parametrize("name,env_id", [("balance3d_s0", "Balance-v0"), ("balance2d_s1", "Balance-v0"),
                                         ("box3d_s0", "Box-v0"), ("box2d_s1", "Box-v0"), ("box3d_overflow", "Box-v0"),
                                         ("done_fall", "Balance-v0"), ("done_maxsteps", "Box-v0")])
def test_make_env_step_matches_reference(name, env_id):
    from walker_gym_b200 import Point, make_env
    g = gu.load(name)
    Point.clear()
    with feed_normal(g["reset_noise"]):
        env = make_env(env_id, **g["env_kwargs"])
    if g["max_steps"] is not None:
        env.max_steps = g["max_steps"]
    assert env.time_step == 0.01 and env.steps == 0
    obs0 = env._body.core.obs[0].cpu().numpy()
    assert gu.same(obs0, g["obs"][0])
    assert env.get_observation_space()["shape"] == (g["obs"].shape[1],)
    assert env.get_action_space()["shape"] == (g["actions"].shape[1],)
    T = min(len(g["actions"]), 150)
    for t in range(T):
        obs, reward, done, info = env.step(g["actions"][t])
        assert isinstance(obs, np.ndarray) and obs.dtype == np.float64 and isinstance(done, bool)
        assert gu.same(obs.astype(np.float32), g["obs"][t + 1]), f"obs @ {t}"
        assert gu.same(np.float32(reward), g["reward"][t]), f"reward @ {t}"
        assert done == bool(g["done"][t]), f"done @ {t}"
        assert info["steps"] == t + 1 == env.steps
        assert gu.same(np.float32(info["total_energy"]), g["energy"][t]), f"energy @ {t}"
        assert gu.same(np.float32(info["centroid_position"]), g["centroid"][t]), f"centroid @ {t}"
        c = env.creature
        assert gu.same(np.stack([p.pos for p in c.phys]), g["pos"][t + 1])
        assert gu.same(np.stack([p.v for p in c.phys]), g["vel"][t + 1])
        assert gu.same(np.stack([p.old_a for p in c.phys]), g["old_a"][t + 1])
        assert gu.same(np.float32([m.x for m in c.muscles]), g["x"][t + 1])
        assert [p.r == 3 for p in c.phys] == list(g["contact_pre"][t])
        assert [p.color == "red" for p in c.phys] == list(g["contact_pre"][t])
        assert gu.same(np.array(c.getstat(env.in3d), np.float32), g["obs"][t + 1])
    Point.clear()


def test_reset_is_jitter_only_like_the_reference():
    from walker_gym_b200 import Point, make_env
    g = gu.load("autoreset_jitter")
    Point.clear()
    with feed_normal(g["reset_noise"]):
        env = make_env("Balance-v0", **g["env_kwargs"])
        env.max_steps = g["max_steps"]
        for t in range(len(g["actions"])):
            obs, reward, done, info = env.step(g["actions"][t])
            assert done == bool(g["done"][t])
            if done:
                obs = env.reset()
                assert env.steps == 0
            assert gu.same(obs.astype(np.float32), g["obs"][t + 1]), f"obs @ {t}"
            assert gu.same(np.stack([p.pos for p in env.creature.phys]), g["pos"][t + 1])
    assert env.seed(5) == [5] and env.seed() == []
    Point.clear()


def test_user_mutations_between_steps_are_honoured():
    """The reference mutates the caller's Point objects in place, and reads them on every step."""
    from walker_gym_b200 import Point, make_env
    g = gu.load("batch_box3d")
    Point.clear()
    with feed_normal(np.zeros(64)):
        env = make_env("Box-v0", in3d=True)
    for e in range(8):
        for n, p in enumerate(env.creature.phys):
            p.pos[:], p.v[:] = g["init_pos"][e][n], g["init_vel"][e][n]
        for m in env.creature.muscles:
            m.x = m.originx
        env.steps = 0
        obs, reward, done, info = env.step(g["actions"][e])
        assert gu.same(obs.astype(np.float32), g["obs"][e]) and gu.same(np.float32(reward), g["reward"][e])
        assert gu.same(np.stack([p.pos for p in env.creature.phys]), g["pos"][e])
    Point.clear()


def test_compat_environment_step_t():
    from walker_gym_b200 import Environment, Point, create_box_creature
    g = gu.load("compat_environment") if False else None
    z = np.load(gu.GOLDEN_DIR + "/compat_environment.npz")
    import json
    kw = json.loads(str(z["env_kwargs"]))
    Point.clear()
    c = create_box_creature()
    with feed_normal(z["reset_noise"]):
        env = Environment([c], **kw)
    t_step = float(z["t_step"])
    for t, a in enumerate(z["actions"]):
        c.act(a)
        if t % 2 == 0:
            assert env.step(t_step) is None
        else:                       # the legacy two-call form
            env.run()
            Point.run1(t_step)
        assert gu.same(np.stack([p.pos for p in c.phys]), z["pos"][t + 1]), f"pos @ {t}"
        assert gu.same(np.stack([p.v for p in c.phys]), z["vel"][t + 1]), f"vel @ {t}"
        assert gu.same(np.stack([p.old_a for p in c.phys]), z["old_a"][t + 1]), f"old_a @ {t}"
        assert gu.same(np.float32([m.x for m in c.muscles]), z["x"][t + 1])
    Point.clear()


def test_batched_save_state_roundtrip(tmp_path):
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv
    from walker_gym_b200.state_io import load_points
    env = BatchedPhysicsEnv("Box-v0", 64, "cuda:0", in3d=True, keep_old_a=True)
    for _ in range(5):
        env.step(torch.rand(64, 4, device="cuda:0") * 2 - 1)
    path = str(tmp_path / "state.pkl")
    env.save_state(path, env_index=17)
    pts, _ = load_points(path)
    assert gu.same(np.stack([p.pos for p in pts]).reshape(-1), env.pos[:, 17].cpu().numpy())
    assert gu.same(np.stack([p.v for p in pts]).reshape(-1), env.vel[:, 17].cpu().numpy())
    other = BatchedPhysicsEnv("Box-v0", 8, "cuda:0", in3d=True)
    other.load_state(path)
    assert gu.same(other.pos[:, 3].cpu().numpy(), env.pos[:, 17].cpu().numpy())
    sd = env.state_dict()
    env2 = BatchedPhysicsEnv("Box-v0", 64, "cuda:0", in3d=True, keep_old_a=True)
    env2.load_state_dict(sd)
    a = torch.rand(64, 4, device="cuda:0") * 2 - 1
    env.step(a); env2.step(a)
    assert gu.same(env.pos.cpu().numpy(), env2.pos.cpu().numpy()) and gu.same(env.obs.cpu().numpy(), env2.obs.cpu().numpy())


def test_step_host_end_to_end_buffers():
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv
    E = 4096
    a = BatchedPhysicsEnv("Balance-v0", E, "cuda:0", in3d=True, seed=3)
    b = BatchedPhysicsEnv("Balance-v0", E, "cuda:0", in3d=True, seed=3)
    h_act = (torch.rand(E, 2) * 2 - 1).pin_memory()
    h_obs = torch.empty(E, a.obs_dim).pin_memory()
    h_rew, h_done = torch.empty(E).pin_memory(), torch.empty(E, dtype=torch.uint8).pin_memory()
    d_act = torch.empty(E, 2, device="cuda:0")
    a.step_host(h_act, d_act, h_obs, h_rew, h_done)
    obs, rew, done, _ = b.step(h_act.cuda())
    torch.cuda.synchronize()
    assert gu.same(h_obs.numpy(), obs.cpu().numpy()) and gu.same(h_rew.numpy(), rew.cpu().numpy())
    assert gu.same(h_done.numpy().astype(bool), done.cpu().numpy())


@pytest.mark.parametrize("name", ["f64act_balance3d", "f64act_box2d"])
def test_facade_float64_ndarray_actions_bit_exact(name):
    """The reference's own usage -- env.step(np.random.uniform(...)), float64 ndarrays -- through the drop-in facade:
    observation (float64, Muscle.x at double precision), reward, done and the Point / Muscle objects, bit for bit."""
    import walker_gym_b200 as wg
    g = gu.load(name)
    env_id = "Balance-v0" if "balance" in name else "Box-v0"
    kw = g["env_kwargs"]
    draws = list(g["reset_noise"].astype(np.float64))
    with mock.patch.object(np.random, "normal", lambda loc=0.0, scale=1.0, size=None: draws.pop(0)):
        wg.Point.clear()
        env = wg.make_env(env_id, **kw)
    for t in range(60):
        obs, rew, done, info = env.step(g["actions"][t])                 # float64 ndarray
        assert obs.dtype == np.float64 and gu.same(obs, g["obs"][t + 1]), t
        assert gu.same(np.float32(rew), np.float32(g["reward"][t])) and done == bool(g["done"][t])
        assert gu.same(np.array([p.pos for p in env.creature.phys]), g["pos"][t + 1])
        assert gu.same(np.array([float(m.x) for m in env.creature.muscles]), g["x"][t + 1])
    assert all(isinstance(m.x, (np.float64, np.float32)) for m in env.creature.muscles)
    assert any(isinstance(m.x, np.float64) for m in env.creature.muscles)
    # python-float / float32 actions afterwards keep working (Muscle.x stays np.float64: still the x64 path)
    env.step([0.1] * len(env.creature.muscles))
    env.reset()
    assert any(isinstance(m.x, np.float64) for m in env.creature.muscles)   # reset() does not touch Muscle.x


def test_batched_x64_many_envs_matches_oracle():
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv
    import walker_oracle as wo
    E, T = 3000, 25
    env = BatchedPhysicsEnv("Box-v0", E, "cuda:0", in3d=True, auto_reset="template", max_steps=9, seed=3, x64=True,
                            keep_old_a=True, initial_reset=False)
    assert env.state_layout == "soa"
    spec = wo.BOX
    body, xb = wo.make_body(spec), wo.make_x64(spec)
    prm = wo.make_params(in3d=True, auto_reset=2, max_steps=9, seed=3)
    st = wo.init_state(body, E)
    st["mx64"], st["mx_weak"] = wo.init_x64(body, xb, E)
    prm.step_index = env.step_count
    env.reset(mode="template")
    wo.reset(body, prm, st, mode=2)
    rng = np.random.default_rng(0)
    for t in range(T):
        a = rng.uniform(-40, 40, (E, 4))                                  # large: the limits clamp often
        prm.step_index = env.step_count
        obs, rew, done, _ = env.step(torch.from_numpy(a).cuda())
        out = wo.step_x64(body, xb, prm, st, a, want_info=False)
        assert gu.same(obs.cpu().numpy(), out["obs"]) and gu.same(done.cpu().numpy(), out["done"].astype(bool)), t
        assert gu.same(env.mx64.cpu().numpy(), st["mx64"]) and gu.same(env.mx_weak.cpu().numpy(), st["mx_weak"]), t
        assert gu.same(env.pos.cpu().numpy(), st["pos"]) and gu.same(env.vel.cpu().numpy(), st["vel"]), t
    assert 0 < st["mx_weak"].mean() < 1 and int(out["done"].sum()) >= 0
    with pytest.raises(ValueError):
        env.step(torch.zeros(E, 4, device="cuda:0"))                      # float32 actions are not x64 actions


def test_host_step_pipeline_matches_synchronous_host_steps():
    """HostStepPipeline (double-buffered, two streams) delivers exactly what back-to-back step_host calls deliver."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, HostStepPipeline
    E, T = 5000, 7
    kw = dict(in3d=True, auto_reset="template", max_steps=4, seed=2)
    a, b = BatchedPhysicsEnv("Balance-v0", E, "cuda:0", **kw), BatchedPhysicsEnv("Balance-v0", E, "cuda:0", **kw)
    g = torch.Generator().manual_seed(0)
    acts = [(torch.rand(E, 2, generator=g) * 2 - 1).pin_memory() for _ in range(T)]
    obs_a = [torch.empty(E, a.obs_dim).pin_memory() for _ in range(T)]
    rew_a = [torch.empty(E).pin_memory() for _ in range(T)]
    done_a = [torch.empty(E, dtype=torch.uint8).pin_memory() for _ in range(T)]
    pipe = HostStepPipeline(a)
    for t in range(T):
        pipe.submit(acts[t], obs_a[t], rew_a[t], done_a[t])
    pipe.drain()                    # blocks the host: the results are in the host tensors now (no synchronize needed)
    snap = [(o.clone(), r.clone(), d.clone()) for o, r, d in zip(obs_a, rew_a, done_a)]
    torch.cuda.synchronize()
    for t in range(T):              # ... and nothing was still in flight when drain() returned
        assert torch.equal(snap[t][0].view(torch.int32), obs_a[t].view(torch.int32)) and torch.equal(snap[t][1].view(torch.int32), rew_a[t].view(torch.int32))
        assert torch.equal(snap[t][2], done_a[t])
    d_act = torch.empty(E, 2, device="cuda:0")
    h_obs, h_rew, h_done = torch.empty(E, b.obs_dim).pin_memory(), torch.empty(E).pin_memory(), torch.empty(E, dtype=torch.uint8).pin_memory()
    for t in range(T):
        b.step_host(acts[t], d_act, h_obs, h_rew, h_done)
        torch.cuda.synchronize()
        assert gu.same(obs_a[t].numpy(), h_obs.numpy()) and gu.same(rew_a[t].numpy(), h_rew.numpy()), t
        assert gu.same(done_a[t].numpy(), h_done.numpy()), t
    assert gu.same(a.pos.cpu().numpy(), b.pos.cpu().numpy())
    # work queued on the caller's stream between drains is ordered before the next submit (reset here)
    a.reset(mode="template"); b.reset(mode="template")
    pipe.submit(acts[0], obs_a[0], rew_a[0], done_a[0])
    pipe.wait_slot(1)
    b.step_host(acts[0], d_act, h_obs, h_rew, h_done)
    torch.cuda.synchronize()
    assert gu.same(obs_a[0].numpy(), h_obs.numpy()) and gu.same(rew_a[0].numpy(), h_rew.numpy())


def test_programmatic_dependent_launch_changes_nothing():
    """The packed step kernel is launched with programmatic stream serialisation by default (WG_TUNE_PDL): its CTAs may
    be scheduled while the previous kernel of the stream drains and wait before their first global access.  Driven by
    actions that a kernel launched IMMEDIATELY before each step writes (the dependency the wait protects), eagerly and
    through a CUDA graph, the trajectory is bit-identical to plain launches."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, _lib
    lib = _lib.load()
    E, T = 1 << 16, 40
    out = {}
    for pdl in (1, 0):
        old = lib.wg_set_tuning(_lib.TUNE_PDL, pdl)
        try:
            env = BatchedPhysicsEnv("Balance-v0", E, "cuda:0", in3d=True, auto_reset="template", max_steps=25, seed=3,
                                    state_layout="packed", graph_safe=True)
            g = torch.Generator(device="cuda:0").manual_seed(11)
            base = torch.rand(E, env.M, device="cuda:0", generator=g) * 2 - 1
            act = torch.empty_like(base)
            rews = torch.zeros(T, E, device="cuda:0")
            for t in range(T // 2):                                     # eager: producer kernel, then the step
                torch.mul(base, 1.0 - 0.02 * t, out=act)
                _, r, _, _ = env.step(act)
                rews[t].copy_(r)
            graph = torch.cuda.CUDAGraph()                              # the same pair captured and replayed
            scale = torch.ones((), device="cuda:0")
            with torch.cuda.graph(graph):
                torch.mul(base, scale, out=act)
                env.step(act)
            for t in range(T // 2, T):
                scale.fill_(1.0 - 0.02 * t)
                graph.replay()
                rews[t].copy_(env.reward)
            torch.cuda.synchronize()
            out[pdl] = (rews.clone(), env.obs.clone(), env.steps.clone())
        finally:
            lib.wg_set_tuning(_lib.TUNE_PDL, old)
    for a, b in zip(out[1], out[0]):
        assert torch.equal(torch.nan_to_num(a.float()), torch.nan_to_num(b.float()))
