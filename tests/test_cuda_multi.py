"""GPU: wg_step_multi (T env-steps per launch, BatchedPhysicsEnv.step_many) against the C oracle stepped T times
on the same seeded inputs, and against T single wg_step launches.  Equality is exact (float32 bit patterns): every
step of the block is the single-step kernel's device code with the state kept in registers."""
import numpy as np
import pytest

import golden_util as gu
import walker_oracle as wo
from test_cuda_vs_oracle import spec_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(name, E, *, in3d, auto_reset, max_steps, seed, k_sub=1):
    from walker_gym_b200 import BatchedPhysicsEnv
    env = BatchedPhysicsEnv(name, E, DEV, in3d=in3d, auto_reset=auto_reset, max_steps=max_steps, seed=seed, k_sub=k_sub,
                            state_layout="packed", track_stats=True, initial_reset=False)
    body = wo.make_body(spec_of(name))
    ar = {None: 0, "jitter": 1, "template": 2}[auto_reset]
    prm = wo.make_params(max_steps=max_steps, k_sub=k_sub, auto_reset=ar, seed=seed, in3d=in3d)
    return env, body, prm, wo.init_state(body, E)


@pytest.mark.parametrize("name", ["balance_v0", "box_v0", "balance", "box2"])
@pytest.mark.parametrize("in3d", [True, False])
@pytest.mark.parametrize("auto_reset", ["template", "jitter", None])
def test_step_many_matches_oracle(name, in3d, auto_reset):
    """Blocks of 1, 7 and 12 steps with episodes ending (max_steps = 5) and being reset inside a block: per-step
    rewards and dones, the observation after each block, the state and the episode statistics equal the oracle's."""
    import torch
    E = 4096 + 37                                   # a partial last tile and a partial last warp
    env, body, prm, st = _pair(name, E, in3d=in3d, auto_reset=auto_reset, max_steps=5, seed=11)
    rng = np.random.default_rng(3)
    nz = (rng.standard_normal((3 * env.N, E)) * 0.1).astype(np.float32)
    env.reset(noise=torch.from_numpy(nz).cuda(), mode="jitter")
    wo.reset(body, prm, st, mode=1, noise=nz)
    ep = (np.zeros(E, np.float32), np.zeros((4, E), np.float32))
    for T in (1, 7, 12):
        acts = rng.uniform(-1, 1, (T, E, env.M)).astype(np.float32)
        first = env.step_count
        obs, rew, done = env.step_many(torch.from_numpy(acts).cuda())
        assert env.step_count == first + T
        for t in range(T):
            prm.step_index = first + t
            out = wo.step(body, prm, st, acts[t], ep_ret=ep[0], fin_stats=ep[1])
            assert gu.same(rew[t].cpu().numpy(), out["reward"]), f"reward @ block {T} step {t}"
            assert gu.same(done[t].cpu().numpy(), out["done"].astype(bool)), f"done @ block {T} step {t}"
        assert gu.same(obs.cpu().numpy(), out["obs"]), f"obs after block {T}"
        assert gu.same(env.pos.cpu().numpy(), st["pos"]) and gu.same(env.vel.cpu().numpy(), st["vel"])
        assert gu.same(env.mx.cpu().numpy(), st["mx"]) and gu.same(env.steps.cpu().numpy(), st["steps"])
    assert gu.same(env.fin_stats.cpu().numpy(), ep[1])
    if auto_reset:
        assert ep[1][3].sum() > 0                   # episodes did end inside the blocks


def test_step_many_equals_single_steps_at_full_size():
    """BASELINE config 3 size (2^20 envs): one 16-step launch == 16 wg_step launches, graph-safe device counter
    included; a block without actions equals steps without actions."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv
    E, T = 1 << 20, 16
    kw = dict(in3d=True, auto_reset="template", max_steps=6, seed=5, state_layout="packed")
    a = BatchedPhysicsEnv("Balance-v0", E, DEV, graph_safe=True, **kw)
    b = BatchedPhysicsEnv("Balance-v0", E, DEV, **kw)
    g = torch.Generator(device=DEV).manual_seed(1)
    acts = torch.rand(T, E, 2, device=DEV, generator=g) * 2 - 1
    obs, rew, done = a.step_many(acts)
    for t in range(T):
        o, r, d, _ = b.step(acts[t])
        assert torch.equal(r.view(torch.int32), rew[t].view(torch.int32)) or gu.same(r.cpu().numpy(), rew[t].cpu().numpy()), t
        assert torch.equal(d, done[t]), t
    assert gu.same(obs.cpu().numpy(), o.cpu().numpy())
    assert torch.equal(a.state.view(torch.int32), b.state.view(torch.int32)) or gu.same(a.state.cpu().numpy(), b.state.cpu().numpy())
    assert int(a._counter.item()) == b.step_count == T + 1
    obs, rew, done = a.step_many(None, n_steps=3)
    for t in range(3):
        o, r, d, _ = b.step(None)
        assert gu.same(r.cpu().numpy(), rew[t].cpu().numpy()) and torch.equal(d, done[t])
    assert gu.same(obs.cpu().numpy(), o.cpu().numpy())


def test_step_many_substeps_and_out_buffers():
    """k_sub = 4 inside a block; results written into caller-owned trajectory slots."""
    import torch
    E, T = 1000, 9
    env, body, prm, st = _pair("box_v0", E, in3d=True, auto_reset="template", max_steps=4, seed=2, k_sub=4)
    rng = np.random.default_rng(8)
    acts = rng.uniform(-1, 1, (T, E, env.M)).astype(np.float32)
    rew = torch.full((T, E), 7.0, device=DEV)
    done = torch.zeros(T, E, dtype=torch.uint8, device=DEV)
    first = env.step_count
    obs, r2, d2 = env.step_many(torch.from_numpy(acts).cuda(), out=(rew, done))
    assert r2 is rew and d2 is done
    for t in range(T):
        prm.step_index = first + t
        out = wo.step(body, prm, st, acts[t])
        assert gu.same(rew[t].cpu().numpy(), out["reward"]) and gu.same(done[t].cpu().numpy().astype(bool), out["done"].astype(bool)), t
    assert gu.same(obs.cpu().numpy(), out["obs"])


def test_step_many_rejects_what_it_cannot_do():
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv
    soa = BatchedPhysicsEnv("Balance-v0", 256, DEV, state_layout="soa")
    with pytest.raises(ValueError):
        soa.step_many(torch.zeros(2, 256, 2, device=DEV))
    env = BatchedPhysicsEnv("Balance-v0", 256, DEV, state_layout="packed")
    with pytest.raises(ValueError):
        env.step_many(torch.zeros(2, 255, 2, device=DEV))
    with pytest.raises(ValueError):
        env.step_many(None)
    with pytest.raises(ValueError):
        env.step_many(torch.zeros(256, 2, device=DEV))                             # action repeat needs n_steps
    with pytest.raises(ValueError):
        env.step_many(torch.zeros(2, 256, 2, device=DEV), n_steps=3)               # n_steps disagrees with the block


@pytest.mark.parametrize("name", ["box", "test", "intrian", "hat", "humanb", "box4", "leg2", "leg", "balance2", "balance3"])
@pytest.mark.parametrize("in3d", [True, False])
def test_step_many_every_packed_walker_py_body(name, in3d):
    """Every body with an ahead-of-time packed kernel has the T-steps-per-launch kernel (gym/walker.py tables),
    balance2 (a 0.1 mass: full IEEE division) and balance3 (a DingPoint) included."""
    import torch
    E = 300
    env, body, prm, st = _pair(name, E, in3d=in3d, auto_reset="template", max_steps=4, seed=21)
    rng = np.random.default_rng(5)
    nz = (rng.standard_normal((3 * env.N, E)) * 0.1).astype(np.float32)
    env.reset(noise=torch.from_numpy(nz).cuda(), mode="jitter")
    wo.reset(body, prm, st, mode=1, noise=nz)
    for T in (5, 3):
        acts = rng.uniform(-1, 1, (T, E, env.M)).astype(np.float32)
        first = env.step_count
        obs, rew, done = env.step_many(torch.from_numpy(acts).cuda())
        for t in range(T):
            prm.step_index = first + t
            out = wo.step(body, prm, st, acts[t])
            assert gu.same(rew[t].cpu().numpy(), out["reward"]), f"reward @ block {T} step {t}"
            assert gu.same(done[t].cpu().numpy(), out["done"].astype(bool)), f"done @ block {T} step {t}"
        assert gu.same(obs.cpu().numpy(), out["obs"]), f"obs after block {T}"
        assert gu.same(env.pos.cpu().numpy(), st["pos"]) and gu.same(env.vel.cpu().numpy(), st["vel"])
        assert gu.same(env.mx.cpu().numpy(), st["mx"]) and gu.same(env.steps.cpu().numpy(), st["steps"])


@pytest.mark.parametrize("name", ["balance_v0", "box_v0"])
def test_step_many_action_repeat(name):
    """One [E, M] action block applied at every step of the block (frame skip) == the oracle stepped T times with it."""
    import torch
    E, T = 2000, 6
    env, body, prm, st = _pair(name, E, in3d=True, auto_reset="jitter", max_steps=4, seed=3)
    rng = np.random.default_rng(6)
    act = rng.uniform(-0.2, 0.2, (E, env.M)).astype(np.float32)
    first = env.step_count
    obs, rew, done = env.step_many(torch.from_numpy(act).cuda(), n_steps=T)
    assert tuple(rew.shape) == (T, E)
    for t in range(T):
        prm.step_index = first + t
        out = wo.step(body, prm, st, act)
        assert gu.same(rew[t].cpu().numpy(), out["reward"]) and gu.same(done[t].cpu().numpy(), out["done"].astype(bool)), t
    assert gu.same(obs.cpu().numpy(), out["obs"]) and gu.same(env.mx.cpu().numpy(), st["mx"])


def test_step_many_host_buffers():
    """wg_step_multi_host: pinned host actions in, last observation + per-step rewards / dones out == the device call."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv
    E, T = 5000, 6
    kw = dict(in3d=True, auto_reset="template", max_steps=4, seed=8, state_layout="packed")
    a = BatchedPhysicsEnv("Box-v0", E, DEV, **kw)
    b = BatchedPhysicsEnv("Box-v0", E, DEV, **kw)
    h_act = torch.empty(T, E, 4).uniform_(-1, 1).pin_memory()
    d_act = torch.empty(T, E, 4, device=DEV)
    h_obs = torch.empty(E, a.obs_dim).pin_memory()
    h_rew = torch.empty(T, E).pin_memory()
    h_done = torch.empty(T, E, dtype=torch.uint8).pin_memory()
    a.step_many_host(h_act, d_act, h_obs, h_rew, h_done)
    obs, rew, done = b.step_many(h_act.to(DEV))
    torch.cuda.synchronize()
    assert gu.same(h_obs.numpy(), obs.cpu().numpy()) and gu.same(h_rew.numpy(), rew.cpu().numpy())
    assert gu.same(h_done.numpy().astype(bool), done.cpu().numpy())
    # action repeat through host buffers
    h_one, d_one = h_act[0].contiguous().pin_memory(), torch.empty(E, 4, device=DEV)
    a.step_many_host(h_one, d_one, h_obs, h_rew[:3], h_done[:3], n_steps=3)
    obs, rew, done = b.step_many(h_one.to(DEV), n_steps=3)
    torch.cuda.synchronize()
    assert gu.same(h_obs.numpy(), obs.cpu().numpy()) and gu.same(h_rew[:3].numpy(), rew.cpu().numpy())


@pytest.mark.parametrize("masses,in3d,ding", [((1, 1, 1, 1, 1), True, ()), ((2, 5, 1, 3, 4), False, ()), ((2, 2, 1, 3, 2), True, ()),
                                               ((2.5, 0.1, 7, 1, 3), True, (3,))])
def test_step_many_runtime_specialised_user_body(masses, in3d, ding):
    """A user-built creature (no ahead-of-time kernel) gets the T-steps-per-launch kernel compiled for its spring graph
    at run time (NVRTC), like its single-step kernel: bit for bit against the oracle stepped T times."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, Point
    from test_cuda_vs_oracle import _custom_creature
    try:
        cr, spec = _custom_creature(masses, ding)
        E = 4100
        env = BatchedPhysicsEnv(cr, E, DEV, in3d=in3d, auto_reset="template", max_steps=5, k_sub=2, seed=4, initial_reset=False)
        assert env.state_layout == "packed" and env.kernel_variant == 0
        body = wo.make_body(spec)
        prm = wo.make_params(in3d=in3d, auto_reset=2, max_steps=5, k_sub=2, seed=4)
        st = wo.init_state(body, E)
        rng = np.random.default_rng(9)
        nz = (rng.standard_normal((3 * env.N, E)) * 0.1).astype(np.float32)
        env.reset(noise=torch.from_numpy(nz).cuda(), mode="jitter")
        wo.reset(body, prm, st, mode=1, noise=nz)
        for T in (8, 3):
            acts = rng.uniform(-1, 1, (T, E, env.M)).astype(np.float32)
            first = env.step_count
            obs, rew, done = env.step_many(torch.from_numpy(acts).cuda())
            for t in range(T):
                prm.step_index = first + t
                out = wo.step(body, prm, st, acts[t])
                assert gu.same(rew[t].cpu().numpy(), out["reward"]) and gu.same(done[t].cpu().numpy(), out["done"].astype(bool)), (T, t)
            assert gu.same(obs.cpu().numpy(), out["obs"])
            assert gu.same(env.pos.cpu().numpy(), st["pos"]) and gu.same(env.vel.cpu().numpy(), st["vel"])
    finally:
        Point.clear()


def test_step_many_sharding_is_invisible():
    """SURVEY 8e for the T-steps launch: shard k of the batch (env_offset) equals the same envs stepped in one piece,
    resets inside the block included (Philox keyed by the global env id and step index + t)."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv
    E, T = 8192, 24
    kw = dict(in3d=True, auto_reset="template", max_steps=7, seed=123, track_stats=True, state_layout="packed")
    whole = BatchedPhysicsEnv("balance_v0", E, DEV, **kw)
    halves = [BatchedPhysicsEnv("balance_v0", E // 2, DEV, env_offset=k * E // 2, **kw) for k in range(2)]
    g = torch.Generator(device=DEV).manual_seed(0)
    acts = torch.rand(T, E, 2, device=DEV, generator=g) * 2 - 1
    obs, rew, done = whole.step_many(acts)
    parts = [h.step_many(acts[:, k * E // 2:(k + 1) * E // 2].contiguous()) for k, h in enumerate(halves)]
    assert gu.same(obs.cpu().numpy(), torch.cat([p[0] for p in parts], 0).cpu().numpy())
    assert gu.same(rew.cpu().numpy(), torch.cat([p[1] for p in parts], 1).cpu().numpy())
    assert torch.equal(done, torch.cat([p[2] for p in parts], 1))
    for name in ("pos", "vel", "mx", "fin_stats"):
        cat = torch.cat([getattr(h, name) for h in halves], dim=1)
        assert gu.same(getattr(whole, name).cpu().numpy(), cat.cpu().numpy()), name
    a, b0, b1 = whole.episode_stats(), halves[0].episode_stats(), halves[1].episode_stats()
    assert a["episodes"] == b0["episodes"] + b1["episodes"] > 0


def test_host_pipeline_submit_many():
    """HostStepPipeline.submit_many: double-buffered T-step launches on pinned host buffers == step_many block by block."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, HostStepPipeline
    E, T, B = 3000, 5, 4
    kw = dict(in3d=True, auto_reset="template", max_steps=4, seed=8, state_layout="packed")
    a = BatchedPhysicsEnv("Balance-v0", E, DEV, **kw)
    b = BatchedPhysicsEnv("Balance-v0", E, DEV, **kw)
    pipe = HostStepPipeline(a)
    h_act = [torch.empty(T, E, 2).uniform_(-1, 1).pin_memory() for _ in range(B)]
    h_rew = [torch.empty(T, E).pin_memory() for _ in range(B)]
    h_done = [torch.empty(T, E, dtype=torch.uint8).pin_memory() for _ in range(B)]
    h_obs = [torch.empty(E, a.obs_dim).pin_memory() for _ in range(B)]
    for i in range(B):
        pipe.submit_many(h_act[i], h_rew[i], h_done[i], h_obs[i] if i % 2 == 0 else None)
    pipe.drain()                    # blocks the host until the results are in the host tensors
    for i in range(B):
        obs, rew, done = b.step_many(h_act[i].to(DEV))
        assert gu.same(h_rew[i].numpy(), rew.cpu().numpy()) and gu.same(h_done[i].numpy().astype(bool), done.cpu().numpy()), i
        if i % 2 == 0:
            assert gu.same(h_obs[i].numpy(), obs.cpu().numpy()), i
    assert gu.same(a.state.cpu().numpy(), b.state.cpu().numpy())


def _gen_cases(M, dt):
    from walker_gym_b200 import CPGActions, ScriptedActions
    rng = np.random.default_rng(M)
    table = rng.uniform(-1, 1, (3, M)).round(2).tolist()
    return [ScriptedActions(table, hold=2), ScriptedActions([[0.5] * M], hold=1),
            CPGActions(amp=rng.uniform(0.2, 3.0, M).tolist(), freq=rng.uniform(0.3, 9.0, M).tolist(),
                       phase=rng.uniform(-3, 3, M).tolist())]


@pytest.mark.parametrize("name,in3d", [("balance_v0", True), ("box_v0", False), ("hat", True), ("balance3", True)])
def test_step_many_with_in_kernel_action_sources_matches_oracle(name, in3d):
    """Scripted phase table (gym/walker.py:356-366) and sinusoidal CPG (after gym/optimized_walker/walker.py:56-90)
    evaluated inside the T-steps kernel from each env's own step counter == the oracle stepped with the same generated
    actions; episodes end and restart the gait inside a block."""
    import torch
    E = 2048 + 19
    for gi in range(3):
        env, body, prm, st = _pair(name, E, in3d=in3d, auto_reset="template", max_steps=7, seed=31 + gi)
        src = _gen_cases(env.M, 0.01)[gi]
        desc = src.describe() if gi < 2 else src.describe(0.01)
        rng = np.random.default_rng(gi)
        nz = (rng.standard_normal((3 * env.N, E)) * 0.1).astype(np.float32)
        env.reset(noise=torch.from_numpy(nz).cuda(), mode="jitter")
        wo.reset(body, prm, st, mode=1, noise=nz)
        # desynchronise the step counters so that one launch sees many gait phases
        steps0 = rng.integers(0, 6, E).astype(np.int32)
        env.set_state(steps=torch.from_numpy(steps0))
        st["steps"][:] = steps0
        for T in (1, 9, 16):
            first = env.step_count
            obs, rew, done = env.step_many(src, n_steps=T)
            for t in range(T):
                prm.step_index = first + t
                acts = wo.gen_actions(desc, st["steps"], env.M)
                out = wo.step(body, prm, st, acts)
                assert gu.same(rew[t].cpu().numpy(), out["reward"]), (gi, T, t)
                assert gu.same(done[t].cpu().numpy(), out["done"].astype(bool)), (gi, T, t)
            assert gu.same(obs.cpu().numpy(), out["obs"]), (gi, T)
            assert gu.same(env.pos.cpu().numpy(), st["pos"]) and gu.same(env.mx.cpu().numpy(), st["mx"]), (gi, T)
            assert gu.same(env.steps.cpu().numpy(), st["steps"])
        assert np.unique(wo.gen_actions(desc, np.arange(64, dtype=np.int32), env.M), axis=0).shape[0] > (1 if gi != 1 else 0)


def test_in_kernel_action_sources_on_a_run_time_compiled_body_and_validation():
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, CPGActions, Creature, Muscle, Point, ScriptedActions, Skeleton
    Point.clear()
    try:
        pts = [Point(1, [0, 0, 0], [0, 0, 0]), Point(1, [60, 0, 0], [0, 0, 0]), Point(1, [30, 50, 0], [0, 0, 0]),
               Point(1, [30, 100, 0], [0, 0, 0]), Point(1, [80, 90, 0], [0, 0, 0])]
        mus = [Muscle(pts[0], pts[2]), Muscle(pts[1], pts[2]), Muscle(pts[3], pts[4])]
        sks = [Skeleton(pts[0], pts[1]), Skeleton(pts[2], pts[3]), Skeleton(pts[2], pts[4])]
        cr = Creature(pts, mus, sks)
        E, T = 4096, 12
        kw = dict(in3d=True, auto_reset="template", max_steps=5, seed=3, state_layout="packed")
        a, b = BatchedPhysicsEnv(cr, E, DEV, **kw), BatchedPhysicsEnv(cr, E, DEV, **kw)
        src = CPGActions(amp=[1.0, 2.0, 0.5], freq=[2.0, 3.5, 7.0], phase=[0.0, 1.0, 2.0])
        obs, rew, done = a.step_many(src, n_steps=T)
        # the same generator evaluated on the host (float32 ops as written in wg_action_gen) and fed as a tensor
        import walker_oracle as wo2
        desc = src.describe(0.01)
        for t in range(T):
            acts = wo2.gen_actions(desc, b.steps.cpu().numpy(), 3)
            o, r, d, _ = b.step(torch.from_numpy(acts).to(DEV))
            assert gu.same(rew[t].cpu().numpy(), r.cpu().numpy()) and torch.equal(done[t], d), t
        assert gu.same(obs.cpu().numpy(), o.cpu().numpy())
        with pytest.raises(ValueError):
            a.step_many(ScriptedActions([[0.0, 0.0]]), n_steps=2)             # 2 values for 3 muscles
        with pytest.raises(ValueError):
            a.step_many(src)                                                   # n_steps missing
        with pytest.raises(ValueError):
            ScriptedActions([[0.0]] * 33)
    finally:
        Point.clear()
