"""GPU: the rollout-side kernels (wg_policy_act, wg_gae) against a plain PyTorch float32 reference of the same op.
Floating-point ML glue, not the bit-exact physics: tolerances are stated per test."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(params=[2, 1, 0], ids=["tcgen05-pipeline", "tcgen05-monolithic", "mma.sync"])
def policy_impl(request):
    """wg_policy_act has three implementations behind one knob (include/walker_gym_b200.h WG_TUNE_POLICY_TC); the default
    is the warp-specialised tcgen05 pipeline."""
    from walker_gym_b200 import _lib
    lib = _lib.load()
    lib.wg_set_tuning(_lib.TUNE_POLICY_TC, request.param)
    yield request.param
    assert lib.wg_policy_tc_status() == 0, "a tcgen05 policy kernel gave up waiting on one of its barriers"
    lib.wg_set_tuning(_lib.TUNE_POLICY_TC, 2)


def make_policy(D, M, seed=0):
    from walker_gym_b200.rollout import FeatureMajorMLP
    torch.manual_seed(seed)
    pol = FeatureMajorMLP(D, M).to(DEV)
    with torch.no_grad():                 # larger weights than the default init: exercises tanh away from 0
        for lin in (pol.l1, pol.l2, pol.mu, pol.v):
            lin.weight.mul_(2.0)
            lin.bias.uniform_(-0.5, 0.5)
        pol.log_std.copy_(torch.linspace(-1.0, 0.2, M)[:, None])
    return pol


def reference(pol, obs):
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            mean, value = pol(obs.double().float())
            # float64 reference of the same function, to measure both implementations against the truth
            p64 = {k: v.double() for k, v in pol.state_dict().items()}
            x = torch.nan_to_num(obs.double() * pol.obs_scale, nan=0.0, posinf=pol.obs_clip, neginf=-pol.obs_clip).clamp(-pol.obs_clip, pol.obs_clip)
            h = torch.tanh(p64["l1.weight"] @ x + p64["l1.bias"][:, None])
            h = torch.tanh(p64["l2.weight"] @ h + p64["l2.bias"][:, None])
            mean64 = p64["mu.weight"] @ h + p64["mu.bias"][:, None]
            value64 = (p64["v.weight"] @ h + p64["v.bias"][:, None])[0]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return mean, value, mean64, value64


@pytest.mark.parametrize("D,M,E", [(38, 2, 4096), (40, 4, 1000), (26, 2, 33), (17, 1, 257), (64, 7, 5000), (9, 3, 128)])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 2e-2)])
def test_policy_matches_torch_fp32(D, M, E, precision, tol, policy_impl):
    """mean / value of the fused kernel vs the torch module (float32, TF32 off) and vs a float64 evaluation.
    Tolerance: fp32 mode 2e-5 absolute (outputs are O(1)); tf32 mode 2e-2."""
    from walker_gym_b200.rollout import FusedPolicy
    pol = make_policy(D, M, seed=D)
    g = torch.Generator(device=DEV).manual_seed(E)
    obs = torch.randn(D, E, device=DEV, generator=g) * 150.0           # obs_scale 1e-2 -> O(1) inputs, some clipped at +-10
    obs[0, ::7] = float("nan")
    obs[1 % D, 3::11] = float("inf")
    obs[2 % D, 5::13] = float("-inf")
    obs[3 % D, ::5] = 5000.0
    mean_t, value_t, mean64, value64 = reference(pol, obs)
    mean, value = torch.full((M, E), 7.0, device=DEV), torch.full((E,), 7.0, device=DEV)
    action, logp = torch.zeros(M, E, device=DEV), torch.zeros(E, device=DEV)
    FusedPolicy(pol, precision).act(obs, action=action, logp=logp, value=value, mean=mean, sample=False)
    mean_r, value_r = torch.zeros_like(mean), torch.zeros_like(value)       # the same from row-major observations
    FusedPolicy(pol, precision).act(obs.t().contiguous(), obs_layout="row", value=value_r, mean=mean_r, sample=False)
    torch.cuda.synchronize()
    assert torch.equal(mean, mean_r) and torch.equal(value, value_r)
    assert torch.isfinite(mean).all() and torch.isfinite(value).all()
    assert (mean - mean_t).abs().max().item() < tol and (value - value_t).abs().max().item() < tol
    assert torch.equal(action, mean)                                    # sample=False
    if precision == "fp32":                                             # as close to the truth as torch's own float32
        err_ours = (mean.double() - mean64).abs().max().item()
        err_torch = (mean_t.double() - mean64).abs().max().item()
        assert err_ours < max(4 * err_torch, 5e-6), (err_ours, err_torch)


def test_policy_sampling_and_logp(policy_impl):
    """action = mean + std * eps with eps ~ N(0,1) (Philox per env / step / action), logp = log N(action; mean, std);
    reproducible, independent of how envs are sharded, different for every step."""
    from walker_gym_b200.rollout import FusedPolicy
    D, M, E = 38, 2, 1 << 16
    pol = make_policy(D, M)
    fp = FusedPolicy(pol, "fp32")
    obs = torch.randn(D, E, device=DEV) * 100
    out = {k: torch.zeros(M, E, device=DEV) for k in ("a", "m", "a2", "a3")}
    logp = torch.zeros(E, device=DEV)
    fp.act(obs, action=out["a"], logp=logp, mean=out["m"], seed=11, step_index=5)
    fp.act(obs, action=out["a2"], seed=11, step_index=5)
    fp.act(obs, action=out["a3"], seed=11, step_index=6)
    std = pol.log_std.detach().exp()
    eps = (out["a"] - out["m"]) / std
    assert torch.equal(out["a"], out["a2"]) and not torch.equal(out["a"], out["a3"])
    assert abs(eps.mean().item()) < 0.02 and abs(eps.var().item() - 1.0) < 0.03
    assert abs(torch.corrcoef(eps)[0, 1].item()) < 0.02                 # the two action dims are independent
    assert abs((eps ** 4).mean().item() - 3.0) < 0.15                   # gaussian kurtosis
    want = (-0.5 * eps * eps - pol.log_std.detach() - 0.5 * math.log(2 * math.pi)).sum(0)
    assert (logp - want).abs().max().item() < 1e-3
    # sharding: envs [E/2, E) evaluated as a separate shard with env_offset = E/2 draw the same noise
    half = torch.zeros(M, E // 2, device=DEV)
    fp.act(obs[:, E // 2:].contiguous(), action=half, seed=11, step_index=5, env_offset=E // 2)
    assert torch.equal(half, out["a"][:, E // 2:])
    # row-major actions
    arow = torch.zeros(E, M, device=DEV)
    fp.act(obs, action=arow, seed=11, step_index=5, act_layout="row")
    assert torch.equal(arow.t().contiguous(), out["a"])


def test_policy_unaligned_and_ragged_observations(policy_impl):
    """Row-major observations that are NOT 16-byte aligned (a view 4 bytes into a buffer) cannot be fetched by TMA bulk
    copies: the kernels gather those rows themselves; the last, ragged tile always takes that path.  Same results as
    the aligned call, bit for bit."""
    from walker_gym_b200.rollout import FusedPolicy
    D, M, E = 38, 2, 128 * 5 + 17
    pol = make_policy(D, M, seed=3)
    g = torch.Generator(device=DEV).manual_seed(1)
    big = torch.randn(E * D + 1, device=DEV, generator=g) * 120.0
    aligned = big[:E * D].clone().view(E, D)
    shifted = big[1:]                                                   # data_ptr() % 16 == 4
    shifted.copy_(aligned.reshape(-1))
    shifted = shifted.view(E, D)
    assert shifted.data_ptr() % 16 == 4 and shifted.is_contiguous()
    out = {}
    for name, obs in (("aligned", aligned), ("shifted", shifted)):
        o = dict(action=torch.zeros(E, M, device=DEV), logp=torch.zeros(E, device=DEV), value=torch.zeros(E, device=DEV),
                 mean=torch.zeros(M, E, device=DEV))
        FusedPolicy(pol, "fp32").act(obs, obs_layout="row", act_layout="row", seed=5, step_index=2, **o)
        out[name] = o
    torch.cuda.synchronize()
    for k in out["aligned"]:
        assert torch.equal(out["aligned"][k], out["shifted"][k]), k
    mean_t, value_t, _, _ = reference(pol, aligned.t().contiguous())
    assert (out["aligned"]["mean"] - mean_t).abs().max().item() < 2e-5
    assert (out["aligned"]["value"] - value_t).abs().max().item() < 2e-5


def test_policy_back_to_back_launches_are_deterministic(policy_impl):
    """200 back-to-back launches on 2^16 envs (the pipeline's barriers, tensor-memory allocation and teardown under
    load): every launch returns the bits of the first."""
    from walker_gym_b200.rollout import FusedPolicy
    D, M, E = 38, 2, 1 << 16
    pol = make_policy(D, M, seed=4)
    obs = torch.randn(E, D, device=DEV) * 120.0
    fp = FusedPolicy(pol, "fp32")
    first = dict(action=torch.zeros(E, M, device=DEV), logp=torch.zeros(E, device=DEV), value=torch.zeros(E, device=DEV))
    fp.act(obs, obs_layout="row", act_layout="row", seed=1, step_index=7, **first)
    cur = {k: torch.zeros_like(v) for k, v in first.items()}
    bad = torch.zeros((), dtype=torch.int64, device=DEV)
    for _ in range(200):
        fp.act(obs, obs_layout="row", act_layout="row", seed=1, step_index=7, **cur)
        for k in cur:
            bad += (cur[k] != first[k]).sum()
    assert bad.item() == 0


def test_gae_matches_torch():
    from walker_gym_b200.rollout import gae
    T, E = 32, 3001
    g = torch.Generator(device=DEV).manual_seed(0)
    rewards = torch.randn(T, E, device=DEV, generator=g) * 50
    rewards[3, ::17] = float("nan")
    rewards[5, ::19] = float("-inf")
    rewards[7, ::23] = 1e9
    values = torch.randn(T + 1, E, device=DEV, generator=g)
    dones = torch.rand(T, E, device=DEV, generator=g) < 0.05
    adv, ret = torch.zeros(T, E, device=DEV), torch.zeros(T, E, device=DEV)
    gamma, lam, clip = 0.99, 0.95, 1e3
    gae(rewards, values, dones.view(torch.uint8), adv, ret, gamma, lam, clip)
    r = torch.nan_to_num(rewards.double(), nan=0.0, posinf=clip, neginf=-clip).clamp(-clip, clip)
    a = torch.zeros(E, dtype=torch.float64, device=DEV)
    want = torch.zeros(T, E, dtype=torch.float64, device=DEV)
    for t in range(T - 1, -1, -1):
        nt = (~dones[t]).double()
        delta = r[t] + gamma * values[t + 1].double() * nt - values[t].double()
        a = delta + gamma * lam * nt * a
        want[t] = a
    scale = want.abs().max().item()
    assert (adv.double() - want).abs().max().item() < 1e-5 * scale      # float32 recurrence vs float64
    assert (ret.double() - (want + values[:T].double())).abs().max().item() < 1e-5 * scale


@pytest.mark.parametrize("layout", ["feature", "row"])
@pytest.mark.parametrize("graph", [False, True])
def test_fused_rollout_collector(graph, layout):
    """The fused collector (2 launches per env step) produces a trajectory that is self-consistent and whose
    physics is the bit-exact step kernel: replaying its actions through a second env gives the same obs/reward/done."""
    from walker_gym_b200 import BatchedPhysicsEnv
    from walker_gym_b200.rollout import FeatureMajorMLP, RolloutCollector
    E, T = 2048, 8
    kw = dict(in3d=True, auto_reset="template", seed=5, obs_layout=layout, act_layout=layout, graph_safe=True)
    env = BatchedPhysicsEnv("Balance-v0", E, DEV, **kw)
    torch.manual_seed(1)
    pol = FeatureMajorMLP(env.obs_dim, env.M).to(DEV)
    col = RolloutCollector(env, pol, T, use_cuda_graph=graph, fused=True, seed=9, fuse_step=True)
    # row-major Balance-v0 on the packed state: policy + env step in ONE launch (wg_policy_step, opt-in); else two launches
    assert col.fused and col.fused_step == (layout == "row")
    assert col.kernel_launches_per_rollout == (T if col.fused_step else 2 * T) + 2
    for _ in range(3 if graph else 1):               # graph: warm-up rollout, capture, replays
        out = {k: v.clone() for k, v in col.collect().items()}
    torch.cuda.synchronize()
    if not graph:
        # the same seeds through a second, identical env + collector: bit-identical trajectory
        env2 = BatchedPhysicsEnv("Balance-v0", E, DEV, **kw)
        # ... with the two separate launches (policy kernel, step kernel): the fused launch returns the same bits
        ref = RolloutCollector(env2, pol, T, use_cuda_graph=False, fused=True, seed=9, fuse_step=False)
        assert not ref.fused_step
        ref = ref.collect()
        for k in ("obs", "actions", "rewards", "dones", "values", "logp", "advantages"):
            assert torch.equal(torch.nan_to_num(out[k].float()), torch.nan_to_num(ref[k].float())), k
        # and the physics inside it is the step kernel: replay the actions through a third env, step by step
        env3 = BatchedPhysicsEnv("Balance-v0", E, DEV, **kw)
        assert torch.equal(torch.nan_to_num(env3.obs), torch.nan_to_num(out["obs"][0]))
        for t in range(T):
            o, r, d, _ = env3.step(out["actions"][t].contiguous())
            assert torch.equal(torch.nan_to_num(o), torch.nan_to_num(out["obs"][t + 1])), t
            assert torch.equal(torch.nan_to_num(r), torch.nan_to_num(out["rewards"][t])) and torch.equal(d, out["dones"][t])
    # self-consistency of the returned trajectory (both modes)
    mean = torch.zeros(env.M, E, device=DEV)
    from walker_gym_b200.rollout import FusedPolicy
    FusedPolicy(pol).act(out["obs"][2].contiguous(), mean=mean, sample=False, obs_layout=layout)
    act2 = out["actions"][2] if layout == "feature" else out["actions"][2].t()
    eps = (act2 - mean) / pol.log_std.detach().exp()
    want = (-0.5 * eps * eps - pol.log_std.detach() - 0.5 * math.log(2 * math.pi)).sum(0)
    assert (out["logp"][2] - want).abs().max().item() < 1e-3
    assert (out["returns"] - (out["advantages"] + out["values"][:T])).abs().max().item() < 1e-3
    assert not torch.equal(out["actions"][0], out["actions"][1])         # fresh noise every step
    stats = col.episode_stats(all_reduce=False)
    assert stats["episodes"] >= 0
