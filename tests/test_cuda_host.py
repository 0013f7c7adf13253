"""GPU: the host-side plumbing around the step -- pinned buffers from the C ABI, results written by the kernel
straight into mapped host memory (zero-copy), CUDA-graph replay of closed-loop steps (StepGraph)."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_pinned_empty_is_pinned_zeroed_and_released():
    import gc
    import torch
    from walker_gym_b200.host import _live, pinned_empty
    n0 = len(_live)
    t = pinned_empty((1000, 38))
    assert t.is_pinned() and t.shape == (1000, 38) and t.dtype == torch.float32 and float(t.abs().sum()) == 0.0
    d = torch.arange(38000, dtype=torch.float32, device=DEV).reshape(1000, 38)
    t.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    assert torch.equal(t, d.cpu())
    u = pinned_empty((7,), dtype=torch.uint8, write_combined=True)
    assert u.is_pinned() and len(_live) == n0 + 2
    del t, u, d
    gc.collect()
    assert len(_live) == n0


@pytest.mark.parametrize("layout", ["packed", "soa"])
def test_zero_copy_step_writes_results_into_pinned_host_memory(layout):
    """out=(pinned host tensors): the kernel's own stores (TMA bulk stores of the observation rows on the packed path)
    land in host memory; same bits as the device buffers of a twin env."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv
    from walker_gym_b200.host import pinned_empty
    E = 4096 + 37
    kw = dict(in3d=True, auto_reset="template", max_steps=5, seed=3, state_layout=layout)
    a, b = BatchedPhysicsEnv("Balance-v0", E, DEV, **kw), BatchedPhysicsEnv("Balance-v0", E, DEV, **kw)
    h = (pinned_empty((E, a.obs_dim)), pinned_empty((E,)), pinned_empty((E,), dtype=torch.uint8))
    g = torch.Generator(device=DEV).manual_seed(0)
    for t in range(8):
        act = torch.rand(E, 2, device=DEV, generator=g) * 2 - 1
        a.step(act, out=h)
        obs, rew, done, _ = b.step(act)
        torch.cuda.synchronize()
        assert gu.same(h[0].numpy(), obs.cpu().numpy()) and gu.same(h[1].numpy(), rew.cpu().numpy()), t
        assert np.array_equal(h[2].numpy().astype(bool), done.cpu().numpy()), t
    with pytest.raises(ValueError):
        a.step(act, out=(torch.empty(E, a.obs_dim), h[1], h[2]))       # pageable host memory is not device-visible


def test_pipeline_zero_copy_matches_copy_pipeline():
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, HostStepPipeline
    from walker_gym_b200.host import pinned_empty
    E, T = 3000, 6
    kw = dict(in3d=True, auto_reset="template", max_steps=4, seed=9)
    a, b = BatchedPhysicsEnv("Box-v0", E, DEV, **kw), BatchedPhysicsEnv("Box-v0", E, DEV, **kw)
    acts = [pinned_empty((E, 4)).uniform_(-1, 1) for _ in range(T)]
    res = [[(pinned_empty((E, a.obs_dim)), pinned_empty((E,)), pinned_empty((E,), dtype=torch.uint8)) for _ in range(T)] for _ in range(2)]
    pa, pb = HostStepPipeline(a), HostStepPipeline(b)
    for t in range(T):
        pa.submit(acts[t], *res[0][t])
        pb.submit_zero_copy(acts[t], *res[1][t])
    pa.drain(); pb.drain()
    for t in range(T):
        for x, y in zip(res[0][t], res[1][t]):
            assert gu.same(x.numpy(), y.numpy()), t
    assert gu.same(a.pos.cpu().numpy(), b.pos.cpu().numpy())


@pytest.mark.parametrize("keep", ["last", "all"])
def test_step_graph_replays_closed_loop_steps_bit_exactly(keep):
    """StepGraph: T env.step launches captured once and replayed == the same steps launched one by one, including the
    Philox index of auto-reset jitter across replays (device step counter)."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, StepGraph
    E, T, R = 2048 + 5, 6, 3
    kw = dict(in3d=True, auto_reset="template", max_steps=4, seed=12, graph_safe=True)
    a, b = BatchedPhysicsEnv("Balance-v0", E, DEV, **kw), BatchedPhysicsEnv("Balance-v0", E, DEV, **kw)
    g = torch.Generator(device=DEV).manual_seed(1)
    acts = torch.rand(T, E, 2, device=DEV, generator=g) * 2 - 1
    sg = StepGraph(a, acts, keep_obs=keep)
    assert int(a._counter.item()) == int(b._counter.item())      # building the graph did not advance the env
    for r in range(R):
        acts.copy_(torch.rand(T, E, 2, device=DEV, generator=g) * 2 - 1)
        sg.replay()
        torch.cuda.synchronize()
        for t in range(T):
            obs, rew, done, _ = b.step(acts[t])
            assert gu.same(sg.rewards[t].cpu().numpy(), rew.cpu().numpy()), (r, t)
            assert np.array_equal(sg.dones[t].cpu().numpy().astype(bool), done.cpu().numpy()), (r, t)
            if keep == "all":
                assert gu.same(sg.obs[t].cpu().numpy(), obs.cpu().numpy()), (r, t)
        assert gu.same(sg.obs[-1].cpu().numpy(), obs.cpu().numpy())
    assert gu.same(a.pos.cpu().numpy(), b.pos.cpu().numpy()) and int(a._counter.item()) == int(b._counter.item())
    assert int(sg.dones.sum()) > 0
