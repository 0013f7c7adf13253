"""PPO rollout collection around the fused step kernel (BASELINE.json config 5).

The caller of the hot path: a small torch MLP policy acts on the observation, the
CUDA kernel steps every env, and the trajectory (obs, action, log-prob, value,
reward, done) is written into preallocated device buffers.  Everything stays on
the device.  To keep the per-step launch overhead off the critical path the whole
T-step rollout (policy kernels + step kernel per step) is captured once into a
CUDA graph and replayed; the env runs in ``graph_safe`` mode so the in-kernel
Philox jitter keeps advancing across replays.  The policy consumes the
observation feature-major ([D, E]) -- the layout the kernel writes straight from
registers with coalesced stores -- and emits feature-major actions ([M, E]).
Returns/advantages (GAE) are computed on the device; episode-return statistics
are reduced by the K3 kernel and all-reduced across ranks (NCCL), once per rollout.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .batched import BatchedPhysicsEnv


class FeatureMajorMLP(torch.nn.Module):
    """obs [D, E] -> (action mean [M, E], value [E]); two tanh hidden layers of ``hidden`` units."""

    def __init__(self, obs_dim: int, act_dim: int, hidden: int = 64, obs_scale: float = 1e-2, obs_clip: float = 10.0):
        super().__init__()
        self.obs_scale, self.obs_clip = obs_scale, obs_clip
        self.l1 = torch.nn.Linear(obs_dim, hidden)
        self.l2 = torch.nn.Linear(hidden, hidden)
        self.mu = torch.nn.Linear(hidden, act_dim)
        self.v = torch.nn.Linear(hidden, 1)
        self.log_std = torch.nn.Parameter(torch.full((act_dim, 1), -0.5))

    def forward(self, obs_fm: torch.Tensor):
        # exploded envs produce inf/NaN observations: sanitise before the first layer
        x = torch.nan_to_num(obs_fm * self.obs_scale, nan=0.0, posinf=self.obs_clip, neginf=-self.obs_clip)
        x = x.clamp_(-self.obs_clip, self.obs_clip)
        h = torch.tanh(torch.addmm(self.l1.bias[:, None], self.l1.weight, x))
        h = torch.tanh(torch.addmm(self.l2.bias[:, None], self.l2.weight, h))
        mean = torch.addmm(self.mu.bias[:, None], self.mu.weight, h)
        value = torch.addmm(self.v.bias[:, None], self.v.weight, h)[0]
        return mean, value


class RolloutCollector:
    """Collect ``horizon`` steps from a ``BatchedPhysicsEnv`` under a policy.

    The env must be built with ``obs_layout="feature"``, ``act_layout="feature"`` and
    ``graph_safe=True``.  ``collect()`` returns views of device buffers:
    obs [T+1, D, E], actions [T, M, E], logp/values/rewards [T(+1), E], dones [T, E]."""

    def __init__(self, env: BatchedPhysicsEnv, policy: torch.nn.Module, horizon: int, *, gamma: float = 0.99,
                 lam: float = 0.95, use_cuda_graph: bool = True, reward_clip: float = 1e3):
        if env.obs_layout != "feature" or env.act_layout != "feature":
            raise ValueError("RolloutCollector needs obs_layout='feature' and act_layout='feature'")
        if use_cuda_graph and env._counter is None:
            raise ValueError("CUDA-graph rollouts need BatchedPhysicsEnv(graph_safe=True)")
        self.env, self.policy, self.T = env, policy, int(horizon)
        self.gamma, self.lam, self.reward_clip = gamma, lam, reward_clip
        E, D, M, dev = env.num_envs, env.obs_dim, env.M, env.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.obs = torch.zeros(self.T + 1, D, E, **f32)
        self.actions = torch.zeros(self.T, M, E, **f32)
        self.logp = torch.zeros(self.T, E, **f32)
        self.values = torch.zeros(self.T + 1, E, **f32)
        self.rewards = torch.zeros(self.T, E, **f32)
        self.dones = torch.zeros(self.T, E, dtype=torch.bool, device=dev)
        self.advantages = torch.zeros(self.T, E, **f32)
        self.returns = torch.zeros(self.T, E, **f32)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._use_graph = use_cuda_graph
        self.kernel_launches_per_rollout = self.T          # step kernels; policy kernels are torch's

    @torch.no_grad()
    def _rollout(self) -> None:
        env, pol = self.env, self.policy
        self.obs[0].copy_(env.obs)
        for t in range(self.T):
            mean, value = pol(self.obs[t])
            std = pol.log_std.exp()
            eps = torch.randn_like(mean)
            act = torch.addcmul(mean, std, eps, out=self.actions[t])
            self.logp[t] = (-0.5 * eps * eps - pol.log_std - 0.5 * math.log(2 * math.pi)).sum(0)
            self.values[t] = value
            obs, rew, done, _ = env.step(act)
            self.obs[t + 1].copy_(obs)
            r = torch.nan_to_num(rew, nan=0.0, posinf=self.reward_clip, neginf=-self.reward_clip)
            self.rewards[t] = r.clamp_(-self.reward_clip, self.reward_clip)
            self.dones[t] = done
        self.values[self.T] = pol(self.obs[self.T])[1]
        # GAE(lambda) on the device
        adv = torch.zeros_like(self.values[0])
        for t in range(self.T - 1, -1, -1):
            nonterminal = (~self.dones[t]).to(torch.float32)
            delta = self.rewards[t] + self.gamma * self.values[t + 1] * nonterminal - self.values[t]
            adv = delta + self.gamma * self.lam * nonterminal * adv
            self.advantages[t] = adv
        torch.add(self.advantages, self.values[: self.T], out=self.returns)

    def collect(self) -> Dict[str, torch.Tensor]:
        if not self._use_graph:
            self._rollout()
        else:
            if self._graph is None:
                s = torch.cuda.Stream(device=self.env.device)
                s.wait_stream(torch.cuda.current_stream(self.env.device))
                with torch.cuda.stream(s):
                    self._rollout()                               # warm-up outside capture (allocator, cuBLAS handles)
                torch.cuda.current_stream(self.env.device).wait_stream(s)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._rollout()
            self._graph.replay()
        return {"obs": self.obs, "actions": self.actions, "logp": self.logp, "values": self.values,
                "rewards": self.rewards, "dones": self.dones, "advantages": self.advantages, "returns": self.returns}

    def episode_stats(self, all_reduce: bool = True) -> dict:
        """Finished-episode return statistics of all ranks: K3 reduction + one NCCL all-reduce of 8 doubles."""
        return self.env.episode_stats(all_reduce=all_reduce)
