#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused walker-gym physics step on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

A *step* is one pass of the hot path (PhysicsEnv.step: act -> physics -> reward /
done / auto-reset -> observation) over every env of the batch: ONE kernel launch.
Workload (BASELINE.json config 3, the throughput config; weak scaling): Balance-v0
(gym/optimized_walker.py:176-199), in3d, 2^20 envs per GPU, U(-1,1) float32
actions resident on the device, template auto-reset, reference semantics as written.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (whole box)"
UNIT = "env-steps/s"
ENV_ID = "Balance-v0"
N_MASS, N_MUSCLE = 4, 2
OBS_DIM = 3 * 3 * N_MASS + N_MUSCLE
# algorithmic bytes per env-step: state R+W 48N, muscle x R+W + action R 12M,
# steps R+W 8, reward 4, done 1; plus the materialised observation 36N + 4M
BYTES_PER_ENV_STEP = 48 * N_MASS + 12 * N_MUSCLE + 13 + 36 * N_MASS + 4 * N_MUSCLE     # 381


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the cpu_baseline sample")
    ap.add_argument("--policy", default="fused-fp32", choices=["fused-fp32", "fused-tf32", "torch"],
                    help="config 5: wg_policy_act (3xTF32 float32-grade / plain TF32 tensor-core MLP) or torch ops")
    ap.add_argument("--probe-stream", action="store_true",
                    help="measure the HBM rate of a pure streaming kernel with the step kernel's read:write mix and exit")
    ap.add_argument("--k-sub", type=int, default=0, help="physics substeps per env step (0 = the config's default)")
    ap.add_argument("--steps-per-launch", type=int, default=1, help="config 6: update_physics calls fused in one launch; config 3: env-steps per launch (wg_step_multi)")
    ap.add_argument("--pkg-body", default="box", help="config 6: body builder of gym/optimized_walker/walker.py")
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5, 6],
                    help="BASELINE.json config: 3 = Balance-v0 throughput (headline, default), 4 = enlarged body "
                         "(4x masses/springs) with 8 substeps, 5 = PPO rollout collection (torch MLP policy + step kernel)")
    ap.add_argument("--body", default="balance", choices=["balance", "box", "legacy_box", "test", "intrian", "hat", "humanb", "box4", "leg", "leg2", "insect", "quad", "balance2", "balance3"],
                    help="config 3 body: Balance-v0 (headline) or Box-v0, both from gym/optimized_walker.py:176-224")
    ap.add_argument("--obs-layout", default="row", choices=["row", "feature"],
                    help="observation layout written by the kernel: row-major [E,D] (default) or feature-major [D,E]")
    ap.add_argument("--generic", action="store_true", help="measure the run-time-topology kernel instead of the specialisation")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_env_step():
    """dram read+write bytes per env-step of the step kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_env_step"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------
# CPU arm: the oracle port (C restatement of the reference, OpenMP over envs)
# ----------------------------------------------------------------------------------
def cpu_run(n_env: int, steps: int, warmup: int, seed: int = 0):
    """Step `n_env` Balance-v0 envs `steps` times with the oracle; returns (env-steps/s, seconds, threads)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import walker_oracle as wo
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    threads = wo.set_threads(ncpu)          # all host threads (torchrun would pin OMP_NUM_THREADS=1)
    body = wo.make_body(wo.BALANCE)
    prm = wo.make_params(in3d=True, auto_reset=2, seed=seed)
    st = wo.init_state(body, n_env)
    st.pop("old_a")
    wo.reset(body, prm, st, mode=2)
    rng = np.random.default_rng(seed)
    ring = [rng.uniform(-1, 1, (n_env, N_MUSCLE)).astype(np.float32) for _ in range(8)]
    # OpenMP pool spin-up on a scratch state (untimed): the first ~20 parallel regions of a fresh pool run 10-25x
    # slower than steady state (thread creation, scheduler spreading the threads over the cores); without this a
    # short --steps run would under-report the CPU arm.  Bounded at 3 s; the timed trajectory is not touched.
    scratch = wo.init_state(body, n_env)
    scratch.pop("old_a")
    wo.reset(body, prm, scratch, mode=2)
    t_spin = time.perf_counter()
    for _ in range(60):
        wo.step(body, prm, scratch, ring[0], want_info=False)
        if time.perf_counter() - t_spin > 3.0:
            break
    del scratch
    for t in range(warmup):
        prm.step_index = t + 1
        wo.step(body, prm, st, ring[t % 8], want_info=False)
    t0 = time.perf_counter()
    for t in range(steps):
        prm.step_index = warmup + t + 1
        wo.step(body, prm, st, ring[t % 8], want_info=False)
    dt = time.perf_counter() - t0
    return n_env * steps / dt, dt, threads


def cpu_baseline(target_seconds: float):
    n_env = 1 << 15
    rate, _, threads = cpu_run(n_env, 20, 3)
    steps = max(10, int(target_seconds * rate / n_env))
    rate, dt, threads = cpu_run(n_env, steps, 3)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{ENV_ID} in3d, {n_env} envs x {steps} steps, oracle/walker_oracle.c (C restatement of the "
                      f"reference, bit-exact vs its golden vectors), OpenMP over envs, {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_env = 1 << 15
    t0 = time.perf_counter()
    rate, dt, threads = cpu_run(n_env, args.steps, max(args.warmup, 3))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ENV_ID} in3d=True, template auto-reset, U(-1,1) f32 actions; CPU sample of "
                               f"{n_env} envs per step (the GPU arm steps 2^20 per GPU)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_env} envs x {args.steps} steps, oracle/walker_oracle.c with OpenMP; the reference "
                                   "itself is Python and cannot travel to the GPU box (2.4k env-steps/s/core measured "
                                   "in the authoring container, BASELINE.md)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from walker_gym_b200 import BatchedPhysicsEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(args.warmup, 3), args.steps
    E = args.envs_per_gpu
    if args.generic:
        from walker_gym_b200 import _lib
        _lib.load().wg_force_generic(1)
    if args.config == 5:
        return run_rollout(args, rank, world, dev)
    if args.config == 6:
        return run_pkg(args, rank, world, dev)
    if args.config == 3 and args.steps_per_launch > 1:
        return run_multi(args, rank, world, dev)
    env_id = {"balance": ENV_ID, "box": "Box-v0", "legacy_box": "box"}.get(args.body, args.body)
    body, k_sub = (env_id, 1) if args.config == 3 else ("quad_balance", 8)
    if args.body == "quad":
        body = "quad_balance"
    if args.k_sub > 0:
        k_sub = args.k_sub

    env = BatchedPhysicsEnv(body, E, dev, in3d=True, auto_reset="template", seed=1234, env_offset=rank * E,
                            track_stats=True, k_sub=k_sub, obs_layout=args.obs_layout)
    bytes_per_env_step = 48 * env.N + 12 * env.M + 13 + 36 * env.N + 4 * env.M
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    ring = [(torch.rand(E, env.M, device=dev, generator=g) * 2 - 1) for _ in range(16)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(W):
        env.step(ring[t % 16])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for t in range(K):
        env.step(ring[t % 16])
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * E * K / (ms * 1e-3)
    stats = env.episode_stats(all_reduce=True)       # the only collective: 8 doubles, off the timed path

    # ---- end to end through the host-buffer C-ABI call: pinned host action in, obs/reward/done out ----
    e2e = None
    if not args.no_e2e and args.obs_layout == "row":
        h_act = torch.empty(E, env.M, dtype=torch.float32).uniform_(-1, 1).pin_memory()
        h_obs = torch.empty(E, env.obs_dim, dtype=torch.float32).pin_memory()
        h_rew = torch.empty(E, dtype=torch.float32).pin_memory()
        h_done = torch.empty(E, dtype=torch.uint8).pin_memory()
        d_act = torch.empty(E, env.M, dtype=torch.float32, device=dev)
        Ke = args.e2e_steps
        for _ in range(3):
            env.step_host(h_act, d_act, h_obs, h_rew, h_done)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(Ke):
            env.step_host(h_act, d_act, h_obs, h_rew, h_done)
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        sync_e2e = {"value": world * E * Ke / (float(t2.item()) * 1e-3), "unit": UNIT, "steps": Ke,
                    "what": "wg_step_host: upload, kernel and download back to back on one stream"}
        # the same traffic double-buffered (HostStepPipeline): step t+1's upload + kernel overlap step t's download
        from walker_gym_b200 import HostStepPipeline
        pipe = HostStepPipeline(env)
        h_obs2, h_rew2, h_done2 = h_obs.clone().pin_memory(), h_rew.clone().pin_memory(), h_done.clone().pin_memory()
        slots = ((h_obs, h_rew, h_done), (h_obs2, h_rew2, h_done2))
        for i in range(4):
            pipe.submit(h_act, *slots[i & 1])
        pipe.drain()
        barrier()
        e0.record()
        for i in range(Ke):
            pipe.submit(h_act, *slots[i & 1])
        pipe.drain()
        e1.record()
        barrier()
        t2p = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2p, op=dist.ReduceOp.MAX)
        e2e = {"value": world * E * Ke / (float(t2p.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": E * env.M * 4, "d2h_bytes_per_step": E * (env.obs_dim * 4 + 4 + 1),
               "steps": Ke,
               "what": "HostStepPipeline: every step uploads its pinned host actions and downloads observations, rewards "
                       "and dones to pinned host memory; two streams and two result slots overlap step t+1's upload + "
                       "kernel with step t's download; PCIe-bound by the 152-byte observation rows",
               "synchronous": sync_e2e}
        # the other usage mode: policy on the device -- observations stay in HBM, only reward/done go to the host
        Kd = max(Ke, 50)
        for i in range(4):
            pipe.submit(h_act, None, *slots[i & 1][1:])
        pipe.drain()
        barrier()
        e0.record()
        for i in range(Kd):
            pipe.submit(h_act, None, *slots[i & 1][1:])
        pipe.drain()
        e1.record()
        barrier()
        t3 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        e2e["device_policy_mode"] = {"value": world * E * Kd / (float(t3.item()) * 1e-3), "unit": UNIT,
                                     "h2d_bytes_per_step": E * env.M * 4, "d2h_bytes_per_step": E * 5, "steps": Kd,
                                     "what": "same pipeline with h_obs = None: observations stay on the device"}

    if rank == 0:
        peak, peak_src = hbm_peak()
        per_launch_s = ms * 1e-3 / K
        achieved = E * bytes_per_env_step / per_launch_s / 1e9
        tr = ncu_traffic_per_env_step() if (args.config == 3 and args.obs_layout == "row" and not args.generic
                                            and args.body == "balance") else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"{env_id} (gym/optimized_walker.py create_{args.body}_creature)" if args.config == 3 else
                                    "quad_balance (4x Balance-v0: N=16, S=20, M=8; BASELINE config 4)") +
                                   f", in3d=True, {E} envs per GPU, 1 kernel launch per env-step, K_sub={k_sub}, template "
                                   "auto-reset, reference semantics as written, U(-1,1) f32 actions from a 16-deep device ring",
                       "baseline_config": args.config,
                       "envs_per_gpu": E, "global_envs": world * E, "state_layout": env.state_layout, "obs": (f"row-major [E,{env.obs_dim}] materialised" if args.obs_layout == "row"
                               else f"feature-major [{env.obs_dim},E] materialised"),
                       "l2": f"state+obs+actions per step = {E * bytes_per_env_step / 1e6:.0f} MB > 126 MB L2 "
                             "(inputs larger than L2, no flush needed)",
                       "parallelism": f"env-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if tr is None else tr * E, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": bytes_per_env_step,
                         "kernel": (("wg::step_static_packed_kernel<TopoBalanceV0, in3d, row-major obs via TMA bulk store, packed "
                                     "float4 state, L2 bulk prefetch, mass pattern [k,k,1,j] at compile time>" if args.body == "balance"
                                     else f"wg::step_static_packed_kernel<{env.creature_name if hasattr(env, 'creature_name') else args.body}, "
                                          f"in3d, {env.state_layout} state>") if args.config == 3
                                    else "wg::step_units_kernel<TopoBalanceV0, in3d, P=4 lanes per env (one per Balance unit), row-major obs via TMA bulk store, MM=3>"),
                         "kernel_us": per_launch_s * 1e6},
            "e2e": e2e, "gpu_launches": K, "clocks": clocks,
            "episode_stats": {k: stats[k] for k in ("episodes", "return_mean", "length_mean")},
        }
        if args.config == 4:
            line["roofline"]["note"] = ("config 4 is fp32-issue bound, not HBM bound: 8 substeps x 20 springs per env-step "
                                        "against 1485 bytes (SURVEY 7.5); the HBM fraction is reported for completeness; "
                                        "with 1 substep the same body reaches 0.51")
        if not args.no_cpu_baseline and world == 1 and args.config == 3 and args.body == "balance":
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_multi(args, rank, world, dev):
    """Next-row workload: T env-steps per launch (wg_step_multi / BatchedPhysicsEnv.step_many) for
    actions known up front.  A bench step = one launch = --steps-per-launch env-steps of every env; value counts
    env-steps.  The state is read and written once per launch, so the kernel is bound by the instruction rate of
    the bit-exact arithmetic, not by HBM: the HBM fraction is reported for completeness."""
    import torch
    import torch.distributed as dist
    from walker_gym_b200 import BatchedPhysicsEnv
    W, K, E, T = max(args.warmup, 3), args.steps, args.envs_per_gpu, args.steps_per_launch
    env_id = {"balance": ENV_ID, "box": "Box-v0", "legacy_box": "box"}.get(args.body, args.body)
    env = BatchedPhysicsEnv(env_id, E, dev, in3d=True, auto_reset="template", seed=1234, env_offset=rank * E,
                            track_stats=True, k_sub=args.k_sub or 1, state_layout="packed")
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    ring = [(torch.rand(T, E, env.M, device=dev, generator=g) * 2 - 1) for _ in range(4)]
    out = (torch.empty(T, E, device=dev), torch.empty(T, E, dtype=torch.uint8, device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(W):
        env.step_many(ring[t % 4], out=out)
    barrier()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for t in range(K):
        env.step_many(ring[t % 4], out=out)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    tmax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    stats = env.episode_stats(all_reduce=True)

    # ---- end to end through the host-buffer C-ABI call (wg_step_multi_host): pinned host actions [T,E,M] in,
    # last observation + per-step rewards / dones out, every launch ----
    e2e = None
    if not args.no_e2e:
        def tmax_of(ms_):
            t_ = torch.tensor([ms_], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return float(t_.item())
        Ke = max(3, args.e2e_steps // 2)
        h_act = torch.empty(T, E, env.M).uniform_(-1, 1).pin_memory()
        h_obs = torch.empty(E, env.obs_dim).pin_memory()
        slots = [(torch.empty(T, E, env.M, device=dev), (torch.empty(T, E, device=dev), torch.empty(T, E, dtype=torch.uint8, device=dev)),
                  torch.empty(T, E).pin_memory(), torch.empty(T, E, dtype=torch.uint8).pin_memory()) for _ in range(2)]
        d_act, d_out, h_rew, h_done = slots[0]
        for _ in range(2):
            env.step_many_host(h_act, d_act, h_obs, h_rew, h_done, out=d_out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(Ke):
            env.step_many_host(h_act, d_act, h_obs, h_rew, h_done, out=d_out)
        e1.record()
        barrier()
        ms_sync = tmax_of(e0.elapsed_time(e1))
        h2d, d2h = T * E * env.M * 4, E * env.obs_dim * 4 + T * E * 5
        e2e = {"value": world * E * T * Ke / (ms_sync * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": Ke, "what": f"wg_step_multi_host: one bench step = one {T}-env-step launch; uploads the pinned host action "
                                    "block, downloads the last observation and the per-step rewards / dones, back to back on one stream"}
        # open-loop mode: observations stay in HBM; HostStepPipeline.submit_many overlaps block i+1's upload and kernel
        # with block i's download (two streams, two staging slots; the kernels serialise: they share the env state)
        from walker_gym_b200 import HostStepPipeline
        pipe = HostStepPipeline(env)
        hres = [(slots[i][2], slots[i][3]) for i in range(2)]

        def submit(i):
            pipe.submit_many(h_act, *hres[i & 1])
        for i in range(2):
            submit(i)
        pipe.drain()
        barrier()
        Kp = 2 * Ke
        t0 = time.perf_counter()
        for i in range(Kp):
            submit(i)
        pipe.drain()
        barrier()
        ms_pipe = tmax_of((time.perf_counter() - t0) * 1e3)
        e2e["open_loop_mode"] = {"value": world * E * T * Kp / (ms_pipe * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                                 "d2h_bytes_per_step": T * E * 5, "steps": Kp,
                                 "what": "HostStepPipeline.submit_many (step_many on two streams with pinned-host copies around "
                                         "it): observations stay on the device, rewards / dones come back every launch; "
                                         "timed host-side across a full drain"}
    if rank == 0:
        peak, peak_src = hbm_peak()
        bytes_per_launch_env = 48 * env.N + 8 * env.M + 8 + 36 * env.N + 4 * env.M + T * (4 * env.M + 5)
        per_launch_s = ms * 1e-3 / K
        achieved = E * bytes_per_launch_env / per_launch_s / 1e9
        line = {"metric": METRIC, "value": world * E * K * T / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{env_id}, in3d=True, {E} envs per GPU, {T} env-steps per launch (wg_step_multi: actions "
                                       f"[T,E,M] known up front, per-step reward / done, observation after the last step), "
                                       "template auto-reset, U(-1,1) f32 actions from a 4-deep device ring of blocks",
                           "baseline_config": "next row f2", "envs_per_gpu": E, "global_envs": world * E,
                           "steps_per_launch": T, "us_per_env_step": per_launch_s * 1e6 / T,
                           "l2": f"state+obs+actions per launch = {E * bytes_per_launch_env / 1e6:.0f} MB > 126 MB L2"},
                "roofline": {"bound": "fp32 issue", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_env_launch": bytes_per_launch_env,
                             "kernel": "wg::step_multi_packed_kernel", "kernel_us": per_launch_s * 1e6},
                "e2e": e2e, "gpu_launches": K, "clocks": clocks,
                "episode_stats": {k: stats[k] for k in ("episodes", "return_mean", "length_mean")}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_pkg(args, rank, world, dev):
    """Next-row workload: the package lineage's Environment.update_physics
    (gym/optimized_walker/env.py:135-184) on one of its own bodies, 2^20 independent copies per GPU.
    A step = one launch = --steps-per-launch updates of every env; value counts env-updates."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import walker_gym_b200.optimized_walker as ow
    W, K, E, T = max(args.warmup, 3), args.steps, args.envs_per_gpu, args.steps_per_launch
    env = ow.Environment(num_envs=E, device=str(dev), ground_level=-8.0)
    getattr(ow, args.pkg_body)(env)
    P, S = len(env._order), len(env.springs)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    env.vel.add_(torch.randn(env.vel.shape, device=dev, generator=g))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        env.update_physics(T)
    barrier()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        env.update_physics(T)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    tmax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    if rank == 0:
        peak, peak_src = hbm_peak()
        bytes_per_launch_env = 48 * P + 12 * P            # pos + vel read and written, old_a written
        per_launch_s = ms * 1e-3 / K
        achieved = E * bytes_per_launch_env / per_launch_s / 1e9
        line = {"metric": "env-updates/sec (whole box)", "value": world * E * K * T / (ms * 1e-3), "unit": "env-updates/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"package lineage Environment.update_physics, body {args.pkg_body} (P={P} points, "
                                       f"S={S} springs), {E} envs per GPU, {T} update(s) per launch, ground at -8, N(0,1) "
                                       "initial velocities", "baseline_config": "next row f3",
                           "l2": f"state per launch = {E * bytes_per_launch_env / 1e6:.0f} MB > 126 MB L2"},
                "roofline": {"bound": "hbm" if T == 1 else "fp32 issue", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_env_launch": bytes_per_launch_env,
                             "kernel": "wg::pkg_update_kernel", "kernel_us": per_launch_s * 1e6},
                "e2e": None, "gpu_launches": K, "clocks": clocks}
        if not args.no_cpu_baseline and world == 1:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import walker_oracle as wo
            threads = wo.set_threads(len(os.sched_getaffinity(0)))
            n = 1 << 14
            system = {"points": [(float(p.m), tuple(map(float, p.pos)), (0.0, 0.0, 0.0), p.fixed) for p in env._order],
                      "springs": [(env._order.index(a), env._order.index(b), float(x), float(k), bool(st))
                                  for a, b, x, k, st in env.springs]}
            so, po, st = wo.make_l2_system(system), wo.make_l2_params(ground_level=-8.0), wo.l2_init_state(system, n)
            st["vel"] += np.random.default_rng(0).normal(0, 1, st["vel"].shape).astype(np.float32)
            wo.l2_step(so, po, st, 20)
            t0 = time.perf_counter()
            wo.l2_step(so, po, st, 2000)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n * 2000 / dt, "unit": "env-updates/s", "cores": threads, "kind": "port",
                                    "sample": f"{n} envs x 2000 updates, oracle wgo_l2_step (OpenMP), {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_rollout(args, rank, world, dev):
    """BASELINE config 5: PPO rollout collection -- torch MLP policy (obs->64->64->M, tanh) + the step kernel,
    262144 envs per GPU, T=32 steps per CUDA-graph replay, episode-return statistics all-reduced over NCCL."""
    import torch
    import torch.distributed as dist
    from walker_gym_b200 import BatchedPhysicsEnv
    from walker_gym_b200.rollout import FeatureMajorMLP, RolloutCollector
    E = args.envs_per_gpu if args.envs_per_gpu != (1 << 20) else (1 << 18)
    T = 32
    torch.manual_seed(7 + rank)
    fused = args.policy != "torch"
    lay = "row" if (fused and args.obs_layout == "row") else "feature"
    env = BatchedPhysicsEnv(ENV_ID, E, dev, in3d=True, auto_reset="template", seed=1234, env_offset=rank * E,
                            obs_layout=lay, act_layout=lay, graph_safe=True)
    pol = FeatureMajorMLP(env.obs_dim, env.M).to(dev)
    col = RolloutCollector(env, pol, T, fused=fused, precision="tf32" if args.policy == "fused-tf32" else "fp32")
    n_roll = max(1, args.steps // T)
    n_warm = max(3, args.warmup // T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(n_warm):
        col.collect()
    barrier()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n_roll):
        col.collect()
    stats = col.episode_stats(all_reduce=True)         # K3 + NCCL all-reduce of 8 doubles, once per timed region
    ev1.record()
    barrier()
    clocks = sampler.stop()
    tmax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    steps = n_roll * T
    if rank == 0:
        line = {
            "metric": METRIC, "value": world * E * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": n_warm * T, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"PPO rollout collection (BASELINE config 5): torch MLP policy {env.obs_dim}->64->64->{env.M} "
                                   f"(tanh, gaussian head, value head) + fused step kernel, {ENV_ID} in3d, {E} envs per GPU, "
                                   f"T={T} steps per CUDA-graph replay, GAE on device, {lay}-major obs/actions",
                       "policy": {"fused-fp32": "wg_policy_act: the torch module's weights evaluated by one CUDA kernel per step, "
                                                "error-compensated 3xTF32 mma.sync (float32-grade, 1e-5 vs torch fp32)",
                                  "fused-tf32": "wg_policy_act with plain TF32 products and tanh.approx",
                                  "torch": "torch eager ops captured in the CUDA graph"}[args.policy],
                       "baseline_config": 5, "envs_per_gpu": E, "global_envs": world * E,
                       "parallelism": f"env-sharded x{world}; one NCCL all-reduce of 8 doubles (episode-return stats) per timed region"},
            "gpu_launches": n_roll * col.kernel_launches_per_rollout, "clocks": clocks,
            "episode_stats": {k: stats[k] for k in ("episodes", "return_mean", "length_mean")},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def probe_stream():
    """HBM rate of pure streaming kernels (wg_stream_probe) at several read : write mixes, 2^20 threads."""
    import torch
    from walker_gym_b200 import _lib
    lib, dev, n = _lib.load(), torch.device("cuda", 0), 1 << 20
    src = torch.zeros(32 * n * 4, dtype=torch.float32, device=dev)
    dst = torch.zeros(32 * n * 4, dtype=torch.float32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    out = {}
    for name, (R, W) in {"copy 16:16": (16, 16), "balance 7.5:16.8 (120 B read, 269 B written per env)": (7, 17),
                         "box 8:17.8": (8, 18), "read only 24:0": (24, 0), "write only 0:24": (0, 24)}.items():
        for _ in range(5):
            lib.wg_stream_probe(src.data_ptr(), dst.data_ptr(), n, R, W, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            lib.wg_stream_probe(src.data_ptr(), dst.data_ptr(), n, R, W, stream)
        e1.record()
        torch.cuda.synchronize()
        out[name] = round(200 * n * 16 * (R + W) / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    print(json.dumps({"probe": "wg_stream_probe, 2^20 threads x 16-byte vectors, GB/s", "hbm_peak_measured": hbm_peak()[0], **out}))


def main():
    args = parse()
    if args.probe_stream:
        return probe_stream()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
