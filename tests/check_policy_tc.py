"""Stand-alone first-contact check of the tcgen05 / tensor-memory policy kernels (run it under `timeout`): the same
policy through the mma.sync kernel (0), the monolithic tcgen05 kernel (1) and the warp-specialised tcgen05 pipeline (2), on
the same observations and seeds; device time per call from CUDA events.
    python tests/check_policy_tc.py [E]
Exit code 0 = both agree within the float32-grade tolerance and the kernel never gave up waiting for its MMAs."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from walker_gym_b200 import _lib                                           # noqa: E402
from walker_gym_b200.rollout import FeatureMajorMLP, FusedPolicy           # noqa: E402

DEV = "cuda:0"
lib = _lib.load()
ok = True
for D, M, E in [(38, 2, 128), (38, 2, 4096 + 77), (40, 4, 1000), (26, 2, 33), (17, 1, 257), (64, 7, 5000), (9, 3, 128),
                (38, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18)]:
    torch.manual_seed(D)
    pol = FeatureMajorMLP(D, M).to(DEV)
    with torch.no_grad():
        for lin in (pol.l1, pol.l2, pol.mu, pol.v):
            lin.weight.mul_(2.0)
            lin.bias.uniform_(-0.5, 0.5)
    g = torch.Generator(device=DEV).manual_seed(E)
    obs = torch.randn(E, D, device=DEV, generator=g) * 150.0
    obs[::7, 0] = float("nan")
    obs[3::11, 1 % D] = float("inf")
    for prec, tol in (("fp32", 3e-5), ("tf32", 3e-2)):
        res = {}
        for impl in (0, 1, 2):
            lib.wg_set_tuning(_lib.TUNE_POLICY_TC, impl)
            out = dict(action=torch.zeros(E, M, device=DEV), logp=torch.zeros(E, device=DEV), value=torch.zeros(E, device=DEV),
                       mean=torch.zeros(M, E, device=DEV))
            fp = FusedPolicy(pol, prec)
            fp.act(obs, obs_layout="row", act_layout="row", seed=3, step_index=9, **out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fp.act(obs, obs_layout="row", act_layout="row", seed=3, step_index=9, **out)
            e1.record()
            torch.cuda.synchronize()
            res[impl] = (out, e0.elapsed_time(e1) / 20 * 1e3)
        lib.wg_set_tuning(_lib.TUNE_POLICY_TC, 2)
        status = lib.wg_policy_tc_status()
        a = res[0][0]
        for impl in (1, 2):
            b = res[impl][0]
            err = {k: (a[k] - b[k]).abs().max().item() for k in a}
            good = status == 0 and all(v < tol for v in err.values()) and all(torch.isfinite(b[k]).all().item() for k in b)
            ok = ok and good
            print(f"D={D} M={M} E={E} {prec} impl {impl}: max |mma.sync - tcgen05| = " + ", ".join(f"{k} {v:.2e}" for k, v in err.items()) +
                  f"; us per call {res[0][1]:.1f} -> {res[impl][1]:.1f}; tc_status {status}; {'OK' if good else 'MISMATCH'}", flush=True)
        if status != 0:
            print("the tcgen05 kernel gave up waiting for its MMAs: stopping", flush=True)
            sys.exit(2)
sys.exit(0 if ok else 1)
