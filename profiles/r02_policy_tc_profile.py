"""ncu driver: the policy kernels on BASELINE config 5's shape, a few calls each.
    python profiles/r02_policy_tc_profile.py [impl ...]     impl 0 mma.sync, 1 monolithic tcgen05, 2 warp-specialised tcgen05"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from walker_gym_b200 import _lib
from walker_gym_b200.rollout import FeatureMajorMLP, FusedPolicy
DEV, D, M, E = "cuda:0", 38, 2, 1 << 18
lib = _lib.load()
torch.manual_seed(0)
pol = FeatureMajorMLP(D, M).to(DEV)
obs = torch.randn(E, D, device=DEV) * 150.0
out = dict(action=torch.zeros(E, M, device=DEV), logp=torch.zeros(E, device=DEV), value=torch.zeros(E, device=DEV))
for impl in ([int(a) for a in sys.argv[1:]] or [2, 1, 0]):
    lib.wg_set_tuning(_lib.TUNE_POLICY_TC, impl)
    for prec in ("fp32", "tf32"):
        fp = FusedPolicy(pol, prec)
        for _ in range(3):
            fp.act(obs, obs_layout="row", act_layout="row", seed=3, step_index=9, **out)
torch.cuda.synchronize()
print("ok", lib.wg_policy_tc_status())
