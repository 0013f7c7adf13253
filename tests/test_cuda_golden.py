"""GPU: the CUDA library, called through its C ABI, must reproduce the reference's
recorded outputs bit for bit -- every golden fixture, both kernels, both layouts."""
import numpy as np
import pytest

import golden_util as gu
from test_oracle_golden import (replay_batch, replay_f64_free_running, replay_f64_teacher_forced, replay_trajectory,
                                replay_x64)

pytestmark = pytest.mark.gpu


@pytest.fixture()
def stepper():
    import cuda_stepper as cs
    cs.obs_layout, cs.force_generic, cs.use_tma = 0, False, False
    yield cs
    cs.obs_layout, cs.force_generic, cs.use_tma = 0, False, False


@pytest.mark.parametrize("name", gu.trajectory_names())
def test_cuda_matches_reference_trajectory(stepper, name):
    assert replay_trajectory(gu.load(name), stepper) is None


@pytest.mark.parametrize("name", gu.trajectory_names())
def test_cuda_generic_kernel_matches_reference_trajectory(stepper, name):
    stepper.force_generic = True
    assert replay_trajectory(gu.load(name), stepper) is None


@pytest.mark.parametrize("name", ["balance3d_s0", "box2d_s1", "custom3d", "insect3d", "autoreset_jitter"])
@pytest.mark.parametrize("generic", [False, True])
def test_cuda_feature_major_obs(stepper, name, generic):
    stepper.obs_layout, stepper.force_generic = 1, generic
    assert replay_trajectory(gu.load(name), stepper) is None


@pytest.mark.parametrize("name", gu.batch_names())
@pytest.mark.parametrize("generic", [False, True])
def test_cuda_matches_reference_batch(stepper, name, generic):
    stepper.force_generic = generic
    assert replay_batch(gu.load(name), stepper) == []


@pytest.mark.parametrize("name", gu.batch_names())
@pytest.mark.parametrize("variant", ["tma", "tma_feature"])
def test_cuda_kernel_variants_match_reference_batch(stepper, name, variant):
    """The persistent TMA-pipelined kernel gives the same bits."""
    stepper.use_tma = variant.startswith("tma")
    stepper.obs_layout = 1 if variant.endswith("feature") else 0
    assert replay_batch(gu.load(name), stepper) == []


from test_oracle_golden import TOLERANCE_FIXTURES  # noqa: E402


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("name", TOLERANCE_FIXTURES)
def test_cuda_float64_actions_teacher_forced(stepper, name, generic):
    """Reference driven with float64 ndarray actions (its own demo loop): single steps within 1e-5, flags exact."""
    stepper.force_generic = generic
    worst, worst_acc, flags = replay_f64_teacher_forced(gu.load(name), stepper)
    assert flags == 0 and worst < 1e-5 and worst_acc < 1e-4, (worst, worst_acc, flags)


def test_cuda_float64_actions_free_running_100_steps(stepper):
    assert replay_f64_free_running(gu.load("f64act_box3d_physical_sign"), stepper, 100) < 1e-3


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("name", gu.f64_names())
def test_cuda_x64_mode_matches_reference_bit_for_bit(stepper, name, layout):
    """wg_step_x64: the reference driven with float64 ndarray actions (its own demo loop), bit for bit -- float64
    muscle lengths, the float32/float64 type switch at the limits, python-float rest lengths, partial actions,
    auto-reset, substeps."""
    stepper.obs_layout = layout
    assert replay_x64(gu.load(name), stepper) is None


@pytest.mark.parametrize("layout,keep_old_a", [("soa", True), ("soa", False), ("auto", False)])
def test_batched_getstat_options_and_step_disp_against_the_reference_recording(layout, keep_old_a):
    """BatchedPhysicsEnv.step_disp (Creature.actdisp + PhysicsEnv.step) and BatchedPhysicsEnv.getstat (Creature.getstat
    with pk / vk / ak / mk / midform / conmid, through wg_getstat) on E copies of the recorded env: every copy reproduces
    the reference's recording (tests/golden/getstat_actdisp_custom3d.npz) bit for bit, in both observation layouts."""
    import json
    import os
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, Creature, DingPoint, Muscle, Point, Skeleton
    z = np.load(os.path.join(gu.GOLDEN_DIR, "getstat_actdisp_custom3d.npz"))
    spec = json.loads(str(z["spec"]))
    Point.clear()
    try:
        pts = [DingPoint(m, list(p)) if f else Point(m, list(p), [0, 0, 0]) for m, p, f in spec["points"]]
        cr = Creature(pts, [Muscle(pts[i], pts[j], **kw) for i, j, kw in spec["muscles"]],
                      [Skeleton(pts[i], pts[j], **kw) for i, j, kw in spec["skeletons"]])
        E, N = 300, len(pts)
        env = BatchedPhysicsEnv(cr, E, "cuda:0", in3d=True, auto_reset=None, state_layout=layout, keep_old_a=keep_old_a,
                                initial_reset=False)
        noise = torch.from_numpy(np.repeat(z["reset_noise"][: 3 * N, None], E, axis=1).copy()).cuda()
        env.reset(noise=noise, mode="jitter")
        variants = json.loads(str(z["variants"]))
        for t in range(z["disp"].shape[0]):
            disp = torch.from_numpy(np.repeat(z["disp"][t][None], E, axis=0).copy()).cuda()
            obs, rew, done, _ = env.step_disp(disp)
            assert gu.same(rew.cpu().numpy(), np.repeat(z["reward"][t], E)), t
            for v, kw in enumerate(variants):
                for lay in ("row", "feature"):
                    got = env.getstat(layout=lay, **kw).cpu().numpy()
                    got = got if lay == "row" else got.T
                    assert gu.same(got, np.repeat(z[f"stat{v}"][t][None], E, axis=0)), (t, v, lay)
            assert gu.same(env.getstat().cpu().numpy(), obs.cpu().numpy())          # the defaults == the step's observation
        with pytest.raises(ValueError):
            env.step_disp(torch.zeros(E + 1, 3, dtype=torch.bool, device="cuda:0"))
    finally:
        Point.clear()
