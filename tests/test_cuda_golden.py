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
