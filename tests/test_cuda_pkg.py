"""GPU: the package lineage's Environment.update_physics (gym/optimized_walker/env.py:135-184) through
wg_pkg_update_physics -- bit-exact against the reference's recorded trajectories and against the oracle."""
import os

import numpy as np
import pytest

import golden_util as gu
import walker_oracle as wo
from test_oracle_golden import replay_l2
from test_pkg_host import BODY_NAMES, body_record, build, replay_body, system_of

pytestmark = pytest.mark.gpu


@pytest.fixture()
def cs():
    import cuda_stepper
    cuda_stepper.force_generic = False
    yield cuda_stepper
    cuda_stepper.force_generic = False


@pytest.mark.parametrize("chunk", [1, 7, 1000])
@pytest.mark.parametrize("name", gu.l2_names())
def test_cuda_matches_reference_package_physics(cs, name, chunk):
    """DingPoints, string springs, explicit rest lengths, ground bounce, non-default parameters; 1, 7 or all
    steps per launch."""
    assert replay_l2(gu.load_l2(name), cs, chunk) is None


@pytest.mark.parametrize("parts", [2, 4, 8])
@pytest.mark.parametrize("name", ["box", "humanb", "insect", "leg2"])
def test_cuda_partitioned_kernel_matches_reference_bodies(cs, part_knob, name, parts):
    """The reference-built trajectories through the point-partitioned kernel (static specialisations off)."""
    from walker_gym_b200 import _lib
    part_knob(parts)
    rec = body_record(name)
    if name == "leg2":                       # has a register-resident kernel: take it out of the way
        rec = dict(rec)
        rec["si"], rec["sj"] = rec["sj"].copy(), rec["si"].copy()      # same springs, endpoints swapped: no static match
        system = system_of(rec)
        st_o, st_c = wo.l2_init_state(system, 5), wo.l2_init_state(system, 5)
        wo.l2_step(wo.make_l2_system(system), wo.make_l2_params(), st_o, 50)
        cs.l2_step(cs.make_l2_system(system), cs.make_l2_params(), st_c, 50)
        assert gu.same(st_c["pos"], st_o["pos"]) and gu.same(st_c["old_a"], st_o["old_a"])
        return
    assert replay_body(rec, cs, 6) is None


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("name", BODY_NAMES)
def test_cuda_matches_reference_bodies(cs, name, generic):
    """Every body the reference ships, through its register-resident specialisation (box, leg2, balance1-3, test)
    and through the run-time-topology kernel."""
    import ctypes as C
    from walker_gym_b200 import _lib
    cs.force_generic = generic
    rec = body_record(name)
    lib = _lib.load()
    old = lib.wg_force_generic(1 if generic else 0)
    variant = lib.wg_pkg_kernel_variant(C.byref(cs.make_l2_system(system_of(rec))))
    lib.wg_force_generic(old)
    assert (variant > 0) == (not generic and name in ("leg2", "balance1", "balance2", "balance3", "test"))
    assert replay_body(rec, cs, 6) is None


def test_cuda_static_and_generic_kernels_agree_on_perturbed_bodies(cs):
    """The specialised leg2 kernel vs the generic kernel vs the oracle on 5000 perturbed envs, strings and a DingPoint."""
    rng = np.random.default_rng(11)
    system = system_of(body_record("leg2"))
    system["points"][3] = system["points"][3][:3] + (True,)
    system["points"][5] = (2.5,) + system["points"][5][1:]
    system["springs"][4] = system["springs"][4][:4] + (True,)
    E = 5000
    st = wo.l2_init_state(system, E)
    st["pos"] += rng.normal(0, 1.0, st["pos"].shape).astype(np.float32)
    st["vel"] += rng.normal(0, 3.0, st["vel"].shape).astype(np.float32)
    kw = dict(ground_level=-17, gravity=(0.0, -50.0, 1.0))
    res = []
    for generic in (False, True):
        cs.force_generic = generic
        s2 = {k: v.copy() for k, v in st.items()}
        cs.l2_step(cs.make_l2_system(system), cs.make_l2_params(**kw), s2, 60)
        res.append(s2)
    wo.l2_step(wo.make_l2_system(system), wo.make_l2_params(**kw), st, 60)
    for k in ("pos", "vel", "old_a"):
        assert gu.same(res[0][k], st[k]) and gu.same(res[1][k], st[k]), k


def random_system(rng, P, S, dings=True):
    pts = []
    for n in range(P):
        m = [1.0, 2.0, 0.5, 3.0, float(rng.uniform(0.3, 7.0)), 1.5][int(rng.integers(0, 6))]
        pts.append((m, tuple(rng.uniform(-40, 40, 3).astype(np.float32).tolist()),
                    tuple(rng.uniform(-5, 5, 3).astype(np.float32).tolist()), bool(dings and rng.random() < 0.15)))
    if all(p[3] for p in pts):
        pts[0] = pts[0][:3] + (False,)
    sps = []
    for _ in range(S):
        i, j = rng.choice(P, 2, replace=False) if P > 1 else (0, 0)
        x = None if rng.random() < 0.6 else float(rng.uniform(5, 60))
        sps.append((int(i), int(j), x, float(rng.choice([50, 100, 250.5, 1000])), bool(rng.random() < 0.3)))
    return {"points": pts, "springs": sps}


@pytest.fixture()
def part_knob():
    from walker_gym_b200 import _lib
    lib = _lib.load()
    saved = lib.wg_set_tuning(_lib.TUNE_PART, -1)

    def set_parts(parts):
        lib.wg_set_tuning(_lib.TUNE_PART, parts)
    yield set_parts
    lib.wg_set_tuning(_lib.TUNE_PART, saved)


@pytest.mark.parametrize("parts", [-1, 0, 2, 4, 8])
@pytest.mark.parametrize("seed,P,S", [(0, 1, 0), (1, 2, 1), (2, 5, 9), (3, 12, 30), (4, 32, 96), (5, 21, 20)])
def test_cuda_matches_oracle_on_random_systems(cs, part_knob, seed, P, S, parts):
    """Random topologies up to the ABI limits, perturbed per env, 3 x 40 steps, ragged env counts; one thread per env
    (parts 0) and 2 / 4 / 8 lanes per env (crossing springs evaluated by both owners): the bits must not change."""
    part_knob(parts)
    rng = np.random.default_rng(seed)
    system = random_system(rng, P, S)
    kw = dict(ground_level=-30, gravity=(0.3, -25.0, -0.2), damping=0.97, air_resistance=0.05, time_step=0.01)
    E = 777
    st_o = wo.l2_init_state(system, E)
    st_o["pos"] += rng.normal(0, 2.0, st_o["pos"].shape).astype(np.float32)
    st_o["vel"] += rng.normal(0, 1.0, st_o["vel"].shape).astype(np.float32)
    st_c = {k: v.copy() for k, v in st_o.items()}
    so, po = wo.make_l2_system(system), wo.make_l2_params(**kw)
    sc, pc = cs.make_l2_system(system), cs.make_l2_params(**kw)
    for _ in range(3):
        wo.l2_step(so, po, st_o, 40)
        cs.l2_step(sc, pc, st_c, 40)
        for k in ("pos", "vel", "old_a"):
            assert gu.same(st_c[k], st_o[k]), k


def test_cuda_package_physics_nonfinite_and_coincident_points(cs):
    """inf / NaN lanes and zero-length springs take the out-of-line IEEE paths: same bits as the oracle."""
    system = {"points": [(1.0, (0, 0, 0), (0, 0, 0), False), (2.0, (0, 0, 0), (0, 0, 0), False),
                         (3.0, (1e-20, 0, 0), (0, 0, 0), False), (1.0, (5, 5, 5), (1e30, 0, 0), False)],
              "springs": [(0, 1, 1.0, 100.0, False), (1, 2, None, 100.0, False), (2, 3, None, 1e30, True), (0, 3, 2.0, 50.0, False)]}
    E = 64
    st_o = wo.l2_init_state(system, E)
    st_o["pos"][0, 1] = np.inf
    st_o["vel"][4, 2] = np.nan
    st_o["pos"][9, 3] = 3e38
    st_c = {k: v.copy() for k, v in st_o.items()}
    kw = dict(time_step=0.5)
    for _ in range(4):
        wo.l2_step(wo.make_l2_system(system), wo.make_l2_params(**kw), st_o, 5)
        cs.l2_step(cs.make_l2_system(system), cs.make_l2_params(**kw), st_c, 5)
        for k in ("pos", "vel", "old_a"):
            assert gu.same(st_c[k], st_o[k]), k
    assert not np.isfinite(st_c["pos"]).all()


def test_cuda_package_physics_large_batch_properties(cs):
    """2^20 envs: duplicated envs stay identical, a slice agrees with the oracle, 1 x 20 steps == 20 x 1 step."""
    import ctypes as C
    import torch
    from walker_gym_b200 import _lib
    system = system_of(body_record("box"))
    E, H = 1 << 20, 1 << 19
    kw = dict(ground_level=-8)
    sysm, prm = cs.make_l2_system(system), cs.make_l2_params(**kw)
    base = wo.l2_init_state(system, 1)
    R = base["pos"].shape[0]
    v_half = np.random.default_rng(3).normal(0, 1, (R, H)).astype(np.float32)
    pos = torch.from_numpy(base["pos"]).to(cs.DEV).repeat(1, E).contiguous()
    vel = torch.from_numpy(np.concatenate([v_half, v_half], 1)).to(cs.DEV)      # second half duplicates the first
    pos2, vel2 = pos.clone(), vel.clone()
    lib, stream = _lib.load(), cs._stream()
    assert lib.wg_pkg_update_physics(C.byref(sysm), C.byref(prm), pos.data_ptr(), vel.data_ptr(), None, E, 20, stream) == 0
    for _ in range(20):
        assert lib.wg_pkg_update_physics(C.byref(sysm), C.byref(prm), pos2.data_ptr(), vel2.data_ptr(), None, E, 1, stream) == 0
    torch.cuda.synchronize()
    assert torch.equal(pos.view(torch.int32), pos2.view(torch.int32)) and torch.equal(vel.view(torch.int32), vel2.view(torch.int32))
    assert torch.equal(pos[:, :H].view(torch.int32), pos[:, H:].view(torch.int32))
    sl = slice(1000, 1512)
    st = dict(pos=np.ascontiguousarray(np.repeat(base["pos"], 512, 1)), vel=np.ascontiguousarray(v_half[:, sl]),
              old_a=np.zeros((R, 512), np.float32))
    wo.l2_step(wo.make_l2_system(system), wo.make_l2_params(**kw), st, 20)
    assert gu.same(pos[:, sl].cpu().numpy(), st["pos"]) and gu.same(vel[:, sl].cpu().numpy(), st["vel"])
    assert (pos[1::3].min() >= -8).item()                    # the ground clamp held everywhere


@pytest.mark.parametrize("name", ["leg2", "balance3", "insect"])
def test_environment_mirror_single_env(name):
    """The drop-in Environment (E = 1): Point objects are refreshed after every update, like the reference's."""
    rec = body_record(name)
    env, creature = build(name)
    pts = env._order
    for t in range(1, 31):
        creature.act(env.time_step)
        env.update_physics()
        assert gu.same(np.array([p.pos for p in pts], np.float32), rec["pos"][t]), t
        assert gu.same(np.array([p.v for p in pts], np.float32), rec["vel"][t]), t
    env.run(steps=90)                                        # 90 more updates in one launch
    assert gu.same(np.array([p.pos for p in pts], np.float32), rec["pos"][120])
    assert env.get_statistics()["frame_count"] == 120
    assert np.isfinite(creature.evaluate_fitness())


def test_environment_mirror_user_edits_and_batched():
    import torch
    import walker_gym_b200.optimized_walker as ow
    rec = body_record("humanb")
    # user edits between updates are picked up (E = 1: the Point objects are the state)
    env, _ = build("humanb")
    env.update_physics(10)
    env.points[0].v[0] += 3.0
    env.update_physics(5)
    system = system_of(rec)
    st = wo.l2_init_state(system, 1)
    so, po = wo.make_l2_system(system), wo.make_l2_params()
    wo.l2_step(so, po, st, 10)
    st["vel"][0, 0] += np.float32(3.0)
    wo.l2_step(so, po, st, 5)
    assert gu.same(np.array([p.pos for p in env._order], np.float32), st["pos"].reshape(-1, 3))
    # batched: every env steps; perturb one env on the device and compare both with the oracle
    ow.Point.clear()
    benv = ow.Environment(num_envs=300)
    c = ow.humanb(benv)
    benv.vel[3, 17] = 2.5
    benv.update_physics(25)
    st = wo.l2_init_state(system, 300)
    st["vel"][3, 17] = 2.5
    wo.l2_step(so, po, st, 25)
    assert gu.same(benv.pos.cpu().numpy(), st["pos"]) and gu.same(benv.old_a.cpu().numpy(), st["old_a"])
    fit = c.evaluate_fitness_batched()
    assert fit.shape == (300,) and torch.isfinite(fit).all() and fit[17] != fit[0]
    with pytest.raises(RuntimeError):
        benv.add_point(1, (0, 0, 0))


def test_environment_continues_from_reference_env_state():
    """Load the reference-written env_state.pkl (25 updates in) and continue 40 updates on the GPU: the
    result is what the reference itself computed after those 40 updates."""
    import walker_gym_b200.optimized_walker as ow
    env = ow.Environment()
    env.load_state(os.path.join(gu.GOLDEN_DIR, "env_state_ref.pkl"))
    env.update_physics(40)
    ref = np.load(os.path.join(gu.GOLDEN_DIR, "env_state_ref_after40.npz"))
    pts = [env.ding_points[0], env.points[0], env.points[1]]      # the reference's (a, b, c)
    assert gu.same(np.array([p.pos for p in pts], np.float32), ref["pos"])
    assert gu.same(np.array([p.v for p in pts], np.float32), ref["vel"])
