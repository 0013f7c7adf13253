"""CPU: the C-ABI library loads and exports what include/walker_gym_b200.h declares,
struct layouts agree between C and ctypes, argument validation works without a GPU,
and the host-side mirror of the reference interface behaves like the reference."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import golden_util as gu
import walker_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "walker_gym_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from walker_gym_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert {"wg_step", "wg_reset", "wg_stats_reduce", "wg_step_host", "wg_abi_version",
            "wg_last_error_string", "wg_obs_dim", "wg_kernel_variant", "wg_force_generic"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert lib.wg_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    from walker_gym_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "walker_gym_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(wg_topology), sizeof(wg_params), '
                    'sizeof(wg_buffers), offsetof(wg_topology, si), offsetof(wg_params, seed_lo), '
                    'offsetof(wg_buffers, obs), sizeof(wg_x64), sizeof(wg_pkg_system), sizeof(wg_pkg_params), '
                    'sizeof(wg_mlp_policy), offsetof(wg_buffers, mx64), offsetof(wg_pkg_system, srest), '
                    'offsetof(wg_topology, sstring), offsetof(wg_buffers, action_gen), sizeof(wg_action_gen), '
                    'offsetof(wg_action_gen, phase0)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    assert out == [C.sizeof(_lib.WgTopology), C.sizeof(_lib.WgParams), C.sizeof(_lib.WgBuffers),
                   _lib.WgTopology.si.offset, _lib.WgParams.seed_lo.offset, _lib.WgBuffers.obs.offset,
                   C.sizeof(_lib.WgX64), C.sizeof(_lib.WgPkgSystem), C.sizeof(_lib.WgPkgParams), C.sizeof(_lib.WgMlpPolicy),
                   _lib.WgBuffers.mx64.offset, _lib.WgPkgSystem.srest.offset,
                   _lib.WgTopology.sstring.offset, _lib.WgBuffers.action_gen.offset, C.sizeof(_lib.WgActionGen),
                   _lib.WgActionGen.phase0.offset]


def test_argument_validation_needs_no_gpu():
    from walker_gym_b200 import _lib, create_balance_creature, make_params
    from walker_gym_b200.topology import topology_from_creature
    lib = _lib.load()
    topo = topology_from_creature(create_balance_creature())
    prm = make_params(in3d=True)
    buf = _lib.WgBuffers()
    assert lib.wg_step(None, C.byref(prm), C.byref(buf), 4, None) == -1
    assert lib.wg_step(C.byref(topo), C.byref(prm), C.byref(buf), 4, None) == -1       # null state pointers
    assert b"pos/vel/mx/steps" in lib.wg_last_error_string()
    prm.k_sub = 0
    assert lib.wg_reset(C.byref(topo), C.byref(prm), C.byref(buf), 4, 1, None, None) == -1
    assert b"k_sub" in lib.wg_last_error_string()
    bad = topology_from_creature(create_balance_creature())
    bad.si[0] = 9
    prm.k_sub = 1
    assert lib.wg_step(C.byref(bad), C.byref(prm), C.byref(buf), 4, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, "wg_step")
    assert lib.wg_obs_dim(C.byref(topo), 1) == 38 and lib.wg_obs_dim(C.byref(topo), 0) == 26


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA instead of computing on the host."""
    import torch
    from walker_gym_b200 import BatchedPhysicsEnv, _lib
    with pytest.raises(_lib.WalkerGymError):
        BatchedPhysicsEnv("Balance-v0", 4, "cpu")
    if not torch.cuda.is_available():
        from walker_gym_b200 import make_env
        with pytest.raises(Exception):
            make_env("Balance-v0")
    for mod in ("walker_gym_b200", "walker_gym_b200.batched", "walker_gym_b200.env", "walker_gym_b200._lib"):
        src = open(sys.modules[mod].__file__).read()
        assert "walker_oracle" not in src and "ref_harness" not in src, f"{mod} must not touch oracle/"


def test_kernel_dispatch():
    from walker_gym_b200 import _lib, make_creature
    from walker_gym_b200.topology import topology_from_creature
    lib = _lib.load()
    want = {"balance_v0": 1, "balance": 1, "balance3": 0, "balance2": 0, "box_v0": 2, "box2": 2, "quad_balance": 3,
            "insect": 4, "box": 5, "test": 6, "intrian": 7, "hat": 8, "humanb": 9, "box4": 10, "leg2": 11, "leg": 12, "balance3": 0}
    for name, variant in want.items():
        topo = topology_from_creature(make_creature(name))
        assert lib.wg_kernel_variant(C.byref(topo)) == variant, name
        # balance2 (mass 0.1) and balance3 (DingPoint) share the Balance spring graph, whose packed kernel also
        # carries the general-mass path; their SoA route is the run-time-topology kernel (variant 0)
        packed = variant in (1, 2, 5, 6, 7, 8, 9, 10, 11, 12) or name in ("balance2", "balance3")
        assert (lib.wg_packed_available(C.byref(topo)) == 1) == packed, name
    old = lib.wg_force_generic(1)
    try:
        assert lib.wg_kernel_variant(C.byref(topology_from_creature(make_creature("box_v0")))) == 0
    finally:
        lib.wg_force_generic(old)


def spec_of(name):
    from walker_gym_b200 import BODIES
    b = BODIES[name]
    ding = set(b.get("ding", ()))
    return {"points": [(m, tuple(p), n in ding) for n, (m, p) in enumerate(b["points"])],
            "muscles": b["muscles"], "skeletons": b["skeletons"]}


@pytest.mark.parametrize("name", ["balance_v0", "box_v0", "test", "balance2", "balance3", "insect", "hat", "quad_balance"])
def test_topology_agrees_with_the_oracles_reading_of_the_morphology(name):
    from walker_gym_b200 import make_creature
    from walker_gym_b200.topology import topology_from_creature
    t, b = topology_from_creature(make_creature(name)), wo.make_body(spec_of(name))
    assert (t.n_mass, t.n_spring, t.n_muscle) == (b.n_mass, b.n_spring, b.n_muscle)
    for f in ("mass", "fixed"):
        assert list(getattr(t, f)[: t.n_mass]) == list(getattr(b, f)[: t.n_mass])
    assert list(t.tmpl_pos[: 3 * t.n_mass]) == list(b.tmpl_pos[: 3 * t.n_mass])
    for f in ("si", "sj", "sk", "sdamp", "srest", "mlo", "mhi"):
        assert list(getattr(t, f)[: t.n_spring]) == list(getattr(b, f)[: t.n_spring]), f


def test_custom_spec_topology_matches_oracle():
    from walker_gym_b200.topology import topology_from_spec
    g = gu.load("custom3d")
    t, b = topology_from_spec(g["spec"]), wo.make_body(g["spec"])
    for f in ("sk", "sdamp", "srest", "mlo", "mhi"):
        assert list(getattr(t, f)[: t.n_spring]) == list(getattr(b, f)[: t.n_spring]), f


def test_muscle_control_semantics():
    """Muscle.act / actdisp / regulation (gym/optimized_walker.py:27-43)."""
    from walker_gym_b200 import Muscle, Point
    Point.clear()
    p, q = Point(1, [0, 0, 0], [0, 0, 0]), Point(1, [3, 4, 0], [0, 0, 0])
    m = Muscle(p, q)
    assert m.x == np.float32(5) and m.originx == m.x and type(m.x) is np.float32
    m.act(np.float32(1.0)); assert m.x == np.float32(6)
    m.act(np.float32(100.0)); assert m.x == np.float32(7.5)          # clamp at originx * maxl
    m.act(np.float32(-100.0)); assert m.x == np.float32(5) * np.float32(0.1)
    m.actdisp(True); assert m.x == np.float32(0.5) + 2
    m.actdisp(False); assert m.x == np.float32(0.5)
    with pytest.raises(NotImplementedError):
        m.run()
    Point.clear()


def test_getstat_formats_like_the_reference():
    """Creature.getstat on mirrored state reproduces the reference's observation list."""
    from walker_gym_b200 import create_balance_creature, Point
    g = gu.load("balance3d_s0")
    Point.clear()
    c = create_balance_creature()
    for t in (1, 50):
        for n, p in enumerate(c.phys):
            p.pos[:], p.v[:], p.old_a = g["pos"][t][n], g["vel"][t][n], g["old_a"][t][n].copy()
        for i, m in enumerate(c.muscles):
            m.x = g["x"][t][i]
        assert gu.same(np.array(c.getstat(True), dtype=np.float32), g["obs"][t])
        assert len(c.getstat(False)) == 26 and len(c.getstat(True, conmid=True)) == 41
    Point.clear()


def test_make_env_rejects_unknown_ids():
    from walker_gym_b200 import make_env
    with pytest.raises(ValueError, match="Unknown environment ID"):
        make_env("Hopper-v0")


def test_state_pkl_roundtrip_and_reference_file(tmp_path):
    """state.pkl: read a file written by the reference, write one the reference's loader accepts."""
    import pickle
    from walker_gym_b200 import Point
    from walker_gym_b200.state_io import load_points, save_points
    pts, rp = load_points(os.path.join(gu.GOLDEN_DIR, "state_ref_box.pkl"))
    assert len(pts) == 4 and rp == {}
    assert gu.same(pts[0].v, np.float32([1.5, -2.0, 0.25])) and gu.same(pts[2].pos, np.float32([51.0, 99.5, -0.5]))
    assert pts[1].m == 1 and pts[1].pos.dtype == np.float32 and isinstance(pts[0], Point)
    out = tmp_path / "state.pkl"
    save_points(str(out), pts, {})
    import pickletools
    ops = [(op.name, arg) for op, arg, _ in pickletools.genops(out.read_bytes())]
    assert ("SHORT_BINUNICODE", "gym.engine") in ops and ("SHORT_BINUNICODE", "Point") in ops
    assert out.read_bytes()[:2] == b"\x80\x04"                                      # protocol 4
    pts2, _ = load_points(str(out))
    for a, b in zip(pts, pts2):
        assert gu.same(a.pos, b.pos) and gu.same(a.v, b.v) and a.m == b.m and a.color == b.color
    evil = tmp_path / "evil.pkl"
    evil.write_bytes(pickle.dumps({"points": [os.system], "r_points": {}}, protocol=4))
    with pytest.raises(pickle.UnpicklingError):
        load_points(str(evil))
    Point.clear(); Point.backup(str(out)); assert len(Point.points) == 4; Point.clear()


@pytest.mark.reference
def test_reference_can_load_our_snapshot(tmp_path):
    """When the reference checkout is present: its own Point.backup reads what we write."""
    import ref_harness as rh
    if not rh.available():
        pytest.skip("reference checkout not present")
    from walker_gym_b200 import Point, create_box_creature
    from walker_gym_b200.state_io import save_points
    Point.clear()
    c = create_box_creature()
    c.phys[1].v[:] = [3, 2, 1]
    out = tmp_path / "state.pkl"
    save_points(str(out), c.phys, {}, module="optimized_engine")
    engine, _, _ = rh.load()
    engine.Point.clear()
    engine.Point.backup(str(out))
    assert len(engine.Point.points) == 4
    assert gu.same(engine.Point.points[1].v, np.float32([3, 2, 1]))
    assert gu.same(engine.Point.positions, np.stack([p.pos for p in c.phys]))
    engine.Point.clear(); Point.clear()


@pytest.mark.reference
def test_oracle_against_live_reference_random_bodies():
    """When the reference checkout is present: fresh random rollouts, oracle == reference."""
    import ref_harness as rh
    if not rh.available():
        pytest.skip("reference checkout not present")
    import warnings
    from test_oracle_golden import replay_trajectory
    rng = np.random.default_rng(77)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for trial in range(4):
            N = int(rng.integers(3, 8))
            pts = [(float(rng.choice([1, 2, 3, 5, 0.5, 0.1, 7.25])), tuple(float(v) for v in rng.uniform(-80, 120, 3)), False)
                   for _ in range(N)]
            pairs = [(i, j) for i in range(N) for j in range(i + 1, N)]
            rng.shuffle(pairs)
            M = int(rng.integers(1, 4))
            spec = {"points": pts, "muscles": [(i, j, {}) for i, j in pairs[:M]],
                    "skeletons": [(i, j, {"k": float(rng.choice([500, 1000, -800]))}) for i, j in pairs[M:M + N]]}
            kw = dict(in3d=bool(trial % 2), g=float(rng.choice([100, 9.8])), dampk=float(rng.choice([0, 0.3])))
            acts = rng.uniform(-1, 1, (60, M)).astype(np.float32)
            noise = (rng.standard_normal(256) * 0.1).astype(np.float32)
            out = rh.rollout(spec, acts, env_kwargs=kw, noise=noise)
            g = dict(out, spec=spec, env_kwargs=kw, actions=acts, k_sub=None, max_steps=None, reset_on_done=None)
            g["reset_noise"] = out["reset_noise"].astype(np.float32)
            for k in ("obs", "reward", "energy", "centroid", "x"):
                g[k] = np.asarray(out[k], np.float32)
            assert replay_trajectory(g, wo) is None, f"trial {trial}"


@pytest.mark.reference
def test_config1_walker_py_body_1000_steps_oracle_vs_live_reference():
    """BASELINE config 1 on the CPU: a gym/walker.py body, 1000 random-action steps, oracle == reference."""
    import ref_harness as rh
    if not rh.available():
        pytest.skip("reference checkout not present")
    import warnings
    from test_oracle_golden import replay_trajectory
    spec = spec_of("humanb")
    rng = np.random.default_rng(123)
    acts = rng.uniform(-1, 1, (1000, 4)).astype(np.float32)
    noise = (rng.standard_normal(64) * 0.1).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = rh.rollout(spec, acts, env_kwargs=dict(in3d=True), noise=noise)
    g = dict(out, spec=spec, env_kwargs=dict(in3d=True), actions=acts, k_sub=None, max_steps=None, reset_on_done=None)
    g["reset_noise"] = out["reset_noise"].astype(np.float32)
    for k in ("obs", "reward", "energy", "centroid", "x"):
        g[k] = np.asarray(out[k], np.float32)
    assert replay_trajectory(g, wo) is None
    assert bool(out["done"][-1]) and not out["done"][:-1].any()


def test_bind_to_device_picks_a_subset_of_the_allowed_cpus():
    import os
    from walker_gym_b200.host import bind_to_device
    before = os.sched_getaffinity(0)
    try:
        info = bind_to_device(0, 2)
        assert set(info["cpus"]) <= before and len(info["cpus"]) >= 1
        assert os.sched_getaffinity(0) == set(info["cpus"])
        other = None
        os.sched_setaffinity(0, before)
        other = bind_to_device(1, 2)
        if len(before) >= 2:
            assert not (set(info["cpus"]) & set(other["cpus"]))      # two ranks never share a core
    finally:
        os.sched_setaffinity(0, before)


def test_reference_bytecode_recipe_and_harness(tmp_path):
    """oracle/make_ref.py byte-compiles the unmodified reference modules; the harness imports them from bytecode and
    steps the same trajectory as from the sources (only where the read-only checkout exists)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import make_ref
    import ref_harness as rh
    if not rh.available(rh.DEFAULT_REF):
        pytest.skip("reference checkout not present")
    assert make_ref.make(quiet=True) and rh.available(rh.COMPILED_REF)
    acts = np.random.default_rng(0).uniform(-1, 1, (30, 2)).astype(np.float32)
    a = rh.rollout("Balance-v0", acts, env_kwargs=dict(in3d=True), seed=3, ref_root=rh.DEFAULT_REF)
    b = rh.rollout("Balance-v0", acts, env_kwargs=dict(in3d=True), seed=3, ref_root=rh.COMPILED_REF)
    for k in ("pos", "vel", "obs", "reward", "done"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    listing = os.listdir(os.path.join(rh.COMPILED_REF, "gym"))
    assert not [f for f in listing if f.endswith((".py", ".pyc"))], "bytecode only: no reference source text in the repo tree"
