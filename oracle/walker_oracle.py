"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the C oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package (``walker_gym_b200``) never does.

A *body spec* is a plain dict (no product classes involved)::

    {"points":    [(m, (x, y, z), fixed), ...],
     "muscles":   [(i, j, {"k":..., "dampk":..., "minl":..., "maxl":..., "x":...}), ...],
     "skeletons": [(i, j, {"k":..., "dampk":..., "x":...}), ...]}

with the reference constructors' defaults (``gym/optimized_walker.py:8-9,70``:
k=1000, dampk=20, minl=0.1, maxl=1.5, rest length = initial distance).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwalker_oracle.so")
MAX_MASS, MAX_SPRING = 32, 96


class Body(C.Structure):
    _fields_ = [
        ("n_mass", C.c_int32), ("n_spring", C.c_int32), ("n_muscle", C.c_int32),
        ("mass", C.c_double * MAX_MASS),
        ("fixed", C.c_uint8 * MAX_MASS),
        ("tmpl_pos", C.c_float * (MAX_MASS * 3)),
        ("si", C.c_int32 * MAX_SPRING), ("sj", C.c_int32 * MAX_SPRING),
        ("sk", C.c_float * MAX_SPRING), ("sdamp", C.c_float * MAX_SPRING),
        ("srest", C.c_float * MAX_SPRING),
        ("mlo", C.c_float * MAX_SPRING), ("mhi", C.c_float * MAX_SPRING),
        ("sstring", C.c_uint8 * MAX_SPRING),
    ]


class Params(C.Structure):
    _fields_ = [
        ("g", C.c_double),
        ("dampk", C.c_float), ("ground", C.c_float), ("fall_thresh", C.c_float),
        ("ground_k", C.c_float), ("ground_damp", C.c_float), ("friction", C.c_float),
        ("dt", C.c_float), ("dt2", C.c_float), ("sigma", C.c_float),
        ("in3d", C.c_int32), ("max_steps", C.c_int32), ("k_sub", C.c_int32),
        ("auto_reset", C.c_int32), ("integrator", C.c_int32),
        ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32),
        ("step_index", C.c_uint32), ("env_offset", C.c_uint32),
    ]


class X64(C.Structure):
    """``wgo_x64``: the double-typed objects the reference keeps once a float64 action made Muscle.x an np.float64."""
    _fields_ = [("sk_d", C.c_double * MAX_SPRING), ("x0_d", C.c_double * MAX_SPRING),
                ("mlo_d", C.c_double * MAX_SPRING), ("mhi_d", C.c_double * MAX_SPRING)]


_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "walker_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-B", "libwalker_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        assert _lib.wgo_sizeof_body() == C.sizeof(Body)
        assert _lib.wgo_sizeof_params() == C.sizeof(Params)
        _lib.wgo_step.restype = C.c_int
        _lib.wgo_reset.restype = C.c_int
    return _lib


def set_threads(n: int) -> int:
    """Set the OpenMP team size of the oracle (torchrun exports OMP_NUM_THREADS=1); returns the size in effect."""
    lib().wgo_set_num_threads(C.c_int(int(n)))
    return int(lib().wgo_get_max_threads())


def make_body(spec) -> Body:
    """Evaluate the morphology exactly as the reference constructors would."""
    b = Body()
    pts = spec["points"]
    mus, sks = spec["muscles"], spec["skeletons"]
    b.n_mass, b.n_muscle, b.n_spring = len(pts), len(mus), len(mus) + len(sks)
    assert b.n_mass <= MAX_MASS and b.n_spring <= MAX_SPRING
    P = [np.array(p[1], dtype=np.float32) for p in pts]   # Point.pos  (optimized_engine.py:55)
    for n, (m, pos, fixed) in enumerate(pts):
        b.mass[n] = float(m)
        b.fixed[n] = 1 if fixed else 0
        for c in range(3):
            b.tmpl_pos[n * 3 + c] = P[n][c]
    for s, (i, j, kw) in enumerate(list(mus) + list(sks)):
        b.si[s], b.sj[s] = i, j
        b.sk[s] = np.float32(kw.get("k", 1000))
        b.sdamp[s] = np.float32(kw.get("dampk", 20))
        x = kw.get("x")
        if x is None:
            x = np.linalg.norm(P[i] - P[j])                # Muscle.distant (optimized_walker.py:23-25)
        b.srest[s] = np.float32(x)
        b.sstring[s] = 1 if kw.get("string", False) else 0
        if s < len(mus):
            b.mlo[s] = np.float32(x * kw.get("minl", 0.1))  # regulation (:27-30), python semantics
            b.mhi[s] = np.float32(x * kw.get("maxl", 1.5))
    return b


def make_params(in3d=False, g=100, dampk=0, ground_high=0, ground_k=1000, ground_damp=100,
                friction=100, rand_sigma=0.1, time_step=0.01, max_steps=1000, k_sub=1,
                auto_reset=0, seed=0, step_index=0, env_offset=0, integrator=0) -> Params:
    p = Params()
    p.g = float(g)
    p.dampk = np.float32(dampk)
    p.ground = np.float32(ground_high)
    p.fall_thresh = np.float32(ground_high - 50)
    p.ground_k = np.float32(ground_k)
    p.ground_damp = np.float32(ground_damp)
    p.friction = np.float32(friction)
    p.dt = np.float32(time_step)
    p.dt2 = np.float32(time_step ** 2)
    p.integrator = 1 if integrator in (1, "run2") else 0
    p.sigma = np.float32(rand_sigma)
    p.in3d, p.max_steps, p.k_sub = int(bool(in3d)), int(max_steps), int(k_sub)
    p.auto_reset = int(auto_reset)
    p.seed_lo, p.seed_hi = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    p.step_index, p.env_offset = step_index, env_offset
    return p


def obs_dim(body: Body, in3d) -> int:
    return 3 * (3 if in3d else 2) * body.n_mass + body.n_muscle


def _ptr(a, ctype):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    return a.ctypes.data_as(C.POINTER(ctype))


def init_state(body: Body, E: int):
    """Template state for E envs in the SoA layout [(n*3+c)][E]."""
    N, M = body.n_mass, body.n_muscle
    tp = np.array(body.tmpl_pos[: N * 3], dtype=np.float32)
    st = dict(
        pos=np.repeat(tp[:, None], E, axis=1).copy(),
        vel=np.zeros((N * 3, E), np.float32),
        old_a=np.zeros((N * 3, E), np.float32),
        mx=np.repeat(np.array(body.srest[:M], dtype=np.float32)[:, None], E, axis=1).copy().reshape(M, E),
        steps=np.zeros(E, np.int32),
    )
    return st


def step(body: Body, prm: Params, st: dict, action: np.ndarray, *, want_info=True,
         ep_ret=None, fin_stats=None, noise=None):
    """One env-step, in place on ``st``.  Returns dict(obs, reward, done, ...)."""
    E = st["pos"].shape[1]
    D = obs_dim(body, prm.in3d)
    action = np.ascontiguousarray(action, dtype=np.float32).reshape(E, -1)
    out = dict(
        obs=np.empty((E, D), np.float32), reward=np.empty(E, np.float32), done=np.empty(E, np.uint8),
        contact_pre=np.empty(E, np.uint32), contact_post=np.empty(E, np.uint32),
    )
    if want_info:
        out["energy"] = np.empty(E, np.float32)
        out["centroid"] = np.empty((3, E), np.float32)
    rc = lib().wgo_step(
        C.byref(body), C.byref(prm), C.c_int64(E),
        _ptr(st["pos"], C.c_float), _ptr(st["vel"], C.c_float), _ptr(st.get("old_a"), C.c_float),
        _ptr(st["mx"], C.c_float), _ptr(st["steps"], C.c_int32),
        _ptr(action, C.c_float), C.c_int32(action.shape[1]),
        _ptr(out["obs"], C.c_float), _ptr(out["reward"], C.c_float), _ptr(out["done"], C.c_uint8),
        _ptr(out["contact_pre"], C.c_uint32), _ptr(out["contact_post"], C.c_uint32),
        _ptr(out.get("energy"), C.c_float), _ptr(out.get("centroid"), C.c_float),
        _ptr(ep_ret, C.c_float), _ptr(fin_stats, C.c_float), _ptr(noise, C.c_float))
    if rc != 0:
        raise RuntimeError(f"wgo_step failed: {rc}")
    return out


def reset(body: Body, prm: Params, st: dict, *, mode=1, mask=None, noise=None):
    """PhysicsEnv.reset for the masked envs (mode 1 = jitter, 2 = template). Returns obs."""
    E = st["pos"].shape[1]
    obs = np.zeros((E, obs_dim(body, prm.in3d)), np.float32)
    rc = lib().wgo_reset(
        C.byref(body), C.byref(prm), C.c_int64(E), C.c_int(mode),
        _ptr(st["pos"], C.c_float), _ptr(st["vel"], C.c_float), _ptr(st.get("old_a"), C.c_float),
        _ptr(st["mx"], C.c_float), _ptr(st["steps"], C.c_int32),
        _ptr(obs, C.c_float), _ptr(mask, C.c_uint8), _ptr(noise, C.c_float))
    if rc != 0:
        raise RuntimeError(f"wgo_reset failed: {rc}")
    return obs


def gen_actions(gen: dict, steps: np.ndarray, M: int) -> np.ndarray:
    """The in-kernel action sources of wg_step_multi for envs whose step counters (before the step) are ``steps``.
    gen = {"mode": 1, "table": [T0][M], "hold": h}  or  {"mode": 2, "amp": [M], "phase0": [M] uint32, "dphase": [M] uint32}."""
    E = int(steps.shape[0])
    out = np.empty((E, M), np.float32)
    table = np.zeros((32, 16), np.float32)
    amp, p0, dp = np.zeros(16, np.float32), np.zeros(16, np.uint32), np.zeros(16, np.uint32)
    n_rows, hold = 1, 1
    if gen["mode"] == 1:
        t = np.asarray(gen["table"], np.float32)
        n_rows, hold = t.shape[0], int(gen["hold"])
        table[:n_rows, :t.shape[1]] = t
    else:
        amp[:M], p0[:M], dp[:M] = gen["amp"], gen["phase0"], gen["dphase"]
    lib().wgo_gen_actions(C.c_int(gen["mode"]), C.c_int(n_rows), C.c_int(hold), _ptr(table, C.c_float), _ptr(amp, C.c_float),
                          _ptr(p0, C.c_uint32), _ptr(dp, C.c_uint32), _ptr(np.ascontiguousarray(steps, np.int32), C.c_int32),
                          C.c_int64(E), C.c_int(M), _ptr(out, C.c_float))
    return out


def getstat(body: Body, st: dict, old_a: np.ndarray, in3d=True, pk=1, vk=1, ak=1, mk=1, midform=True, conmid=False) -> np.ndarray:
    """Creature.getstat (gym/optimized_walker.py:129-162) for every env of an SoA state: [E, 3*d*N + M (+3)] float32.
    NumPy float32 arithmetic in the reference's order: sequential `mid += pos`, `mid /= N`, `(pos - mid) * pk`."""
    N, M = body.n_mass, body.n_muscle
    E = st["pos"].shape[1]
    d = 3 if in3d else 2
    pos, vel = st["pos"].reshape(N, 3, E), st["vel"].reshape(N, 3, E)
    oa = np.asarray(old_a, np.float32).reshape(N, 3, E)
    mid = np.zeros((3, E), np.float32)
    if midform:
        for n in range(N):
            mid += pos[n]
        mid /= N
    cols = []
    for n in range(N):
        cols += [(((pos[n, c] - mid[c]) if midform else pos[n, c]) * np.float32(pk)) for c in range(d)]
        cols += [vel[n, c] * np.float32(vk) for c in range(d)]
        cols += [oa[n, c] * np.float32(ak) for c in range(d)]
    if conmid:
        cols += [mid[c] for c in range(3)]
    cols += [st["mx"][m] * np.float32(mk) for m in range(M)]
    return np.stack(cols, axis=1).astype(np.float32)


def normal3(seed: int, env: int, step: int, mass: int) -> np.ndarray:
    out = (C.c_float * 3)()
    lib().wgo_normal3(C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32((seed >> 32) & 0xFFFFFFFF),
                      C.c_uint32(env), C.c_uint32(step), C.c_uint32(mass), out)
    return np.array(out[:], dtype=np.float32)


# The two morphologies of gym/optimized_walker.py:176-224 as specs (data only).
BALANCE = {
    "points": [(5, (-50, 100, 0), False), (5, (50, 100, 0), False), (1, (0, 0, 0), False), (3, (0, 100, 0), False)],
    "muscles": [(0, 2, {}), (1, 2, {})],
    "skeletons": [(0, 1, {}), (0, 3, {}), (1, 3, {})],
}
BOX = {
    "points": [(1, (-50, 0, 0), False), (1, (-50, 100, 0), False), (1, (50, 100, 0), False), (1, (50, 0, 0), False)],
    "muscles": [(0, 1, {}), (0, 2, {}), (3, 1, {}), (3, 2, {})],
    "skeletons": [(1, 2, {})],
}


def make_x64(spec) -> X64:
    """The python / NumPy objects the reference's Muscle keeps, as doubles: float(k); originx (np.float32 from
    distant(), or the python float the user passed); originx * minl and originx * maxl as regulation() forms them."""
    xb = X64()
    P = [np.array(p[1], dtype=np.float32) for p in spec["points"]]
    for s, (i, j, kw) in enumerate(list(spec["muscles"]) + list(spec["skeletons"])):
        xb.sk_d[s] = float(kw.get("k", 1000))
        x = kw.get("x")
        if x is None:
            x = np.linalg.norm(P[i] - P[j])                # np.float32
        xb.x0_d[s] = float(x)
        if s < len(spec["muscles"]):
            xb.mlo_d[s] = float(x * kw.get("minl", 0.1))    # np.float32 * python float -> np.float32; python * python -> double
            xb.mhi_d[s] = float(x * kw.get("maxl", 1.5))
    return xb


def init_x64(body: Body, xb: X64, E: int):
    """mx64 [M, E] = originx, mx_weak [M, E] = 1 (a fresh Muscle holds the constructor's object)."""
    M = body.n_muscle
    x0 = np.array([xb.x0_d[m] for m in range(M)], np.float64)
    return np.repeat(x0[:, None], E, 1).copy().reshape(M, E), np.ones((M, E), np.uint8)


def step_x64(body: Body, xb: X64, prm: Params, st: dict, action64, *, want_info=True, noise=None):
    """``PhysicsEnv.step`` with float64 actions [E, A]; ``st`` additionally holds ``mx64`` and ``mx_weak``."""
    l = lib()
    assert l.wgo_sizeof_x64() == C.sizeof(X64)
    E = st["pos"].shape[1]
    D = obs_dim(body, prm.in3d)
    action64 = np.ascontiguousarray(action64, dtype=np.float64).reshape(E, -1)
    out = dict(obs=np.zeros((E, D), np.float32), reward=np.zeros(E, np.float32), done=np.zeros(E, np.uint8),
               contact_pre=np.zeros(E, np.uint32), contact_post=np.zeros(E, np.uint32))
    if want_info:
        out["energy"] = np.zeros(E, np.float32)
        out["centroid"] = np.zeros((3, E), np.float32)
    rc = l.wgo_step_x64(C.byref(body), C.byref(xb), C.byref(prm), C.c_int64(E),
                        _ptr(st["pos"], C.c_float), _ptr(st["vel"], C.c_float), _ptr(st.get("old_a"), C.c_float),
                        _ptr(st["mx"], C.c_float), _ptr(st["mx64"], C.c_double), _ptr(st["mx_weak"], C.c_uint8),
                        _ptr(st["steps"], C.c_int32), _ptr(action64, C.c_double), C.c_int32(action64.shape[1]),
                        _ptr(out["obs"], C.c_float), _ptr(out["reward"], C.c_float), _ptr(out["done"], C.c_uint8),
                        _ptr(out["contact_pre"], C.c_uint32), _ptr(out["contact_post"], C.c_uint32),
                        _ptr(out.get("energy"), C.c_float), _ptr(out.get("centroid"), C.c_float), None, None,
                        _ptr(noise, C.c_float))
    if rc != 0:
        raise RuntimeError(f"wgo_step_x64 failed: {rc}")
    return out


# ---- L2 (package lineage): Environment.update_physics of gym/optimized_walker/env.py ------------------
class L2System(C.Structure):
    _fields_ = [("n_point", C.c_int32), ("n_spring", C.c_int32),
                ("mass", C.c_double * MAX_MASS), ("ding", C.c_uint8 * MAX_MASS),
                ("si", C.c_int32 * MAX_SPRING), ("sj", C.c_int32 * MAX_SPRING),
                ("sx", C.c_float * MAX_SPRING), ("sk", C.c_float * MAX_SPRING), ("sstring", C.c_uint8 * MAX_SPRING)]


class L2Params(C.Structure):
    _fields_ = [("gravity", C.c_float * 3), ("damping", C.c_float), ("drag_c", C.c_float),
                ("ground_level", C.c_float), ("restitution", C.c_float), ("friction", C.c_float), ("dt", C.c_float),
                ("min_dist", C.c_float), ("ground", C.c_int32)]


def make_l2_system(system) -> L2System:
    """system: {"points": [(m, pos, vel, ding)], "springs": [(i, j, x_or_None, k, string)]}"""
    s = L2System()
    pts, sps = system["points"], system["springs"]
    s.n_point, s.n_spring = len(pts), len(sps)
    P = [np.array(p[1], dtype=np.float32) for p in pts]
    for n, (m, pos, vel, ding) in enumerate(pts):
        s.mass[n], s.ding[n] = float(m), 1 if ding else 0
    for q, (i, j, x, k, string) in enumerate(sps):
        s.si[q], s.sj[q] = i, j
        if x is None:                                       # add_spring (env.py:105-107)
            x = np.linalg.norm(P[i] - P[j]).astype(np.float32)
        s.sx[q], s.sk[q], s.sstring[q] = np.float32(x), np.float32(k), 1 if string else 0
    return s


def make_l2_params(gravity=(0, -9.8, 0), damping=0.99, ground=True, ground_level=-50, ground_restitution=0.8,
                   air_resistance=0.01, friction=0.5, time_step=0.01) -> L2Params:
    p = L2Params()
    g = np.array(gravity, dtype=np.float32)
    for c in range(3):
        p.gravity[c] = g[c]
    p.damping, p.drag_c = np.float32(damping), np.float32(-0.5 * air_resistance)
    p.ground_level, p.restitution = np.float32(ground_level), np.float32(ground_restitution)
    p.friction, p.dt, p.min_dist, p.ground = np.float32(friction), np.float32(time_step), np.float32(16e-36), int(bool(ground))
    return p


def l2_init_state(system, E: int):
    pos = np.array([p[1] for p in system["points"]], np.float32).reshape(-1)
    vel = np.array([p[2] for p in system["points"]], np.float32).reshape(-1)
    return dict(pos=np.repeat(pos[:, None], E, 1).copy(), vel=np.repeat(vel[:, None], E, 1).copy(),
                old_a=np.zeros((pos.size, E), np.float32))


def l2_step(sysm: L2System, prm: L2Params, st: dict, n_steps: int = 1):
    l = lib()
    assert l.wgo_sizeof_l2_system() == C.sizeof(L2System) and l.wgo_sizeof_l2_params() == C.sizeof(L2Params)
    E = st["pos"].shape[1]
    rc = l.wgo_l2_step(C.byref(sysm), C.byref(prm), C.c_int64(E), C.c_int32(n_steps),
                       _ptr(st["pos"], C.c_float), _ptr(st["vel"], C.c_float), _ptr(st.get("old_a"), C.c_float))
    if rc != 0:
        raise RuntimeError(f"wgo_l2_step failed: {rc}")
