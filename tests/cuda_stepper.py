"""The CUDA library behind the oracle's calling convention (tests only).

Every call uploads the numpy state, invokes ``wg_step`` / ``wg_reset`` through
the C ABI on ``cuda:0`` and downloads the result, so the golden replay code in
``test_oracle_golden.py`` can drive the oracle and the GPU identically.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from walker_gym_b200 import _lib
from walker_gym_b200.batched import make_params as _make_params
from walker_gym_b200.topology import topology_from_spec

DEV = "cuda:0"
obs_layout = 0          # module-level switches the tests flip
force_generic = False
use_tma = False


class Body:
    def __init__(self, spec):
        self.topo = topology_from_spec(spec)
        self.n_mass, self.n_muscle, self.n_spring = self.topo.n_mass, self.topo.n_muscle, self.topo.n_spring
        self.srest = self.topo.srest
        self.tmpl_pos = self.topo.tmpl_pos


def make_body(spec):
    return Body(spec)


def make_params(**kw):
    return _make_params(**kw)


def obs_dim(body, in3d):
    return 3 * (3 if in3d else 2) * body.n_mass + body.n_muscle


def init_state(body, E):
    N, M = body.n_mass, body.n_muscle
    tp = np.array(body.tmpl_pos[: N * 3], dtype=np.float32)
    return dict(pos=np.repeat(tp[:, None], E, axis=1).copy(), vel=np.zeros((N * 3, E), np.float32),
                old_a=np.zeros((N * 3, E), np.float32),
                mx=np.repeat(np.array(body.srest[:M], dtype=np.float32)[:, None], E, axis=1).copy().reshape(M, E),
                steps=np.zeros(E, np.int32))


def _dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return C.c_void_p(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)


def _upload(st):
    return {k: _dev(st[k]) for k in ("pos", "vel", "old_a", "mx", "steps")}


def _download(st, d):
    for k in ("pos", "vel", "old_a", "mx", "steps"):
        st[k][...] = d[k].cpu().numpy()


def step(body, prm, st, action, *, want_info=True, ep_ret=None, fin_stats=None, noise=None):
    lib = _lib.load()
    E = st["pos"].shape[1]
    D = obs_dim(body, prm.in3d)
    d = _upload(st)
    act = _dev(np.ascontiguousarray(action, dtype=np.float32).reshape(E, -1))
    f32 = dict(dtype=torch.float32, device=DEV)
    out = dict(obs=torch.zeros((E, D) if obs_layout == 0 else (D, E), **f32), reward=torch.zeros(E, **f32),
               done=torch.zeros(E, dtype=torch.uint8, device=DEV),
               contact_pre=torch.zeros(E, dtype=torch.int32, device=DEV),
               contact_post=torch.zeros(E, dtype=torch.int32, device=DEV))
    if want_info:
        out["energy"] = torch.zeros(E, **f32)
        out["centroid"] = torch.zeros(3, E, **f32)
    d_ep, d_fin, d_noise = _dev(ep_ret), _dev(fin_stats), _dev(noise)
    b = _lib.WgBuffers()
    b.pos, b.vel, b.old_a, b.mx, b.steps = (_ptr(d[k]) for k in ("pos", "vel", "old_a", "mx", "steps"))
    if body.n_muscle == 0:
        b.mx = b.steps
    b.action, b.act_dim, b.obs_layout = _ptr(act), act.shape[1], obs_layout
    b.obs, b.reward, b.done = _ptr(out["obs"]), _ptr(out["reward"]), _ptr(out["done"])
    b.contact_pre, b.contact_post = _ptr(out["contact_pre"]), _ptr(out["contact_post"])
    b.energy, b.centroid = _ptr(out.get("energy")), _ptr(out.get("centroid"))
    b.ep_ret, b.fin_stats, b.noise = _ptr(d_ep), _ptr(d_fin), _ptr(d_noise)
    old = lib.wg_force_generic(1 if force_generic else 0)
    old_tma = lib.wg_set_tuning(_lib.TUNE_TMA, 1 if use_tma else 0)
    try:
        rc = lib.wg_step(C.byref(body.topo), C.byref(prm), C.byref(b), E, _stream())
    finally:
        lib.wg_force_generic(old)
        lib.wg_set_tuning(_lib.TUNE_TMA, old_tma)
    _lib.check(rc, "wg_step")
    torch.cuda.synchronize()
    _download(st, d)
    if ep_ret is not None:
        ep_ret[...] = d_ep.cpu().numpy()
    if fin_stats is not None:
        fin_stats[...] = d_fin.cpu().numpy()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    if obs_layout == 1:
        res["obs"] = np.ascontiguousarray(res["obs"].T)
    res["contact_pre"] = res["contact_pre"].astype(np.uint32)
    res["contact_post"] = res["contact_post"].astype(np.uint32)
    return res


def reset(body, prm, st, *, mode=1, mask=None, noise=None):
    lib = _lib.load()
    E = st["pos"].shape[1]
    D = obs_dim(body, prm.in3d)
    d = _upload(st)
    obs = torch.zeros((E, D) if obs_layout == 0 else (D, E), dtype=torch.float32, device=DEV)
    d_noise, d_mask = _dev(noise), _dev(mask)
    b = _lib.WgBuffers()
    b.pos, b.vel, b.old_a, b.mx, b.steps = (_ptr(d[k]) for k in ("pos", "vel", "old_a", "mx", "steps"))
    if body.n_muscle == 0:
        b.mx = b.steps
    b.obs, b.obs_layout, b.noise = _ptr(obs), obs_layout, _ptr(d_noise)
    rc = lib.wg_reset(C.byref(body.topo), C.byref(prm), C.byref(b), E, mode, _ptr(d_mask), _stream())
    _lib.check(rc, "wg_reset")
    torch.cuda.synchronize()
    _download(st, d)
    o = obs.cpu().numpy()
    return np.ascontiguousarray(o.T) if obs_layout == 1 else o


# ---- package lineage (Environment.update_physics) through wg_pkg_update_physics -------------------------
def make_l2_system(system):
    import walker_oracle as wo
    o = wo.make_l2_system(system)            # same host-side evaluation of rest lengths; copy into the ABI struct
    s = _lib.WgPkgSystem()
    s.n_point, s.n_spring = o.n_point, o.n_spring
    for n in range(o.n_point):
        s.mass[n], s.fixed[n] = o.mass[n], o.ding[n]
    for q in range(o.n_spring):
        s.si[q], s.sj[q], s.srest[q], s.sk[q], s.sstring[q] = o.si[q], o.sj[q], o.sx[q], o.sk[q], o.sstring[q]
    return s


def make_l2_params(**kw):
    import walker_oracle as wo
    o = wo.make_l2_params(**kw)
    p = _lib.WgPkgParams()
    for c in range(3):
        p.gravity[c] = o.gravity[c]
    p.damping, p.drag_c, p.ground_level, p.restitution = o.damping, o.drag_c, o.ground_level, o.restitution
    p.friction, p.dt, p.min_dist, p.ground = o.friction, o.dt, o.min_dist, o.ground
    return p


def l2_init_state(system, E):
    import walker_oracle as wo
    return wo.l2_init_state(system, E)


def l2_step(sysm, prm, st, n_steps=1):
    lib = _lib.load()
    E = st["pos"].shape[1]
    d = {k: _dev(st[k]) for k in ("pos", "vel", "old_a")}
    old = lib.wg_force_generic(1 if force_generic else 0)
    try:
        rc = lib.wg_pkg_update_physics(C.byref(sysm), C.byref(prm), _ptr(d["pos"]), _ptr(d["vel"]), _ptr(d["old_a"]),
                                       E, n_steps, _stream())
    finally:
        lib.wg_force_generic(old)
    _lib.check(rc, "wg_pkg_update_physics")
    torch.cuda.synchronize()
    for k in ("pos", "vel", "old_a"):
        st[k][...] = d[k].cpu().numpy()


# ---- x64 mode (float64 actions) through wg_step_x64 ------------------------------------------------------------------
def make_x64(spec):
    import walker_oracle as wo
    o = wo.make_x64(spec)
    x = _lib.WgX64()
    for k in range(_lib.MAX_SPRING):
        x.sk_d[k], x.x0_d[k], x.mlo_d[k], x.mhi_d[k] = o.sk_d[k], o.x0_d[k], o.mlo_d[k], o.mhi_d[k]
    return x


def init_x64(body, xb, E):
    M = body.n_muscle
    x0 = np.array([xb.x0_d[m] for m in range(M)], np.float64)
    return np.repeat(x0[:, None], E, 1).copy().reshape(M, E), np.ones((M, E), np.uint8)


def step_x64(body, xb, prm, st, action64, *, want_info=True, noise=None):
    lib = _lib.load()
    E = st["pos"].shape[1]
    D = obs_dim(body, prm.in3d)
    d = _upload(st)
    d64, dw = _dev(st["mx64"]), _dev(st["mx_weak"])
    act = _dev(np.ascontiguousarray(action64, dtype=np.float64).reshape(E, -1))
    f32 = dict(dtype=torch.float32, device=DEV)
    out = dict(obs=torch.zeros((E, D) if obs_layout == 0 else (D, E), **f32), reward=torch.zeros(E, **f32),
               done=torch.zeros(E, dtype=torch.uint8, device=DEV),
               contact_pre=torch.zeros(E, dtype=torch.int32, device=DEV),
               contact_post=torch.zeros(E, dtype=torch.int32, device=DEV))
    if want_info:
        out["energy"] = torch.zeros(E, **f32)
        out["centroid"] = torch.zeros(3, E, **f32)
    d_noise = _dev(noise)
    b = _lib.WgBuffers()
    b.pos, b.vel, b.old_a, b.mx, b.steps = (_ptr(d[k]) for k in ("pos", "vel", "old_a", "mx", "steps"))
    if body.n_muscle == 0:
        b.mx = b.steps
        b.mx64, b.mx_weak = b.steps, b.steps
    else:
        b.mx64, b.mx_weak = _ptr(d64), _ptr(dw)
    b.action64, b.act_dim, b.obs_layout = _ptr(act), act.shape[1], obs_layout
    b.obs, b.reward, b.done = _ptr(out["obs"]), _ptr(out["reward"]), _ptr(out["done"])
    b.contact_pre, b.contact_post = _ptr(out["contact_pre"]), _ptr(out["contact_post"])
    b.energy, b.centroid, b.noise = _ptr(out.get("energy")), _ptr(out.get("centroid")), _ptr(d_noise)
    rc = lib.wg_step_x64(C.byref(body.topo), C.byref(xb), C.byref(prm), C.byref(b), E, _stream())
    _lib.check(rc, "wg_step_x64")
    torch.cuda.synchronize()
    _download(st, d)
    if body.n_muscle:
        st["mx64"][...] = d64.cpu().numpy()
        st["mx_weak"][...] = dw.cpu().numpy()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    if obs_layout == 1:
        res["obs"] = np.ascontiguousarray(res["obs"].T)
    res["contact_pre"] = res["contact_pre"].astype(np.uint32)
    res["contact_post"] = res["contact_post"].astype(np.uint32)
    return res
