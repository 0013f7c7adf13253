"""Shared helpers: load a golden fixture and replay it through a stepper.

A *stepper* is anything with the oracle's calling convention (the C oracle in
the CPU tests, the CUDA library through its C ABI in the GPU tests), so the
same replay code checks both against the reference's recorded outputs.
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def trajectory_names():
    return [n for n in golden_names() if not n.startswith(("batch_", "compat_", "l2_", "env_state_", "f64act_", "getstat_"))]


def f64_names():
    return golden_names("f64act_")


def l2_names():
    return [n for n in golden_names("l2_") if n != "l2_bodies"]


def load_l2(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in ("pos", "vel", "old_a")}
    sysm = json.loads(str(z["system"]))
    g["system"] = {"points": [(m, tuple(p), tuple(v), bool(d)) for m, p, v, d in sysm["points"]],
                   "springs": [(i, j, x, k, bool(s)) for i, j, x, k, s in sysm["springs"]]}
    g["env_kwargs"] = json.loads(str(z["env_kwargs"]))
    return g


def batch_names():
    return golden_names("batch_")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["spec"] = json.loads(str(g["spec"]))
    # JSON turns tuples into lists; normalise
    g["spec"]["points"] = [(m, tuple(p), bool(f)) for m, p, f in g["spec"]["points"]]
    g["spec"]["muscles"] = [(i, j, dict(kw)) for i, j, kw in g["spec"]["muscles"]]
    g["spec"]["skeletons"] = [(i, j, dict(kw)) for i, j, kw in g["spec"]["skeletons"]]
    g["env_kwargs"] = json.loads(str(g["env_kwargs"]))
    for k in ("k_sub", "max_steps", "reset_on_done", "integrator"):
        g[k] = int(g[k]) if k in g else None
    return g


def same(a, b):
    """Bit-level equality up to NaN payload and the sign of zero."""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.array_equal(a, b, equal_nan=True))


def soa(x):
    """[N,3] (or [E,N,3]) -> SoA [(n*3+c), E]."""
    x = np.asarray(x, np.float32)
    if x.ndim == 2:
        x = x[None]
    E = x.shape[0]
    return np.ascontiguousarray(x.reshape(E, -1).T)


def aos(x, N):
    """SoA [(n*3+c), E] -> [E, N, 3]."""
    return np.ascontiguousarray(np.asarray(x).T).reshape(-1, N, 3)


def mask_bits(mask, N):
    return np.array([[(int(m) >> n) & 1 for n in range(N)] for m in np.asarray(mask).reshape(-1)], dtype=bool)
