"""Build ``libwalkergym_b200.so`` in-tree with plain nvcc for sm_100a.

One object per translation unit, compiled in parallel, linked into a shared
library next to this file (it is git-ignored but travels with the repo
snapshot to the GPU box).  ``-fmad=false`` is part of the numerics contract:
the kernels reproduce the reference's separately rounded float32 operations.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OUT = os.environ.get("WG_LIB_OUT", os.path.join(HERE, "libwalkergym_b200.so"))
OBJ_DIR = os.environ.get("WG_OBJ_DIR", os.path.join(HERE, "build"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
NVCC_FLAGS += os.environ.get("WG_EXTRA_NVCC_FLAGS", "").split()      # tuning experiments only


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    stamp, digest = os.path.join(OBJ_DIR, "stamp"), _digest(deps)
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, srcs))
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log for _, log in results))
    if verbose:
        print("\n".join(log for _, log in results))
    r = subprocess.run([nvcc, "-shared", "-o", OUT, *[o for o, _ in results], "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
