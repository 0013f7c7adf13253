"""Host-side plumbing for the host-buffer calls (``wg_step_host`` / ``HostStepPipeline``): CPU affinity of a rank
next to its GPU, and pinned buffers allocated through the C ABI (``wg_host_alloc``) by the bound thread.

The reference keeps its arrays in pageable NumPy memory (gym/optimized_engine.py:84-89); PCIe copies from / to such
memory are staged by the driver.  A caller that wants the wire rate copies its actions into, and reads its results
from, buffers made here.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import List, Optional, Sequence

import torch

from . import _lib


def _nvml_cpu_affinity(index: int) -> Optional[List[int]]:
    """CPUs the driver reports as local to GPU ``index`` (None when NVML is missing)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        return [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
    except Exception:
        return None


def bind_to_device(local_rank: int, ranks_on_node: int = 1) -> dict:
    """Pin the calling process to CPUs next to GPU ``local_rank``: the CPUs NVML reports as local to the GPU,
    intersected with what the process may use, then split evenly between the ranks that share them, so that every
    rank's copy-issuing thread and its pinned pages sit on the GPU's NUMA node and no two ranks share a core.
    Call it before allocating pinned memory.  Returns what was done (for logs)."""
    allowed = sorted(os.sched_getaffinity(0))
    near = _nvml_cpu_affinity(local_rank)
    pool = [c for c in allowed if near is None or c in set(near)] or allowed
    n = max(1, int(ranks_on_node))
    per = max(1, len(pool) // n)
    mine = pool[(local_rank % n) * per: (local_rank % n) * per + per] or pool
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        mine = allowed
    return {"cpus": mine, "gpu_local_cpus": None if near is None else len(near), "allowed": len(allowed)}


_live = {}


def pinned_empty(shape: Sequence[int], dtype: torch.dtype = torch.float32, write_combined: bool = False) -> torch.Tensor:
    """A pinned, device-mapped host tensor from ``wg_host_alloc`` (cudaHostAlloc by the calling thread, zeroed).
    ``tensor.is_pinned()`` is true; the memory is released when the tensor is garbage collected."""
    lib = _lib.load()
    n = 1
    for s in shape:
        n *= int(s)
    nbytes = max(1, n * torch.empty((), dtype=dtype).element_size())
    p = C.c_void_p()
    _lib.check(lib.wg_host_alloc(C.byref(p), nbytes, 1 if write_combined else 0), "wg_host_alloc")
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    t = torch.frombuffer(buf, dtype=dtype, count=n).reshape(tuple(int(s) for s in shape))
    addr = p.value
    _live[addr] = buf
    weakref.finalize(t.untyped_storage(), _release, addr)
    return t


def _release(addr: int) -> None:
    _live.pop(addr, None)
    try:
        _lib.load().wg_host_free(C.c_void_p(addr))
    except Exception:
        pass
