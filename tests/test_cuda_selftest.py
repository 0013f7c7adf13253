"""GPU: exhaustive / randomised self-tests of the exact-arithmetic primitives the kernels are built from
(csrc/wg_math.cuh), each against the IEEE operation it replaces, through the C ABI (wg_selftest_*).  The reference's
arithmetic is NumPy's IEEE float32 / float64 (gym/optimized_engine.py:104-106, gym/optimized_walker.py:45-67); the
kernels replace divisions and square roots by shorter exact sequences, and these tests are the proof that the
replacement never changes a bit: the count of mismatching inputs must be zero."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(fn, *args):
    import torch
    from walker_gym_b200 import _lib
    lib = _lib.load()
    out = torch.zeros(1, dtype=torch.int64, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)
    with torch.cuda.device(DEV):
        _lib.check(getattr(lib, fn)(*args, out.data_ptr(), stream), fn)
    torch.cuda.synchronize()
    return int(out.item())


def _body_table_masses():
    from walker_gym_b200 import BODIES
    ms = set()
    for b in BODIES.values():
        for m, _ in b["points"]:
            if float(m) == int(m) and 1 <= int(m) <= 2048:
                ms.add(int(m))
        ms.add(len(b["points"]))              # the division by the number of masses (centroid / means)
    return sorted(ms)


def test_div_smallint_exhaustive_over_all_float32_for_every_body_table_mass():
    """x / m for every one of the 2^32 float32 bit patterns x (normals, subnormals, +-0, +-inf, every NaN) and every
    integer mass / mass count of the in-tree body tables: zero mismatches against IEEE division."""
    bad = {m: _run("wg_selftest_div_smallint", C.c_float(m), 0, 1 << 32) for m in _body_table_masses()}
    assert all(v == 0 for v in bad.values()), bad


def test_div_smallint_exhaustive_over_all_float32_for_every_integer_divisor_up_to_2048():
    """The whole admitted divisor range m = 2 .. 2048 x all 2^32 float32 x (8.8e12 quotients)."""
    import torch
    from walker_gym_b200 import _lib
    lib = _lib.load()
    out = torch.zeros(2049, dtype=torch.int64, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)
    with torch.cuda.device(DEV):
        for m in range(2, 2049):
            _lib.check(lib.wg_selftest_div_smallint(C.c_float(m), 0, 1 << 32, out[m:].data_ptr(), stream), "selftest")
    torch.cuda.synchronize()
    bad = {m: int(v) for m, v in enumerate(out.tolist()) if v}
    assert not bad, bad


def test_sqrt_exhaustive_over_all_non_negative_float32():
    assert _run("wg_selftest_sqrt") == 0


@pytest.mark.parametrize("general", [0, 1])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_div3_len_on_2_pow_33_random_and_boundary_inputs(mode, general):
    """direction / current_dist with one shared reciprocal vs three IEEE divisions: independent random bit patterns
    (subnormals, inf and NaN included), L = norm(d), and exponents at the fast-path guards."""
    n = (1 << 33) if mode == 0 else (1 << 32)
    assert _run("wg_selftest_div3", mode, general, 1234 + mode, n) == 0


@pytest.mark.parametrize("m", [1.0, 2.0, 3.0, 5.0, 7.0, 13.0, 16.0, 2048.0, 0.1, 2.5])
def test_forced_list_on_random_pairs(m):
    """float32(float64(a) + float64(f) / m): the exact-remainder double quotient vs IEEE double division."""
    assert _run("wg_selftest_forced_list", C.c_double(m), 99, 1 << 32) == 0


def test_selftest_detects_a_wrong_reciprocal():
    """The harness itself: a divisor outside the proven range is rejected, and the counters do count (a deliberately
    inexact identity -- dividing by 3 through the m = 5 path cannot be run through the ABI, so check the counter
    plumbing with the sqrt test's complement instead: zero stays zero, and results accumulate across calls)."""
    import torch
    from walker_gym_b200 import _lib
    lib = _lib.load()
    out = torch.full((1,), 5, dtype=torch.int64, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)
    assert lib.wg_selftest_div_smallint(C.c_float(4096.0), 0, 16, out.data_ptr(), stream) == -1
    assert lib.wg_selftest_div_smallint(C.c_float(2.5), 0, 16, out.data_ptr(), stream) == -1
    _lib.check(lib.wg_selftest_div_smallint(C.c_float(3.0), 0, 1 << 20, out.data_ptr(), stream), "selftest")
    torch.cuda.synchronize()
    assert int(out.item()) == 5               # adds to the counter; no mismatches
