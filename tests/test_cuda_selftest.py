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


def _run(fn, *args, dump=False):
    """Runs one wg_selftest_* call; returns the mismatch count (and, with dump, the details of one failing input)."""
    import torch
    from walker_gym_b200 import _lib
    lib = _lib.load()
    out = torch.zeros(2, dtype=torch.int64, device=DEV)
    dmp = torch.zeros(10, dtype=torch.float32, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)
    extra = (out.data_ptr(), dmp.data_ptr(), stream) if dump else (out.data_ptr(), stream)
    with torch.cuda.device(DEV):
        _lib.check(getattr(lib, fn)(*args, *extra), fn)
    torch.cuda.synchronize()
    if dump:
        return int(out[0].item()), int(out[1].item()) - 1, [float(v) for v in dmp.tolist()], dmp.view(torch.int32).tolist()
    return int(out[0].item())


def _admitted(m: int) -> bool:
    """The divisors the host hands to the 3-FMA quotient (make_const_div): 1, powers of two, odd integers <= 2047."""
    return m == 1 or (m & (m - 1)) == 0 or (m % 2 == 1 and 3 <= m <= 2047)


def _body_table_divisors():
    from walker_gym_b200 import BODIES
    ms = set()
    for b in BODIES.values():
        for m, _ in b["points"]:
            if float(m) == int(m) and 1 <= int(m) <= 2048:
                ms.add(int(m))
        ms.add(len(b["points"]))              # the division by the number of masses (centroid / means)
    return sorted(ms)


def test_div_smallint_exhaustive_over_all_float32_for_every_body_table_divisor():
    """x / m for every one of the 2^32 float32 bit patterns x (normals, subnormals, +-0, +-inf, every NaN) and every
    integer mass / mass count of the in-tree body tables that takes the 3-FMA quotient: zero mismatches against IEEE
    division.  Even non-powers of two (6 masses: humanb, box4) are not admitted -- they take the IEEE division, because
    a subnormal quotient by such a divisor can be an exact tie (this test found it) -- and the ABI refuses them here."""
    import torch
    from walker_gym_b200 import _lib
    divisors = _body_table_divisors()
    assert {1, 3, 4, 5, 6, 13, 16} <= set(divisors)
    bad = {m: _run("wg_selftest_div_smallint", C.c_float(m), 0, 1 << 32) for m in divisors if _admitted(m)}
    assert all(v == 0 for v in bad.values()), bad
    out = torch.zeros(2, dtype=torch.int64, device=DEV)
    for m in divisors:
        if not _admitted(m):
            assert _lib.load().wg_selftest_div_smallint(C.c_float(m), 0, 16, out.data_ptr(), None) == -1, m


def test_div_smallint_exhaustive_over_all_float32_for_every_admitted_divisor_up_to_2048():
    """The whole admitted divisor range -- every odd m in 3 .. 2047 and every power of two up to 2048 -- x all 2^32
    float32 x (4.4e12 quotients)."""
    import torch
    from walker_gym_b200 import _lib
    lib = _lib.load()
    out = torch.zeros(2050, dtype=torch.int64, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)
    ms = [m for m in range(2, 2049) if _admitted(m)]
    with torch.cuda.device(DEV):
        for m in ms:
            _lib.check(lib.wg_selftest_div_smallint(C.c_float(m), 0, 1 << 32, out[m:].data_ptr(), stream), "selftest")
    torch.cuda.synchronize()
    bad = {m: int(out[m].item()) for m in ms if int(out[m].item())}
    assert len(ms) == 1023 + 11 and not bad, bad


def test_sqrt_exhaustive_over_all_non_negative_float32():
    assert _run("wg_selftest_sqrt") == 0


@pytest.mark.parametrize("mode,general", [(0, 1), (1, 0), (1, 1), (2, 0), (2, 1)])
def test_div3_len_on_2_pow_33_random_and_boundary_inputs(mode, general):
    """direction / current_dist with one shared reciprocal vs three IEEE divisions: independent random bit patterns
    (subnormals, inf and NaN included; arbitrary numerators are the package-lineage variant's contract, general = 1),
    L = norm(d) (the relation the step kernels have), and exponents at the fast-path guards."""
    n = (1 << 33) if mode == 0 else (1 << 32)
    bad, idx, vals, bits = _run("wg_selftest_div3", mode, general, 1234 + mode, 0, n, dump=True)
    assert bad == 0, (bad, idx, vals, [hex(b & 0xFFFFFFFF) for b in bits])


@pytest.mark.parametrize("m", [1.0, 2.0, 3.0, 5.0, 6.0, 7.0, 12.0, 13.0, 16.0, 2047.0, 2048.0, 0.1, 2.5])
def test_forced_list_on_random_pairs(m):
    """float32(float64(a) + float64(f) / m): the exact-remainder double quotient vs IEEE double division."""
    assert _run("wg_selftest_forced_list", C.c_double(m), 99, 1 << 32) == 0


def test_selftest_plumbing():
    """The harness itself: a divisor outside the admitted class is rejected; the counter accumulates across calls."""
    import torch
    from walker_gym_b200 import _lib
    lib = _lib.load()
    out = torch.tensor([5, 0], dtype=torch.int64, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)
    for m in (4096.0, 2.5, 6.0, 0.0):
        assert lib.wg_selftest_div_smallint(C.c_float(m), 0, 16, out.data_ptr(), stream) == -1, m
    _lib.check(lib.wg_selftest_div_smallint(C.c_float(3.0), 0, 1 << 20, out.data_ptr(), stream), "selftest")
    torch.cuda.synchronize()
    assert int(out[0].item()) == 5               # adds to the counter; no mismatches
