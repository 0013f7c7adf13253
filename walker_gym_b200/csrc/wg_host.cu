// wg_host.cu -- host-side helpers of the C ABI that are not kernels of the step path:
//   * pinned host buffers for the host-buffer calls (wg_step_host / wg_step_multi_host), allocated and first-touched
//     by the calling thread so that they land on the NUMA node the caller bound itself to;
//   * exhaustive / randomised self-tests of the exact-arithmetic primitives of wg_math.cuh against the IEEE
//     operations they replace (the bit-exactness of every kernel rests on them).
#include <cstdio>
#include <cstring>

#include "wg_launch.cuh"

namespace wg {

// ---- self-test kernels ------------------------------------------------------------------------------------
// equal values (+0 == -0: the kernels' contract is "same bits up to NaN payload and the sign of zero" -- the exact
// quotient of -0 by a constant comes out as +0, and no operation of the step divides by, or takes the sign of, a zero)
__device__ __forceinline__ bool same_f32(float a, float b) { return (a != a && b != b) || a == b; }
__device__ __forceinline__ bool same_bits(float a, float b) { return (a != a && b != b) || __float_as_uint(a) == __float_as_uint(b); }
__device__ __forceinline__ void count_mismatch(bool bad, unsigned long long* out) {
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)__popc(m));
}

// div_smallint(x, m, RN(1/m)) against IEEE x / m for every float32 bit pattern x in [x0, x0 + n)
__global__ void selftest_div_smallint_kernel(float m, float r, uint64_t x0, uint64_t n, unsigned long long* out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t k = 0; k < rounds; k++, i += stride) {
        bool bad = false;
        if (i < n) {
            const float x = __uint_as_float((uint32_t)(x0 + i));
            bad = !same_f32(div_smallint(x, m, r), __fdiv_rn(x, m));
            // the packed twin the kernels use (FMUL2 / FFMA2): x in one half, a bijective scramble of it in the other
            const float y = __uint_as_float(((uint32_t)(x0 + i) * 0x9E3779B1u) ^ 0x5bd1e995u);
            const float2 q2 = div_smallint2(make_float2(x, y), m, r), p2 = div_smallint2(make_float2(y, x), m, r);
            bad = bad || !same_f32(q2.x, __fdiv_rn(x, m)) || !same_f32(q2.y, __fdiv_rn(y, m)) || !same_f32(p2.y, __fdiv_rn(x, m));
        }
        count_mismatch(bad, out);
    }
}

// forced_list(a, f, m, RN(1/m), kind) against float32(float64(a) + float64(f) / m) on Philox-random bit patterns
__global__ void selftest_forced_list_kernel(double m, double rd, int kind, uint32_t seed, uint64_t n, unsigned long long* out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t k = 0; k < rounds; k++, i += stride) {
        bool bad = false;
        if (i < n) {
            uint32_t c[4] = { (uint32_t)i, (uint32_t)(i >> 32), 0x464c5354u, 0u };
            philox4x32_10(c, seed, 0x5eedu);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float a = __uint_as_float(c[2 * q]), f = __uint_as_float(c[2 * q + 1]);
                const float want = (float)((double)a + (double)f / m);
                bad = bad || !same_f32(forced_list(a, f, m, rd, kind), want);
            }
        }
        count_mismatch(bad, out);
    }
}

// sqrt_rn(x) against IEEE sqrt for every non-negative float32 bit pattern and every NaN
__global__ void selftest_sqrt_kernel(unsigned long long* out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n = 1ull << 32, rounds = (n + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t k = 0; k < rounds; k++, i += stride) {
        bool bad = false;
        if (i < n) {
            const float x = __uint_as_float((uint32_t)i);
            if (!(x < 0.0f) && __float_as_uint(x) != 0x80000000u)       // a sum of squares is +0, positive or NaN
                bad = !same_bits(sqrt_rn(x), __fsqrt_rn(x));
        }
        count_mismatch(bad, out);
    }
}

// div3_len against `if L > 0: d = d / L` (gym/optimized_walker.py:52-54) on random inputs.
//   mode 0: d0..d2 and L independent random bit patterns (L >= 0: it is a norm)
//   mode 1: L = np_norm3(d) (the relation the kernels have), d random bit patterns
//   mode 2: boundary exponents: L around 2^-2 / 2^120 / subnormal / huge, quotients around 2^-100
// A NaN length is skipped: the reference leaves d untouched, the kernel makes it NaN, and every consumer multiplies
// d by a NaN factor derived from the same length (f_size, dk), so both give NaN -- checked at the spring level by
// the trajectory tests.
template <bool GENERAL>
__global__ void selftest_div3_kernel(int mode, uint32_t seed, uint64_t first, uint64_t n, unsigned long long* out, float* dump) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t k = 0; k < rounds; k++, i0 += stride) {
        bool bad = false;
        if (i0 < n) {
            const uint64_t i = first + i0;
            uint32_t c[4] = { (uint32_t)i, (uint32_t)(i >> 32), 0x44495633u, (uint32_t)mode };
            philox4x32_10(c, seed, 0x5eedu);
            float d[3] = { __uint_as_float(c[0]), __uint_as_float(c[1]), __uint_as_float(c[2]) };
            float L = __uint_as_float(c[3] & 0x7fffffffu);
            if (mode == 1) {
                L = np_norm3(d[0], d[1], d[2]);
            } else if (mode == 2) {
                const int sel = (c[3] >> 24) & 7;
                const int eL = sel == 0 ? 125 : sel == 1 ? 124 : sel == 2 ? 247 : sel == 3 ? 248 : sel == 4 ? 0 : sel == 5 ? 1
                                        : sel == 6 ? 254 : 127;
                L = __uint_as_float(((uint32_t)eL << 23) | (c[3] & 0x7fffffu));
#pragma unroll
                for (int q = 0; q < 3; q++) {           // quotient exponent within +-4 of 2^-100, or anything
                    const uint32_t bits = c[q];
                    if (bits & 0x40000000u) {
                        int e = eL - 100 + (int)((bits >> 26) & 7) - 4;
                        e = e < 0 ? 0 : (e > 254 ? 254 : e);
                        d[q] = __uint_as_float((bits & 0x807fffffu) | ((uint32_t)e << 23));
                    }
                }
            }
            if (L == L) {
                float q0 = d[0], q1 = d[1], q2 = d[2];
                div3_len<GENERAL>(q0, q1, q2, L);
                V3 pk = v3(d[0], d[1], d[2]);                    // the packed twin the kernels use: must return the same bits
                div3_len<GENERAL>(pk, L);
                bool twin = same_f32(pk.xy.x, q0) && same_f32(pk.xy.y, q1) && same_f32(pk.z, q2);
                if (!GENERAL && mode == 1) {                     // L == norm(d): the fused routine the springs call
                    V3 ud = v3(d[0], d[1], d[2]);
                    const float uL = unit_dir(ud);
                    twin = twin && same_f32(uL, L) && same_f32(ud.xy.x, q0) && same_f32(ud.xy.y, q1) && same_f32(ud.z, q2);
                }
                float w0 = d[0], w1 = d[1], w2 = d[2];
                if (L > 0.0f) { w0 = __fdiv_rn(d[0], L); w1 = __fdiv_rn(d[1], L); w2 = __fdiv_rn(d[2], L); }
                bad = !(twin && same_f32(q0, w0) && same_f32(q1, w1) && same_f32(q2, w2));      // -0 / L comes out as +0: sign of zero
                if (bad && atomicCAS(out + 1, 0ull, (unsigned long long)(i + 1)) == 0ull && dump) {
                    dump[0] = d[0]; dump[1] = d[1]; dump[2] = d[2]; dump[3] = L;
                    dump[4] = q0; dump[5] = q1; dump[6] = q2; dump[7] = w0; dump[8] = w1; dump[9] = w2;
                }
            }
        }
        count_mismatch(bad, out);
    }
}

static int run_selftest(unsigned long long* d_out, cudaStream_t s, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "self-test launch: %s", cudaGetErrorString(e));
    (void)d_out; (void)s; (void)what;
    return WG_OK;
}

}  // namespace wg

using namespace wg;

extern "C" {

int wg_host_alloc(void** out, uint64_t bytes, int write_combined) {
    if (!out || bytes == 0) return fail(WG_ERR_BAD_ARG, "wg_host_alloc: null / empty request%s", "");
    void* p = nullptr;
    const unsigned flags = cudaHostAllocPortable | cudaHostAllocMapped | (write_combined ? cudaHostAllocWriteCombined : 0u);
    cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, flags);
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e));
    memset(p, 0, (size_t)bytes);            // first touch from the calling thread: pages land on its NUMA node
    *out = p;
    return WG_OK;
}

int wg_host_free(void* p) {
    if (!p) return WG_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFreeHost: %s", cudaGetErrorString(e));
    return WG_OK;
}

int wg_selftest_div_smallint(float m, uint64_t x_begin, uint64_t x_count, uint64_t* d_mismatches, void* cuda_stream) {
    // the divisors the host ever hands to div_smallint: 1, powers of two, odd integers in [3, 2047] (make_const_div)
    const int kind = (m >= 1.0f && m <= 2048.0f) ? make_const_div(m).kind : 3;
    if (!d_mismatches || kind == 3 || x_begin + x_count > (1ull << 32))
        return fail(WG_ERR_BAD_ARG, "wg_selftest_div_smallint: m must be 1, a power of two <= 2048 or an odd integer in [3, 2047]; x range within 2^32%s", "");
    if (x_count == 0) return WG_OK;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    selftest_div_smallint_kernel<<<148 * 16, 256, 0, s>>>(m, 1.0f / m, x_begin, x_count, (unsigned long long*)d_mismatches);
    return run_selftest((unsigned long long*)d_mismatches, s, "div_smallint");
}

int wg_selftest_forced_list(double m, uint32_t seed, uint64_t n_pairs, uint64_t* d_mismatches, void* cuda_stream) {
    if (!d_mismatches || !(m > 0.0)) return fail(WG_ERR_BAD_ARG, "wg_selftest_forced_list: bad argument%s", "");
    if (n_pairs == 0) return WG_OK;
    // the kind the host picks for this mass (fill_args): 0 unit, 1 power of two, 2 integer in [2, 2048], 3 anything else
    const int kind = make_const_div((float)m).kind;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    selftest_forced_list_kernel<<<148 * 16, 256, 0, s>>>(m, 1.0 / m, kind, seed, (n_pairs + 1) / 2, (unsigned long long*)d_mismatches);
    return run_selftest((unsigned long long*)d_mismatches, s, "forced_list");
}

int wg_selftest_sqrt(uint64_t* d_mismatches, void* cuda_stream) {
    if (!d_mismatches) return fail(WG_ERR_BAD_ARG, "wg_selftest_sqrt: null counter%s", "");
    cudaStream_t s = (cudaStream_t)cuda_stream;
    selftest_sqrt_kernel<<<148 * 16, 256, 0, s>>>((unsigned long long*)d_mismatches);
    return run_selftest((unsigned long long*)d_mismatches, s, "sqrt");
}

int wg_selftest_div3(int mode, int general, uint32_t seed, uint64_t first, uint64_t n, uint64_t* d_mismatches, float* d_dump,
                     void* cuda_stream) {
    if (!d_mismatches || mode < 0 || mode > 2) return fail(WG_ERR_BAD_ARG, "wg_selftest_div3: mode must be 0, 1 or 2%s", "");
    if (n == 0) return WG_OK;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (general) selftest_div3_kernel<true><<<148 * 16, 256, 0, s>>>(mode, seed, first, n, (unsigned long long*)d_mismatches, d_dump);
    else selftest_div3_kernel<false><<<148 * 16, 256, 0, s>>>(mode, seed, first, n, (unsigned long long*)d_mismatches, d_dump);
    return run_selftest((unsigned long long*)d_mismatches, s, "div3_len");
}

}  // extern "C"
