// does the 3-FMA exact-remainder quotient equal IEEE division for ALL float32 x, for arbitrary (non-integer) divisors m?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float div_smallint(float x, float m, float r) {
    const float q0 = __fmul_rn(x, r);
    const float rem = fmaxf(__fmaf_rn(-m, q0, x), -3.402823466e38f);
    return __fmaf_rn(rem, r, q0);
}
__device__ __forceinline__ double div_d(double f, double m, double rd) {
    const double q0 = __dmul_rn(f, rd);
    const double rem = fma(-m, q0, f);
    return (fabs(q0) == (double)__int_as_float(0x7f800000)) ? q0 : fma(rem, rd, q0);
}
__global__ void k(float m, float r, double md, double rd, unsigned long long* out) {
    unsigned long long bad32 = 0, bad64 = 0, badsub = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)i);
        const float a = div_smallint(x, m, r), b = __fdiv_rn(x, m);
        const bool same = (__float_as_uint(a) == __float_as_uint(b)) || (a != a && b != b) || (a == 0.0f && b == 0.0f);
        if (!same) { bad32++; if (fabsf(b) < 1.17549435e-38f) badsub++; }
        if (x == x) {
            const double qa = div_d((double)x, md, rd), qb = __ddiv_rn((double)x, md);
            if (!(qa == qb || (qa != qa && qb != qb))) bad64++;
        }
    }
    atomicAdd(out, bad32); atomicAdd(out + 1, bad64); atomicAdd(out + 2, badsub);
}
int main() {
    unsigned long long* out; cudaMallocManaged(&out, 24);
    const double ms[] = {0.1, 0.3, 0.7, 1.5, 2.5, 7.3, 0.05, 12.75, 100.1, 3.14159, 0.333333, 6.0, 10.0, 0.9, 1.1};
    for (double md0 : ms) {
        const float m = (float)md0; const double md = (double)m;      // the reference's mass as float32 and as the double it divides by
        out[0] = out[1] = out[2] = 0;
        k<<<148 * 16, 256>>>(m, 1.0f / m, md0, 1.0 / md0, out);
        cudaDeviceSynchronize();
        printf("m = %-10g float32 quotient mismatches %llu (of which IEEE result subnormal: %llu); float64 quotient (f float32, m python float) mismatches %llu\n", md0, out[0], out[2], out[1]);
    }
    return 0;
}
