// throughput of packed fp32 (FFMA2 / FADD2 / FMUL2, sm_100) against scalar FFMA / FADD / FMUL: same flops, half the instructions
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x[8]; float2 y[4];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 4; i++) y[i] = make_float2(x[2 * i], x[2 * i + 1]);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], a, b);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 4; i++) y[i] = __ffma2_rn(y[i], a2, b2);
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = __fadd_rn(__fmul_rn(x[i], a), b);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) y[i] = __fadd2_rn(__fmul2_rn(y[i], a2), b2);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += y[i].x + y[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 2000;
    const char* names[4] = {"scalar FFMA       ", "packed FFMA2      ", "scalar FMUL + FADD", "packed FMUL2+FADD2"};
    for (int warps = 8; warps <= 64; warps *= 2)
    for (int mode = 0; mode < 4; mode++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int blocks = 148 * warps / 8;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 1) k<1><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 2) k<2><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 3) k<3><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * 256 * iters * 8 * 8 * ((mode & 2) ? 2 : 1);   // scalar-equivalent fp32 instructions
        printf("%d warps/SM  %s: %.3f ms, %.1f scalar-equivalent fp32 ops / clk / SM\n", warps, names[mode], ms, ops / (ms * 1e-3) / 1.965e9 / 148);
    }
    return 0;
}
