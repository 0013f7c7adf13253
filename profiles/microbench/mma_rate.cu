// microbenchmark: cycles per tcgen05.mma kind::tf32 (M=128, N, K=8), A from TMEM (TS) or smem (SS), back to back
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t rows) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((rows * 16u) >> 4 & 0x3fffu) << 16;
    d |= (uint64_t)(128u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc(int n, int kind_tf32) {
    // tf32: a,b format 2 ; f16 kind with bf16: format 1
    return (1u << 4) | ((kind_tf32 ? 2u : 1u) << 7) | ((kind_tf32 ? 2u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
template <int MODE>   // 0 TS tf32, 1 SS tf32, 2 TS bf16(kind::f16), 3 SS bf16
__global__ void __launch_bounds__(128) k(int n, int reps, int nacc, long long* out) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += 128) sm[i] = 0.001f * (i & 63);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint64_t db = desc(smem_u32(sm), n), da = desc(smem_u32(sm + 8192), 128);
        const uint32_t id = idesc(n, MODE < 2);
        for (int pass = 0; pass < 2; pass++) {
            t0 = clock64();
            for (int r = 0; r < reps; r++) {
                const uint32_t dcol = tmem + (nacc > 1 ? (uint32_t)((r % nacc) * (n < 64 ? 64 : n)) % 256u : 0u);
                if (MODE == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" :: "r"(dcol), "r"(tmem + 256 - 16), "l"(db), "r"(id), "r"(1) : "memory");
                if (MODE == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" :: "r"(dcol), "l"(da), "l"(db), "r"(id), "r"(1) : "memory");
                if (MODE == 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" :: "r"(dcol), "r"(tmem + 256 - 16), "l"(db), "r"(id), "r"(1) : "memory");
                if (MODE == 3) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" :: "r"(dcol), "l"(da), "l"(db), "r"(id), "r"(1) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0; int it = 0;
            while (!ok && it++ < (1 << 22)) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(pass) : "memory");
            t1 = clock64();
        }
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}
int main() {
    long long* out; cudaMallocManaged(&out, 8);
    const int reps = 512;
    const char* names[4] = {"TS tf32", "SS tf32", "TS bf16", "SS bf16"};
    for (int mode = 0; mode < 4; mode++)
        for (int ctas = 1; ctas <= 2; ctas++)
            for (int nacc = 1; nacc <= 2; nacc++)
                for (int n : {16, 32, 64, 128}) {
                    if (nacc == 2 && n == 128 ) { }
                    void (*kern)(int, int, int, long long*) = mode == 0 ? k<0> : mode == 1 ? k<1> : mode == 2 ? k<2> : k<3>;
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
                    *out = 0;
                    kern<<<148 * ctas, 128, 65536>>>(n, reps, nacc, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    printf("%s ctas/SM %d nacc %d N %3d: %.1f cycles per MMA (%s)\n", names[mode], ctas, nacc, n, (double)*out / reps, cudaGetErrorString(e));
                }
    return 0;
}
