// microbenchmarks: (A) tcgen05.mma issue rate with several issuing warps per CTA, (B) tcgen05.ld / st throughput
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
__host__ __device__ constexpr uint32_t idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// W issuing warps, each `reps` MMAs (TS, N = 64, unrolled by 8 with accumulate) into its own 64 columns
template <int W, int N>
__global__ void __launch_bounds__(256) kA(int reps, long long* out) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 16384; i += 256) sm[i] = 0.001f * (i & 63);
    if (tid == 0) { for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[i]))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp < W) {
        const uint64_t db = desc(smem_u32(sm), N);
        const uint32_t dcol = tmem + warp * 64, acol = tmem + 256 + warp * 16;
        constexpr uint32_t id = idesc(N);
        long long t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < reps; r += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" :: "r"(dcol), "r"(acol), "l"(db + (uint64_t)(u & 1) * 128), "r"(id), "r"(1) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar[warp])) : "memory");
        }
        long long t1 = clock64();
        uint32_t ok = 0; int it = 0;
        while (!ok && it++ < (1 << 22)) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(0) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0 && (tid & 31) == 0) { out[2 * warp] = t1 - t0; out[2 * warp + 1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}
// tcgen05.ld / st throughput: every warp of the CTA moves 32 columns x 32 lanes (4 KB) per instruction, reps times
template <int ST>
__global__ void __launch_bounds__(512) kB(int reps, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 32;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = tid + i;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        if (ST) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                    "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
        } else {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] ^ v[31];
        }
    }
    if (ST) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}
template <int W, int N> void runA(long long* out) {
    const int reps = 512;
    cudaFuncSetAttribute(kA<W, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int i = 0; i < 8; i++) out[i] = 0;
    kA<W, N><<<148, 256, 65536>>>(reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("A: %d issuing warps, N %3d:", W, N);
    for (int w = 0; w < W; w++) printf("  warp %d issue %.1f done %.1f cyc/MMA", w, (double)out[2 * w] / reps, (double)out[2 * w + 1] / reps);
    printf(" (%s)\n", cudaGetErrorString(e));
}
int main() {
    long long* out; cudaMallocManaged(&out, 64);
    uint32_t* sink; cudaMalloc(&sink, 4);
    runA<1, 64>(out); runA<2, 64>(out); runA<4, 64>(out); runA<1, 16>(out); runA<4, 16>(out); runA<1, 128>(out); runA<2, 128>(out);
    for (int st = 0; st < 2; st++)
        for (int threads : {128, 256, 512}) {
            const int reps = 256;
            *out = 0;
            if (st) kB<1><<<148, threads>>>(reps, out, sink); else kB<0><<<148, threads>>>(reps, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            printf("B: tcgen05.%s 32x32b.x32, %d warps: %.1f cycles per instruction per warp, %.1f B/cycle/SM (%s)\n", st ? "st" : "ld", threads / 32,
                   (double)*out / reps, 4096.0 * (threads / 32) * reps / (double)*out, cudaGetErrorString(e));
        }
    return 0;
}
