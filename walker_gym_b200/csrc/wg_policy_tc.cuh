// wg_policy_tc.cuh -- the rollout policy (BASELINE config 5) on the 5th-generation tensor cores.
//
// Same function as policy_act_kernel (wg_policy.cuh): obs -> gaussian MLP (D -> 64 -> 64 -> {M means, 1 value}, tanh)
// -> sampled action, log-prob, value, one launch per env step, reading the observation as the step kernel wrote it.
// The two wide GEMMs run as tcgen05.mma (kind::tf32, M = 128 envs per tile, accumulators in tensor memory):
//
//   layer 1   A = the tile's observations [128 x K1] in TENSOR MEMORY: thread r owns TMEM lane r = env r, sanitises its
//             observation row and writes it with tcgen05.st; B = W1 [64 x K1] in shared memory (canonical K-major
//             layout, no swizzle); D -> TMEM columns [0, 64).  Row-major observation tiles (one contiguous 128 x D
//             block) arrive by TMA bulk copy into a double-buffered staging area, one tile ahead.
//   layer 2   A = tanh(layer 1) [128 x 72] in tensor memory (written there by the epilogue: an activation row never
//             touches shared memory), B = [W2, b2] [64 x 72] in smem, D -> TMEM columns [0, 64)
//   heads     M + 1 <= 9 outputs of depth 64: 192 MACs per env.  As MMAs they cost a third accumulate / commit / wait
//             round trip per tile plus the split and tcgen05.st of the second activation; instead the layer-2 epilogue
//             multiplies its 32 activations (in registers, float32) with the head weights (broadcast reads from shared
//             memory) and the two threads of an env add their halves through shared memory.
// The biases ride in the GEMMs: every A operand carries a constant 1 in the column after its last feature and every B
// operand its bias there, so the epilogues are tanh + split only.  256 threads per tile: two warps share each quarter
// of the TMEM lanes and split an env row's columns between them.
//
// With SPLIT every product is the 3-term error-compensated a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (float32-grade, like the
// mma.sync kernel): three MMAs per k-step into the same accumulator, hi / lo planes of A in TMEM columns [64, 136) /
// [144, 216) and of B in shared memory.
//
// What the measurements on B200 said (profiles/microbench/mma_rate*.cu, tc_trace.cu):
//  * one tcgen05.mma M128 N64 K8 costs its 32-cycle floor (N16: 9) when issued from WARP-UNIFORM code (warp 0, one
//    elected lane); issued under `if (threadIdx.x == 0)` the compiler wraps every instruction in a lane-election loop and
//    the issue alone costs 60-70 cycles per MMA -- 4000 of the first version's 10800 cycles per tile;
//  * the epilogues are bound by the XU pipe (ex2 + rcp per float32-grade tanh, 16 lanes / clock / SM), so the hi / lo
//    split uses integer rounding instead of cvt.rna.tf32 (also XU);
//  * Philox + Box-Muller (a ~150-instruction dependent chain per env) is evaluated by the threads that would otherwise
//    only wait for the layer-2 MMAs, and handed to the sampling threads through shared memory.
// Two CTAs per SM (2 x 256 TMEM columns, 2 x ~107 KB of shared memory) overlap one tile's epilogue with the other's MMAs.
// Every mbarrier wait is bounded: a descriptor mistake ends the kernel with an error flag instead of hanging the GPU.
#pragma once
#include "wg_policy.cuh"

namespace wg {

// development aid (profiles/microbench/tc_trace.cu): thread 0 of CTA 0 stamps the clock at the phase boundaries of its tiles
#ifdef WG_TC_TRACE
__device__ long long g_tc_trace[16 * 64];
#define WG_TC_STAMP(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0 && it < 64) g_tc_trace[it * 16 + (slot)] = clock64(); } while (0)
#else
#define WG_TC_STAMP(slot) do { } while (0)
#endif

constexpr int kTcTile = 128;          // envs per tile = MMA M = TMEM lanes
constexpr int kTcThreads = 256;       // two warps per TMEM lane quarter: each thread owns one env row and half of its columns
constexpr int kTcCols = 256;          // TMEM columns per CTA: D [0,64)  A_hi [64,136)  A_lo [144,216)
constexpr int kTcKH = 72;             // depth of layer 2: 64 hidden units + the bias column (a constant 1) + padding

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: element (r, k) of a [R x K] float matrix lives at byte
// (k / 4) * (R * 16) + r * 16 + (k % 4) * 4 from the start address, i.e. 16-byte chunks of 4 consecutive k, all R rows of a
// chunk contiguous: core matrices (8 rows x 16 B) are 128 B apart along the rows (stride byte offset) and R * 16 B apart
// along k (leading byte offset).  One MMA (K = 8) reads two chunks.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t rows) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);                 // start address            bits [0, 14)
    d |= (uint64_t)((rows * 16u) >> 4 & 0x3fffu) << 16;      // leading byte offset      bits [16, 30)
    d |= (uint64_t)(128u >> 4) << 32;                        // stride byte offset       bits [32, 46)
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    return d;                                                // base offset 0, layout type 0 = no swizzle
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t tc_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// bounded wait on an mbarrier phase; false = gave up (a fraction of a second)
__device__ __forceinline__ bool tc_wait_bar(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
#pragma unroll 1
    for (int it = 0; it < (1 << 18); it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
#ifdef WG_TC_WAIT_SLEEP
        __nanosleep(WG_TC_WAIT_SLEEP);          // a polling warp competes for issue slots with the warps that do the work
#endif
    }
    return false;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kTcMaxHeads = 9;        // M <= 8 action means + the value

// shared memory of one CTA (floats): B planes of the two layers (hi, lo), the head weights in float32, two staging buffers
// for raw observation tiles (row-major observations: one TMA bulk copy per tile, issued one tile ahead), the gaussian
// noise of the tile (float2 per env and action pair) and the column-half partial sums of the heads
template <int K1>
struct TcSmem {
    static constexpr int W1 = 64 * K1, W2 = 64 * kTcKH, WH = kTcMaxHeads * 64;
    static constexpr int o_w1 = 0, o_w2 = o_w1 + 2 * W1, o_wh = o_w2 + 2 * W2, o_zs = o_wh + WH + 4 /* pad to 16 B */;
    static constexpr int o_hx = o_zs + 4 * 2 * kTcTile, o_ls = o_hx + kTcMaxHeads * kTcTile, o_st = o_ls + 32;
    static constexpr int st_floats(int D) { return ((kTcTile * D + 3) / 4) * 4; }
    static constexpr int o_bar(int D) { return o_st + 2 * st_floats(D); }
    static constexpr size_t bytes(int D) { return sizeof(float) * o_bar(D) + 64; }          // + 3 mbarriers, TMEM slot
};

// One weight matrix [64 x K] into its hi (and lo) plane in the canonical layout, 16-byte chunk by chunk.  Element
// (n, k): w[n * ld + k] for k < k_valid; column k_valid is the BIAS column (the A operand carries a constant 1 there, so
// the tensor core adds the bias); zero beyond.  All of a thread's loads are issued before the first conversion (the
// 296 CTAs read the same few KB from L2 at the same time: one round trip instead of one per chunk).
template <bool SPLIT, int K>
__device__ __forceinline__ void tc_fill_b(float* hi, float* lo, int k_valid, const float* __restrict__ w, int ld,
                                          const float* __restrict__ bias) {
    constexpr int n_chunks = 64 * (K / 4), NPT = (n_chunks + kTcThreads - 1) / kTcThreads;
    float v[NPT][4];
#pragma unroll
    for (int i = 0; i < NPT; i++) {
        const int c = threadIdx.x + i * kTcThreads, n = c & 63, k4 = c >> 6;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = 4 * k4 + q;
            v[i][q] = 0.0f;
            if (c < n_chunks) { if (k < k_valid) v[i][q] = __ldg(w + n * ld + k); else if (k == k_valid) v[i][q] = __ldg(bias + n); }
        }
    }
#pragma unroll
    for (int i = 0; i < NPT; i++) {
        const int c = threadIdx.x + i * kTcThreads;
        if (c < n_chunks) {
            float4 h, l;
            h.x = __uint_as_float(to_tf32(v[i][0])); h.y = __uint_as_float(to_tf32(v[i][1]));
            h.z = __uint_as_float(to_tf32(v[i][2])); h.w = __uint_as_float(to_tf32(v[i][3]));
            l.x = v[i][0] - h.x; l.y = v[i][1] - h.y; l.z = v[i][2] - h.z; l.w = v[i][3] - h.w;
            reinterpret_cast<float4*>(hi)[c] = h;                  // chunk (n, k4) at (k4 * 64 + n) * 16 bytes
            if (SPLIT) reinterpret_cast<float4*>(lo)[c] = l;
        }
    }
}

// nan_to_num + clamp of the torch reference, then the hi / lo split of one A element
template <bool SPLIT>
__device__ __forceinline__ void tc_obs_elem(float v, float scale, float clip, uint32_t& hi, uint32_t& lo) {
    v = v * scale;
    v = (v != v) ? 0.0f : fminf(fmaxf(v, -clip), clip);
    act_split<SPLIT>(v, hi, lo);
}

// K1 = layer-1 depth: obs_dim + 1 (the bias column) rounded up to a multiple of 8; SPLIT = float32-grade 3xTF32
template <int K1, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 2)
policy_act_tc_kernel(const __grid_constant__ PolicyArgs A, int* __restrict__ error_flag) {
    using L = TcSmem<K1>;
    extern __shared__ __align__(128) float tsm[];
    const int D = A.D, M = A.M;
    float* const W1h = tsm + L::o_w1; float* const W1l = W1h + L::W1;
    float* const W2h = tsm + L::o_w2; float* const W2l = W2h + L::W2;
    float* const WH = tsm + L::o_wh;                                    // [M + 1][64] float32: rows < M means, row M value
    float2* const ZS = reinterpret_cast<float2*>(tsm + L::o_zs);        // [4 pairs][128 envs] gaussian noise of the tile
    float* const HX = tsm + L::o_hx;                                    // [9][128] head partial sums of column half 1
    float* const LS = tsm + L::o_ls;                                    // log_std[16], head biases[16]
    float* const HB = LS + 16;
    float* const ST0 = tsm + L::o_st;                                   // two raw observation tiles [128][D]
    const int st_floats = L::st_floats(D);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(tsm + L::o_bar(D));   // [0] MMA commits, [1], [2] staging buffers full
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // warp-uniform for the compiler too (MMA issue below)
    const int row = tid & (kTcTile - 1), half = warp >> 2;             // env row of the tile; which half of the columns
    const int64_t E = A.E;
    const int64_t n_tiles = (E + kTcTile - 1) / kTcTile;
    // row-major observations whose tiles are whole and 16-byte aligned arrive by TMA bulk copy, one tile ahead
    const uint32_t tile_bytes = (uint32_t)(kTcTile * D * 4);
    const bool tma_ok = A.obs_layout == 0 && ((reinterpret_cast<uintptr_t>(A.obs) & 15u) == 0);
    auto tile_by_tma = [&](int64_t t) { return tma_ok && (t + 1) * kTcTile <= E; };

    if (tid == 0) {
        for (int i = 0; i < 3; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int64_t t0 = blockIdx.x;                                  // the first tile's observations: in flight during the setup
        if (t0 < n_tiles && tile_by_tma(t0)) tc_bulk_g2s(ST0, A.obs + t0 * kTcTile * D, tile_bytes, bar + 1);
    }
    // ---- one-time setup: weights with their bias columns (hi / lo planes), head weights, tensor memory ----
    tc_fill_b<SPLIT, K1>(W1h, W1l, D, A.w1, D, A.b1);
    tc_fill_b<SPLIT, kTcKH>(W2h, W2l, 64, A.w2, 64, A.b2);
    for (int i = tid; i < (M + 1) * 64; i += kTcThreads) WH[i] = i < M * 64 ? __ldg(A.w_mu + i) : __ldg(A.w_v + (i - M * 64));
    if (tid < 16) { LS[tid] = tid < M ? A.log_std[tid] : 0.0f; HB[tid] = tid < M ? A.b_mu[tid] : (tid == M ? A.b_v[0] : 0.0f); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTcCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the weight planes -> visible to the tensor core's reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;                                  // lane 0, first column of this CTA's allocation
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's quarter of the lanes
    constexpr uint32_t cD = 0, cAh = 64, cAl = 144;
    uint32_t phase = 0, st_phase0 = 0, st_phase1 = 0;
    bool alive = true;

    const uint32_t step = A.step_index + (A.step_counter ? __ldg(A.step_counter) : 0u);
    const uint64_t dW1h = tc_smem_desc(smem_u32(W1h), 64), dW1l = tc_smem_desc(smem_u32(W1l), 64);
    const uint64_t dW2h = tc_smem_desc(smem_u32(W2h), 64), dW2l = tc_smem_desc(smem_u32(W2l), 64);
    // a k-step of 8 advances a B operand by two 16-byte chunks = 2 * 64 * 16 bytes (descriptor units of 16 bytes)
    constexpr uint32_t kStepB = 2 * 64;
    constexpr uint32_t id64 = tc_idesc(64);
    constexpr int NC = K1 / 8, NC0 = (NC + 1) / 2;                     // obs chunks of 8 columns: half 0 takes [0, NC0)
    const int n_pairs = A.sample ? (M + 1) / 2 : 0;
    const bool even_d = (D & 1) == 0;                                  // staged rows are then 8-byte aligned: 64-bit reads

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles && alive; tile += gridDim.x, it++) {
        const int buf = it & 1;
        WG_TC_STAMP(0);
        const int64_t e = tile * kTcTile + row;
        const bool ev = e < E;
        const int64_t nxt = tile + gridDim.x;
        // ---- this env's observation row (this thread's half of it) -> sanitise -> hi / lo planes of the layer-1 A
        // operand in tensor memory; column D carries the constant 1 that multiplies the bias column of W1 ----
        const bool staged = tile_by_tma(tile);
        if (staged) {
            alive = tc_wait_bar(bar + 1 + buf, buf ? st_phase1 : st_phase0);
            if (buf) st_phase1 ^= 1; else st_phase0 ^= 1;
        }
        WG_TC_STAMP(1);
        const float* srow = ST0 + buf * st_floats + row * D;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            if ((j < NC0) != (half == 0)) continue;                 // warp-uniform
            uint32_t hi[8], lo[8];
            if (8 * j + 8 <= D) {                                   // a whole chunk of observation entries (warp-uniform)
                float x[8];
                if (staged) {
                    if (even_d) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float2 t = reinterpret_cast<const float2*>(srow + 8 * j)[q];
                            x[2 * q] = t.x; x[2 * q + 1] = t.y;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; q++) x[q] = srow[8 * j + q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        x[q] = ev ? __ldg(A.obs + (A.obs_layout ? (int64_t)(8 * j + q) * E + e : e * D + 8 * j + q)) : 0.0f;
                }
#pragma unroll
                for (int q = 0; q < 8; q++) tc_obs_elem<SPLIT>(x[q], A.obs_scale, A.obs_clip, hi[q], lo[q]);
            } else {                                                // the chunk with the end of the row, the 1 and the padding
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int f = 8 * j + q;
                    float x = 0.0f;
                    if (f < D) {
                        if (staged) x = srow[f];
                        else if (ev) x = __ldg(A.obs + (A.obs_layout ? (int64_t)f * E + e : e * D + f));
                    }
                    tc_obs_elem<SPLIT>(x, A.obs_scale, A.obs_clip, hi[q], lo[q]);
                    if (f >= D) { hi[q] = f == D ? 0x3f800000u : 0u; lo[q] = 0u; }
                }
            }
            tc_st8(t_lane + cAh + 8 * j, hi);
            if (SPLIT) tc_st8(t_lane + cAl + 8 * j, lo);
        }
        tc_wait_st();
        tc_fence_before();
        WG_TC_STAMP(2);
        alive = __syncthreads_and(alive);                          // also: every thread is done with staging[buf]
        WG_TC_STAMP(3);
        // ---- layer 1: D[0,64) = [obs, 1] * [W1, b1]^T (A from tensor memory); issued by one elected lane of warp 0 in
        // warp-uniform control flow (straight-line UTCHMMA, no lane-election loop around each instruction) ----
        if (warp == 0) {
            if (alive && tc_elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < NC; kk++) {
                    tc_mma_ts(tmem + cD, tmem + cAh + 8 * kk, dW1h + kk * kStepB, id64, kk > 0);
                    if (SPLIT) {
                        tc_mma_ts(tmem + cD, tmem + cAl + 8 * kk, dW1h + kk * kStepB, id64, 1);
                        tc_mma_ts(tmem + cD, tmem + cAh + 8 * kk, dW1l + kk * kStepB, id64, 1);
                    }
                }
                tc_commit(bar);
                // the next tile's observations: its staging buffer was last read one tile ago (several barriers since)
                if (nxt < n_tiles && tile_by_tma(nxt))
                    tc_bulk_g2s(ST0 + (buf ^ 1) * st_floats, A.obs + nxt * kTcTile * D, tile_bytes, bar + 1 + (buf ^ 1));
            }
            __syncwarp();
        }
        WG_TC_STAMP(4);
        if (alive) alive = tc_wait_bar(bar, phase);
        WG_TC_STAMP(5);
        alive = __syncthreads_and(alive);                          // uniform verdict: nobody waits at a barrier others left
        phase ^= 1;
        tc_fence_after();
        if (!alive) break;
        // ---- epilogue 1: tanh(D) -> hi / lo planes of layer 2's A operand, in tensor memory; this thread owns 32 of its
        // env's 64 hidden units; half 1 also writes the chunk with the constant 1 of the bias column ----
        {
            uint32_t a[16], b[16];
            tc_ld16(t_lane + cD + 32 * half, a);
            tc_ld16(t_lane + cD + 32 * half + 16, b);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float y = pol_tanh<SPLIT>(__uint_as_float(c ? b[i] : a[i]));
                    act_split<SPLIT>(y, hi[i], lo[i]);
                }
                tc_st16(t_lane + cAh + 32 * half + 16 * c, hi);
                if (SPLIT) tc_st16(t_lane + cAl + 32 * half + 16 * c, lo);
            }
            if (half) {
                const uint32_t one[8] = { 0x3f800000u, 0u, 0u, 0u, 0u, 0u, 0u, 0u }, zero[8] = { 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u };
                tc_st8(t_lane + cAh + 64, one);
                if (SPLIT) tc_st8(t_lane + cAl + 64, zero);
            }
        }
        WG_TC_STAMP(6);
        tc_wait_st();
        tc_fence_before();
        __syncthreads();
        WG_TC_STAMP(7);
        // ---- layer 2: D[0,64) = [tanh(h1), 1] * [W2, b2]^T ----
        if (warp == 0) {
            if (tc_elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < kTcKH / 8; kk++) {
                    tc_mma_ts(tmem + cD, tmem + cAh + 8 * kk, dW2h + kk * kStepB, id64, kk > 0);
                    if (SPLIT) {
                        tc_mma_ts(tmem + cD, tmem + cAl + 8 * kk, dW2h + kk * kStepB, id64, 1);
                        tc_mma_ts(tmem + cD, tmem + cAh + 8 * kk, dW2l + kk * kStepB, id64, 1);
                    }
                }
                tc_commit(bar);
            }
            __syncwarp();
        }
        WG_TC_STAMP(8);
        // ---- while the tensor core works: the tile's gaussian noise (Philox + Box-Muller per env and action pair); pair p
        // is evaluated by the thread of column half (p + 1) & 1 and handed to the sampling thread through shared memory ----
#pragma unroll
        for (int pr = 0; pr < 4; pr++)
            if (pr < n_pairs && half == ((pr + 1) & 1) && ev)
                ZS[pr * kTcTile + row] = pol_normal2(A.seed_lo, A.seed_hi, A.env_offset + (uint32_t)e, step, (uint32_t)pr);
        alive = __syncthreads_and(tc_wait_bar(bar, phase));
        WG_TC_STAMP(9);
        phase ^= 1;
        tc_fence_after();
        if (!alive) break;
        // ---- epilogue 2 + heads: y = tanh(D); output n = b_n + sum_k y_k W_n[k] in float32 on the CUDA cores, this
        // thread's 32 columns first, the two halves of an env added through shared memory ----
        float hs[kTcMaxHeads];
        {
            uint32_t a[16], b[16];
            tc_ld16(t_lane + cD + 32 * half, a);
            tc_ld16(t_lane + cD + 32 * half + 16, b);
            tc_wait_ld();
            float y[32];
#pragma unroll
            for (int i = 0; i < 16; i++) { y[i] = pol_tanh<SPLIT>(__uint_as_float(a[i])); y[16 + i] = pol_tanh<SPLIT>(__uint_as_float(b[i])); }
            WG_TC_STAMP(13);
#pragma unroll
            for (int n = 0; n < kTcMaxHeads; n++) {
                hs[n] = 0.0f;
                if (n <= M) {                                       // warp-uniform
                    const float4* wn = reinterpret_cast<const float4*>(WH + n * 64 + 32 * half);
                    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                    for (int q = 0; q < 8; q += 2) {
                        const float4 w0 = wn[q], w1 = wn[q + 1];
                        s0 = __fmaf_rn(y[4 * q + 0], w0.x, s0); s0 = __fmaf_rn(y[4 * q + 1], w0.y, s0);
                        s0 = __fmaf_rn(y[4 * q + 2], w0.z, s0); s0 = __fmaf_rn(y[4 * q + 3], w0.w, s0);
                        s1 = __fmaf_rn(y[4 * q + 4], w1.x, s1); s1 = __fmaf_rn(y[4 * q + 5], w1.y, s1);
                        s1 = __fmaf_rn(y[4 * q + 6], w1.z, s1); s1 = __fmaf_rn(y[4 * q + 7], w1.w, s1);
                    }
                    hs[n] = s0 + s1;
                    if (half) HX[n * kTcTile + row] = hs[n];
                }
            }
        }
        tc_fence_before();                      // the next tile's MMAs overwrite D only after this thread's loads (and barriers)
        WG_TC_STAMP(10);
        __syncthreads();
        WG_TC_STAMP(11);
        // ---- heads (one thread per env): outputs n < M means, n == M value; gaussian sample, log-prob ----
        if (half == 0 && ev) {
            float lp = 0.0f, mean[8], act[8], val = 0.0f;
#pragma unroll
            for (int pr = 0; pr < 4; pr++) {
                if (2 * pr < M) {
                    float2 z = make_float2(0.0f, 0.0f);
                    if (A.sample) z = ZS[pr * kTcTile + row];
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const int n = 2 * pr + j;
                        if (n < M) {
                            const float ls = LS[n], eps = j ? z.y : z.x;
                            mean[n] = hs[n] + HX[n * kTcTile + row] + HB[n];
                            act[n] = A.sample ? __fmaf_rn(__expf(ls), eps, mean[n]) : mean[n];
                            lp += -0.5f * eps * eps - ls - 0.9189385332046727f;
                        }
                    }
                }
            }
#pragma unroll
            for (int n = 0; n < kTcMaxHeads; n++) if (n == M) val = hs[n] + HX[n * kTcTile + row] + HB[n];
            WG_TC_STAMP(12);
#pragma unroll
            for (int n = 0; n < 8; n++) {
                if (n < M) {
                    if (A.mean) A.mean[(int64_t)n * E + e] = mean[n];
                    if (A.action) A.action[A.act_layout ? (int64_t)n * E + e : e * M + n] = act[n];
                }
            }
            if (A.value) A.value[e] = val;
            if (A.logp) A.logp[e] = lp;
        }
        WG_TC_STAMP(14);
    }
    if (!alive && error_flag && tid == 0) atomicExch(error_flag, 1);
    // ---- teardown: the allocating warp frees the tensor memory once every warp is done with it ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTcCols) : "memory");
}

}  // namespace wg
