// wg_policy_tc.cuh -- the rollout policy (BASELINE config 5) on the 5th-generation tensor cores.
//
// Same function as policy_act_kernel (wg_policy.cuh): obs -> gaussian MLP (D -> 64 -> 64 -> {M means, 1 value}, tanh)
// -> sampled action, log-prob, value, one launch per env step, reading the observation as the step kernel wrote it.
// The three GEMMs run as tcgen05.mma (kind::tf32, M = 128 envs per tile, accumulators in tensor memory):
//
//   layer 1   A = the tile's observations [128 x K1] in TENSOR MEMORY: thread r owns TMEM lane r = env r, sanitises its
//             observation row and writes it with tcgen05.st; B = W1 [64 x K1] in shared memory (canonical K-major
//             layout, no swizzle); D -> TMEM columns [0, 64).  Row-major observation tiles (one contiguous 128 x D
//             block) arrive by TMA bulk copy into a double-buffered staging area, one tile ahead.
//   layer 2   A = tanh(layer 1) [128 x 72] in tensor memory (written there by the epilogue: an activation row never
//             touches shared memory), B = [W2, b2] [64 x 72] in smem, D -> TMEM columns [0, 64)
//   heads     A = tanh(layer 2) in TMEM, B = [w_mu, b_mu; w_v, b_v; 0] [16 x 72] in smem, D -> TMEM columns [224, 240)
// The biases ride in the GEMMs: every A operand carries a constant 1 in the column after its last feature and every B
// operand its bias there, so the epilogues are tanh + split only.  256 threads per tile: two warps share each quarter
// of the TMEM lanes and split an env row's columns between them.
//
// With SPLIT every product is the 3-term error-compensated a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (float32-grade, like the
// mma.sync kernel): three MMAs per k-step into the same accumulator, hi / lo planes of A in TMEM columns [64, 144) /
// [144, 224) and of B in shared memory.  One thread issues the MMAs of a layer and commits them to an mbarrier; all
// 128 threads then run that layer's epilogue (tcgen05.ld -> bias, tanh, split -> tcgen05.st).  Two CTAs per SM
// (2 x 256 TMEM columns, 2 x ~100 KB of shared memory) overlap one tile's epilogue with the other's MMAs.
// Every mbarrier wait is bounded: a descriptor mistake ends the kernel with an error flag instead of hanging the GPU.
#pragma once
#include "wg_policy.cuh"

namespace wg {

constexpr int kTcTile = 128;          // envs per tile = MMA M = TMEM lanes
constexpr int kTcThreads = 256;       // two warps per TMEM lane quarter: each thread owns one env row and half of its columns
constexpr int kTcCols = 256;          // TMEM columns per CTA: D [0,64)  A_hi [64,144)  A_lo [144,224)  heads [224,240)
constexpr int kTcHeadN = 16;          // heads MMA N (smallest N for M = 128); rows 0..M-1 means, row M value, rest 0
constexpr int kTcKH = 72;             // depth of layer 2 / the heads: 64 hidden units + the bias column (a constant 1) + padding

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
    for (int it = 0; it < (1 << 18); it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
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

// shared memory of one CTA (floats): B planes of the three layers (hi, lo) and two staging buffers for raw observation
// tiles (row-major observations: one TMA bulk copy per tile, issued one tile ahead)
template <int K1>
struct TcSmem {
    static constexpr int W1 = 64 * K1, W2 = 64 * kTcKH, WH = kTcHeadN * kTcKH;
    static constexpr int o_w1 = 0, o_w2 = o_w1 + 2 * W1, o_wh = o_w2 + 2 * W2, o_st = o_wh + 2 * WH;
    static constexpr int st_floats(int D) { return ((kTcTile * D + 3) / 4) * 4; }
    static constexpr int o_ls(int D) { return o_st + 2 * st_floats(D); }                    // log_std[16]
    static constexpr size_t bytes(int D) { return sizeof(float) * (o_ls(D) + 16) + 64; }    // + 3 mbarriers, TMEM slot
};

// One weight matrix [rows x K] into its hi (and lo) plane in the canonical layout, 16-byte chunk by chunk.  Element
// (n, k): w[n * ld + k] for k < k_valid (row n_valid from w_last, rows beyond: zero); column k_valid is the BIAS column
// (bias[n], or bias_last[0] for row n_valid): the A operand carries a constant 1 there, so the tensor core adds the bias.
template <bool SPLIT>
__device__ __forceinline__ void tc_fill_b(float* hi, float* lo, int rows, int K, int n_valid, int k_valid,
                                          const float* __restrict__ w, int ld, const float* __restrict__ w_last,
                                          const float* __restrict__ bias, const float* __restrict__ bias_last) {
    const int n_chunks = rows * (K / 4);
    for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
        const int n = c % rows, k4 = c / rows;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = 4 * k4 + q;
            v[q] = 0.0f;
            if (n < n_valid) { if (k < k_valid) v[q] = __ldg(w + n * ld + k); else if (k == k_valid) v[q] = __ldg(bias + n); }
            else if (n == n_valid && w_last) { if (k < k_valid) v[q] = __ldg(w_last + k); else if (k == k_valid) v[q] = __ldg(bias_last); }
        }
        float4 h, l;
        h.x = __uint_as_float(to_tf32(v[0])); h.y = __uint_as_float(to_tf32(v[1]));
        h.z = __uint_as_float(to_tf32(v[2])); h.w = __uint_as_float(to_tf32(v[3]));
        l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
        reinterpret_cast<float4*>(hi)[k4 * rows + n] = h;          // chunk (n, k4) at (k4 * rows + n) * 16 bytes
        if (SPLIT) reinterpret_cast<float4*>(lo)[k4 * rows + n] = l;
    }
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
// nan_to_num + clamp of the torch reference, then the hi / lo split of one A element
template <bool SPLIT>
__device__ __forceinline__ void tc_obs_elem(float v, float scale, float clip, uint32_t& hi, uint32_t& lo) {
    v = v * scale;
    v = (v != v) ? 0.0f : fminf(fmaxf(v, -clip), clip);
    hi = to_tf32(v);
    lo = SPLIT ? __float_as_uint(v - __uint_as_float(hi)) : 0u;
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
    float* const WHh = tsm + L::o_wh; float* const WHl = WHh + L::WH;
    float* const ST0 = tsm + L::o_st;                                   // two raw observation tiles [128][D]
    const int st_floats = L::st_floats(D);
    float* const LS = tsm + L::o_ls(D);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(LS + 16);        // [0] MMA commits, [1], [2] staging buffers full
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (kTcTile - 1), half = tid >> 7;              // env row of the tile; which half of the columns
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
    // ---- one-time setup: weights with their bias columns (hi / lo planes), tensor memory ----
    tc_fill_b<SPLIT>(W1h, W1l, 64, K1, 64, D, A.w1, D, nullptr, A.b1, nullptr);
    tc_fill_b<SPLIT>(W2h, W2l, 64, kTcKH, 64, 64, A.w2, 64, nullptr, A.b2, nullptr);
    tc_fill_b<SPLIT>(WHh, WHl, kTcHeadN, kTcKH, M, 64, A.w_mu, 64, A.w_v, A.b_mu, A.b_v);
    if (tid < 16) LS[tid] = tid < M ? A.log_std[tid] : 0.0f;
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
    constexpr uint32_t cD = 0, cAh = 64, cAl = 144, cH = 224;
    uint32_t phase = 0, st_phase0 = 0, st_phase1 = 0;
    bool alive = true;

    const uint32_t step = A.step_index + (A.step_counter ? __ldg(A.step_counter) : 0u);
    const uint64_t dW1h = tc_smem_desc(smem_u32(W1h), 64), dW1l = tc_smem_desc(smem_u32(W1l), 64);
    const uint64_t dW2h = tc_smem_desc(smem_u32(W2h), 64), dW2l = tc_smem_desc(smem_u32(W2l), 64);
    const uint64_t dWHh = tc_smem_desc(smem_u32(WHh), kTcHeadN), dWHl = tc_smem_desc(smem_u32(WHl), kTcHeadN);
    // a k-step of 8 advances a B operand by two 16-byte chunks = 2 * rows * 16 bytes (descriptor units of 16 bytes)
    constexpr uint32_t kStepB64 = 2 * 64, kStepBH = 2 * kTcHeadN;
    constexpr uint32_t id64 = tc_idesc(64), idH = tc_idesc(kTcHeadN);
    constexpr int NC = K1 / 8, NC0 = (NC + 1) / 2;                     // obs chunks of 8 columns: half 0 takes [0, NC0)

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles && alive; tile += gridDim.x, it++) {
        const int buf = it & 1;
        const int64_t e = tile * kTcTile + row;
        const bool ev = e < E;
        // the next tile's observations: its staging buffer was last read one tile ago (a __syncthreads since)
        const int64_t nxt = tile + gridDim.x;
        if (tid == 0 && nxt < n_tiles && tile_by_tma(nxt))
            tc_bulk_g2s(ST0 + (buf ^ 1) * st_floats, A.obs + nxt * kTcTile * D, tile_bytes, bar + 1 + (buf ^ 1));
        // ---- this env's observation row (this thread's half of it) -> sanitise -> hi / lo planes of the layer-1 A
        // operand in tensor memory; column D carries the constant 1 that multiplies the bias column of W1 ----
        const bool staged = tile_by_tma(tile);
        if (staged) {
            alive = tc_wait_bar(bar + 1 + buf, buf ? st_phase1 : st_phase0);
            if (buf) st_phase1 ^= 1; else st_phase0 ^= 1;
        }
        const float* srow = ST0 + buf * st_floats + row * D;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            if ((j < NC0) != (half == 0)) continue;                 // warp-uniform
            uint32_t hi[8], lo[8];
            if (8 * j + 8 <= D) {                                   // a whole chunk of observation entries (warp-uniform)
                float x[8];
                if (staged) {
#pragma unroll
                    for (int q = 0; q < 8; q++) x[q] = srow[8 * j + q];
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
        alive = __syncthreads_and(alive);                          // also: every thread is done with staging[buf]
        // ---- layer 1: D[0,64) = [obs, 1] * [W1, b1]^T (A from tensor memory) ----
        if (tid == 0 && alive) {
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < NC; kk++) {
                tc_mma_ts(tmem + cD, tmem + cAh + 8 * kk, dW1h + kk * kStepB64, id64, kk > 0);
                if (SPLIT) {
                    tc_mma_ts(tmem + cD, tmem + cAl + 8 * kk, dW1h + kk * kStepB64, id64, 1);
                    tc_mma_ts(tmem + cD, tmem + cAh + 8 * kk, dW1l + kk * kStepB64, id64, 1);
                }
            }
            tc_commit(bar);
        }
        if (alive) alive = tc_wait_bar(bar, phase);
        alive = __syncthreads_and(alive);                          // uniform verdict: nobody waits at a barrier others left
        phase ^= 1;
        tc_fence_after();
        // ---- epilogues 1 and 2: tanh(D) -> hi / lo planes of the next layer's A operand, in tensor memory; this thread
        // owns 32 of its env's 64 hidden units; half 1 also writes the chunk with the constant 1 of the bias column ----
#pragma unroll 1
        for (int layer = 0; layer < 2 && alive; layer++) {
            uint32_t v[32];
            {
                uint32_t a[16], b[16];
                tc_ld16(t_lane + cD + 32 * half, a);
                tc_ld16(t_lane + cD + 32 * half + 16, b);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i++) { v[i] = a[i]; v[16 + i] = b[i]; }
            }
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float y = pol_tanh<SPLIT>(__uint_as_float(v[16 * c + i]));
                    hi[i] = to_tf32(y);
                    lo[i] = SPLIT ? __float_as_uint(y - __uint_as_float(hi[i])) : 0u;
                }
                tc_st16(t_lane + cAh + 32 * half + 16 * c, hi);
                if (SPLIT) tc_st16(t_lane + cAl + 32 * half + 16 * c, lo);
            }
            if (half) {
                const uint32_t one[8] = { 0x3f800000u, 0u, 0u, 0u, 0u, 0u, 0u, 0u }, zero[8] = { 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u };
                tc_st8(t_lane + cAh + 64, one);
                if (SPLIT) tc_st8(t_lane + cAl + 64, zero);
            }
            tc_wait_st();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t dcol = layer ? cH : cD;
                const uint64_t bh = layer ? dWHh : dW2h, bl = layer ? dWHl : dW2l;
                const uint32_t ks = layer ? kStepBH : kStepB64, id = layer ? idH : id64;
#pragma unroll
                for (int kk = 0; kk < kTcKH / 8; kk++) {
                    tc_mma_ts(tmem + dcol, tmem + cAh + 8 * kk, bh + kk * ks, id, kk > 0);
                    if (SPLIT) {
                        tc_mma_ts(tmem + dcol, tmem + cAl + 8 * kk, bh + kk * ks, id, 1);
                        tc_mma_ts(tmem + dcol, tmem + cAh + 8 * kk, bl + kk * ks, id, 1);
                    }
                }
                tc_commit(bar);
            }
            alive = __syncthreads_and(tc_wait_bar(bar, phase));
            phase ^= 1;
            tc_fence_after();
        }
        if (!alive) break;
        // ---- heads (one thread per env): outputs n < M means, n == M value (biases included); gaussian sample, log-prob ----
        if (half == 0) {
            uint32_t hv[16];
            tc_ld16(t_lane + cH, hv);
            tc_wait_ld();
            if (ev) {
                float lp = 0.0f;
#pragma unroll
                for (int pr = 0; pr < 4; pr++) {
                    if (2 * pr < M) {
                        float2 z = make_float2(0.0f, 0.0f);
                        if (A.sample) z = pol_normal2(A.seed_lo, A.seed_hi, A.env_offset + (uint32_t)e, step, (uint32_t)pr);
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const int n = 2 * pr + j;
                            if (n < M) {
                                const float mean = __uint_as_float(hv[n]), ls = LS[n], eps = j ? z.y : z.x;
                                const float act = A.sample ? __fmaf_rn(__expf(ls), eps, mean) : mean;
                                lp += -0.5f * eps * eps - ls - 0.9189385332046727f;
                                if (A.mean) A.mean[(int64_t)n * E + e] = mean;
                                if (A.action) A.action[A.act_layout ? (int64_t)n * E + e : e * M + n] = act;
                            }
                        }
                    }
                }
                if (A.value) {
                    float val = 0.0f;
#pragma unroll
                    for (int n = 0; n < 8; n++) if (n == M) val = __uint_as_float(hv[n]);
                    A.value[e] = val;
                }
                if (A.logp) A.logp[e] = lp;
            }
        }
        tc_fence_before();                      // the next tile's MMAs overwrite the heads' columns only after further barriers
    }
    if (!alive && error_flag && tid == 0) atomicExch(error_flag, 1);
    // ---- teardown: the allocating warp frees the tensor memory once every warp is done with it ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTcCols) : "memory");
}

}  // namespace wg
