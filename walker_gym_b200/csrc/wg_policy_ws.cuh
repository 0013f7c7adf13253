// wg_policy_ws.cuh -- the rollout policy (BASELINE config 5) as a WARP-SPECIALISED tcgen05 pipeline: one persistent CTA
// per SM, 25 warps in five roles, three to four tiles (128 envs each) in flight in different stages, all hand-offs through
// mbarriers, tensor memory (all 512 columns of the SM) as the only inter-stage storage for activations.
//
//   role            warps    per tile
//   P   producers   0..3     observation row of env r (TMA-staged in shared memory) -> nan_to_num / clip -> hi / lo TF32
//                            planes of the layer-1 A operand in tensor memory (tcgen05.st); thread 0 issues the TMA bulk
//                            copy of the tile two ahead
//   E1  epilogue 1  4..11    D1 (tcgen05.ld) -> tanh -> hi / lo planes of the layer-2 A operand (tcgen05.st)
//   E2  epilogue 2  12..19   D2 -> tanh -> heads (M + 1 <= 9 outputs of depth 64) as float32 FMAs against the head weights
//                            in shared memory; each thread's 32-column partial sums go to shared memory (double buffered)
//   O   output      20..23   one thread per env: Philox + Box-Muller noise of the first action pair (evaluated while the
//                            tile is still upstream), heads = the two halves + bias -> gaussian sample, log-prob, value
//                            -> global memory
//   MMA             24       one elected lane: layer 1 of tile i, then layer 2 of tile i-1 (tcgen05.mma kind::tf32, M 128,
//                            N 64, A from tensor memory, B = weight planes in shared memory, biases folded in as an extra
//                            k-column); tcgen05.commit -> the consumers' "full" and the producers' "free" barriers
//
// Tensor-memory columns: obs planes [0, 2 K1) | D1 x 2 | layer-2 A planes 2 x 72 | D2 x 2 (x 1 when K1 > 48).  D1 and D2
// are double buffered so that the tensor core runs one tile ahead of each epilogue; the layer-2 A planes are single
// buffered (E1 computes a column chunk in registers and waits for the previous tile's layer-2 MMAs only before it
// stores).  Why this shape (measured on B200): the monolithic kernel (wg_policy_tc.cuh: every
// thread walks obs -> MMA -> tanh -> MMA -> tanh -> heads, two CTAs per SM) is latency bound -- 27 % issue-slot
// utilisation, each phase waiting for the previous one -- while no single resource is busy more than a third of the
// time (XU pipe 2100 cycles per tile for the float32-grade tanh, tensor pipe 1350, instruction issue 2000, against
// 7800 cycles per tile achieved).  With the stages decoupled the slowest resource sets the pace -- provided the code of
// the five roles fits the instruction caches together: loops over heads, action pairs and outputs are ROLLED on purpose
// (the first, fully unrolled version had 6600 SASS instructions and lost a quarter to half of every role's issue slots to
// instruction fetch).
//
// With SPLIT every product is the 3-term error-compensated a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (float32-grade) and the
// factor 2 log2(e) of tanh = 1 - 2 / (2^(2 log2(e) x) + 1) is folded into weights and biases; without, plain TF32 (hi planes
// only) and tanh.approx.  Every mbarrier wait is bounded; a role that gives up raises the error flag and stops touching
// memory (the producers keep their named barrier company until their loop ends).
#pragma once
#include "wg_policy_tc.cuh"
#include "wg_policy_step.cuh"

namespace wg {

// development aid (profiles/microbench/ws_trace.cu): one thread per role of CTA 0 stamps the clock at its phase boundaries
#ifdef WG_WS_TRACE
__device__ long long g_ws_trace[4 * 16 * 8];
#define WG_WS_STAMP(role, who, slot) do { if (blockIdx.x == 0 && (who) && i < 16) g_ws_trace[((role) * 16 + i) * 8 + (slot)] = clock64(); } while (0)
#else
#define WG_WS_STAMP(role, who, slot) do { } while (0)
#endif

constexpr int kWsWarpsP = 4, kWsWarpsE = 8, kWsWarpsO = 4;
constexpr int kWsThreads = 32 * (kWsWarpsP + 2 * kWsWarpsE + kWsWarpsO + 1);      // 800 (the policy alone)
constexpr int kWsWarpE1 = kWsWarpsP, kWsWarpE2 = kWsWarpE1 + kWsWarpsE, kWsWarpO = kWsWarpE2 + kWsWarpsE;
// the fused kernel (wg_policy_step.cuh) has TWO groups of output warps that take alternate tiles: the env step they carry
// is ~1400 instructions per env, too long for one warp per scheduler to keep the pipeline's pace
template <class SA> struct WsCfg {
    static constexpr int kGroupsO = SA::kFused ? 2 : 1;
    static constexpr int kWarpMma = kWsWarpO + kWsWarpsO * kGroupsO;
    static constexpr int kThreads = 32 * (kWarpMma + 1);
};

// barriers (uint64_t each)
enum { kBarObsFull = 0 /* x2 */, kBarA1Ready = 2, kBarA1Free = 3, kBarD1Full = 4 /* x2 */, kBarE1Done = 6, kBarHFree = 7,
       kBarD2Full = 8 /* x2 */, kBarD2Free = 10 /* x2 */, kBarHxFull = 12 /* x2 */, kBarHxFree = 14 /* x2 */, kBarCount = 16 };

template <int K1, int OBS_OUT_FLOATS = 0>
struct WsSmem {
    static constexpr int W1 = 64 * K1, W2 = 64 * kTcKH, WH = kTcMaxHeads * 64;
    static constexpr int o_w1 = 0, o_w2 = o_w1 + 2 * W1, o_wh = o_w2 + 2 * W2, o_hx = o_wh + WH + 4;
    static constexpr int o_ls = o_hx + 2 * 2 * kTcMaxHeads * kTcTile, o_ot = o_ls + 32;      // o_ot: fused step, next observations
    static constexpr int o_st = o_ot + ((OBS_OUT_FLOATS + 3) / 4) * 4;
    static constexpr int st_floats(int D) { return ((kTcTile * D + 3) / 4) * 4 + 8; }     // + 8: the last row's tail chunk reads past its end
    static constexpr int o_bar(int D) { return o_st + 2 * st_floats(D); }
    static constexpr size_t bytes(int D) { return sizeof(float) * o_bar(D) + 8 * kBarCount + 16; }
};

__device__ __forceinline__ void ws_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_group_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

// One weight matrix [64 x k_valid] (row-major, leading dimension k_valid) plus its bias into the hi (and lo) plane of the
// canonical K-major layout: element (n, k) at float index ((k / 4) * 64 + n) * 4 + (k % 4); column k_valid = bias.  The
// global reads are coalesced and all in flight at once; the planes were zeroed before (padding columns).
// `pre` scales weights and bias: the float32-grade tanh is 1 - 2 / (2^(x * 2 log2 e) + 1), and the constant factor rides in
// the GEMM instead of costing a multiplication per activation.
template <bool SPLIT, int K, int NT>
__device__ __forceinline__ void ws_fill_b(float* hi, float* lo, int k_valid, const float* __restrict__ w,
                                          const float* __restrict__ bias, float pre) {
    constexpr int NPT = (64 * K + NT - 1) / NT;
    const int total = 64 * k_valid;
    float v[NPT];
#pragma unroll
    for (int j = 0; j < NPT; j++) {
        const int i = threadIdx.x + j * NT;
        v[j] = i < total ? __ldg(w + i) * pre : 0.0f;
    }
    const float bv = threadIdx.x < 64 ? __ldg(bias + threadIdx.x) * pre : 0.0f;
#pragma unroll
    for (int j = 0; j < NPT; j++) {
        const int i = threadIdx.x + j * NT;
        if (i < total) {
            const int n = i / k_valid, k = i - n * k_valid, idx = ((k >> 2) * 64 + n) * 4 + (k & 3);
            const float h = __uint_as_float(to_tf32(v[j]));
            hi[idx] = h;
            if (SPLIT) lo[idx] = v[j] - h;
        }
    }
    if (threadIdx.x < 64) {
        const int idx = ((k_valid >> 2) * 64 + threadIdx.x) * 4 + (k_valid & 3);
        const float h = __uint_as_float(to_tf32(bv));
        hi[idx] = h;
        if (SPLIT) lo[idx] = bv - h;
    }
}

// tanh of a pre-activation that the GEMM already scaled by 2 log2(e) (SPLIT) / left alone (plain TF32: tanh.approx)
template <bool SPLIT>
__device__ __forceinline__ float ws_tanh(float x) {
    if (!SPLIT) return pol_tanh<false>(x);
    float e, r;                            // exp -> inf gives 1 - 2 / inf = 1, exp -> 0 gives 1 - 2 = -1: no clamp needed
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return __fmaf_rn(-2.0f, r, 1.0f);
}

// K1 = layer-1 depth: obs_dim + 1 (the bias column) rounded up to a multiple of 8; SPLIT = float32-grade 3xTF32;
// SA = NoStep (the policy alone) or FusedStep<...> (the output warps also run the env step, wg_policy_step.cuh)
template <int K1, bool SPLIT, class SA = NoStep>
__global__ void __launch_bounds__(WsCfg<SA>::kThreads, 1)
policy_act_ws_kernel(const __grid_constant__ PolicyArgs A, int* __restrict__ error_flag, const __grid_constant__ SA S) {
    using L = WsSmem<K1, SA::kObsFloats * WsCfg<SA>::kGroupsO>;
    constexpr int NT = WsCfg<SA>::kThreads, kWarpMma = WsCfg<SA>::kWarpMma;
    constexpr int ND2 = (2 * K1 + 128 + 2 * kTcKH + 128 <= 512) ? 2 : 1;         // D2 buffers that fit next to the rest
    constexpr uint32_t cOh = 0, cOl = K1, cD1 = 2 * K1, cHh = cD1 + 128, cHl = cHh + kTcKH, cD2 = cHl + kTcKH;
    static_assert(cD2 + 64 * ND2 <= 512, "tensor memory budget");
    extern __shared__ __align__(128) float tsm[];
    const int D = A.D, M = A.M;
    float* const W1h = tsm + L::o_w1; float* const W1l = W1h + L::W1;
    float* const W2h = tsm + L::o_w2; float* const W2l = W2h + L::W2;
    float* const WH = tsm + L::o_wh;                                    // [M + 1][64] float32: rows < M means, row M value
    float* const HX = tsm + L::o_hx;                                    // [2 buffers][2 halves][9][128] head partial sums of the column halves
    float* const LS = tsm + L::o_ls;                                    // log_std[16], head biases[16]
    float* const HB = LS + 16;
    float* const ST0 = tsm + L::o_st;                                   // two raw observation tiles [128][D]
    const int st_floats = L::st_floats(D);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(tsm + L::o_bar(D));
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bar + kBarCount);
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // warp-uniform for the compiler too
    const int64_t E = A.E;
    const int64_t n_tiles = (E + kTcTile - 1) / kTcTile;
    const int n_my = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);     // tiles of this CTA: blockIdx.x + i * gridDim.x
    const uint32_t tile_bytes = (uint32_t)(kTcTile * D * 4);
    const bool tma_ok = A.obs_layout == 0 && ((reinterpret_cast<uintptr_t>(A.obs) & 15u) == 0);
    auto tile_of = [&](int i) { return (int64_t)blockIdx.x + (int64_t)i * gridDim.x; };
    auto tile_by_tma = [&](int64_t t) { return tma_ok && (t + 1) * kTcTile <= E; };

    // ---- one-time setup ----
    if (tid == 0) {
        const int counts[kBarCount] = { 1, 1, 32 * kWsWarpsP, 1, 1, 1, 32 * kWsWarpsE, 1, 1, 1, 32 * kWsWarpsE, 32 * kWsWarpsE,
                                        32 * kWsWarpsE, 32 * kWsWarpsE, 32 * kWsWarpsO, 32 * kWsWarpsO };
        for (int i = 0; i < kBarCount; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar + i)), "r"(counts[i]) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Programmatic dependent launch (the launch carries cudaLaunchAttributeProgrammaticStreamSerialization): this CTA may
    // have been scheduled while the previous kernel of the stream -- the env step that writes the observations -- was still
    // draining.  Barrier initialisation, the zero fill of the weight planes and the tensor-memory allocation touch no
    // global memory and run ahead; everything else waits for the previous kernel's memory here.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int i = tid; i < (2 * L::W1 + 2 * L::W2) / 4; i += NT) reinterpret_cast<float4*>(tsm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < 2 && i < n_my; i++)                        // the first two tiles' observations: in flight during the setup
            if (tile_by_tma(tile_of(i))) tc_bulk_g2s(ST0 + i * st_floats, A.obs + tile_of(i) * kTcTile * D, tile_bytes, bar + kBarObsFull + i);
    }
    constexpr float kPre = SPLIT ? 2.8853900817779268f : 1.0f;         // 2 log2(e), see ws_tanh
    ws_fill_b<SPLIT, K1, NT>(W1h, W1l, D, A.w1, A.b1, kPre);
    ws_fill_b<SPLIT, kTcKH, NT>(W2h, W2l, 64, A.w2, A.b2, kPre);
    for (int i = tid; i < (M + 1) * 64; i += NT) WH[i] = i < M * 64 ? __ldg(A.w_mu + i) : __ldg(A.w_v + (i - M * 64));
    if (tid < 16) { LS[tid] = tid < M ? A.log_std[tid] : 0.0f; HB[tid] = tid < M ? A.b_mu[tid] : (tid == M ? A.b_v[0] : 0.0f); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the weight planes -> visible to the tensor core's reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's quarter of the lanes
    const int row = ((warp & 3) << 5) | (tid & 31);                    // env row of the tile = TMEM lane
    bool alive = true;

    if (warp < kWsWarpE1) {
        // ================= P: observations -> layer-1 A planes =================
        constexpr int NC = K1 / 8;
        const bool even_d = (D & 1) == 0;
        for (int i = 0; i < n_my; i++) {
            const int buf = i & 1;
            const int64_t tile = tile_of(i), e = tile * kTcTile + row;
            const bool ev = e < E, staged = tile_by_tma(tile);
            float* const srow = ST0 + buf * st_floats + row * D;
            WG_WS_STAMP(0, tid == 0, 0);
            if (alive && staged) alive = tc_wait_bar(bar + kBarObsFull + buf, (uint32_t)(i >> 1) & 1u);
            if (!staged) {          // ragged last tile, feature-major or unaligned observations: this thread gathers its own row
#pragma unroll 1
                for (int f = 0; f < D; f++) srow[f] = ev ? __ldg(A.obs + (A.obs_layout ? (int64_t)f * E + e : e * D + f)) : 0.0f;
            }
            WG_WS_STAMP(0, tid == 0, 1);
            if (alive && i > 0) alive = tc_wait_bar(bar + kBarA1Free, (uint32_t)(i - 1) & 1u);   // layer 1 of the previous tile has read the planes
            WG_WS_STAMP(0, tid == 0, 2);
            if (alive) {
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    uint32_t hi[8], lo[8];
                    float x[8];                                         // the tail chunk reads past the row's end (padding / next row)
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
#pragma unroll
                    for (int q = 0; q < 8; q++) tc_obs_elem<SPLIT>(x[q], A.obs_scale, A.obs_clip, hi[q], lo[q]);
                    if (8 * j + 8 > D) {                                // the chunk with the end of the row: the 1 of the bias column, zero padding
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            if (8 * j + q >= D) { hi[q] = 8 * j + q == D ? 0x3f800000u : 0u; lo[q] = 0u; }
                    }
                    tc_st8(t_lane + cOh + 8 * j, hi);
                    if (SPLIT) tc_st8(t_lane + cOl + 8 * j, lo);
                }
                tc_wait_st();
                tc_fence_before();
                ws_arrive(bar + kBarA1Ready);
            }
            WG_WS_STAMP(0, tid == 0, 3);
            ws_group_sync(1, 32 * kWsWarpsP);                           // every producer is done with staging[buf]
            WG_WS_STAMP(0, tid == 0, 4);
            if (tid == 0 && alive && i + 2 < n_my && tile_by_tma(tile_of(i + 2)))
                tc_bulk_g2s(ST0 + buf * st_floats, A.obs + tile_of(i + 2) * kTcTile * D, tile_bytes, bar + kBarObsFull + buf);
        }
    } else if (warp == kWarpMma) {
        // ================= MMA: layer 1 of tile i, then layer 2 of tile i - 1 =================
        if (tc_elect_one()) {
            const uint64_t dW1h = tc_smem_desc(smem_u32(W1h), 64), dW1l = tc_smem_desc(smem_u32(W1l), 64);
            const uint64_t dW2h = tc_smem_desc(smem_u32(W2h), 64), dW2l = tc_smem_desc(smem_u32(W2l), 64);
            constexpr uint32_t kStepB = 2 * 64;                         // a k-step of 8 = two 16-byte chunks of 64 rows, in 16-byte units
            constexpr uint32_t id64 = tc_idesc(64);
            for (int i = 0; i <= n_my && alive; i++) {
                WG_WS_STAMP(1, true, 0);
                if (i < n_my) {
                    alive = tc_wait_bar(bar + kBarA1Ready, (uint32_t)i & 1u);
                    WG_WS_STAMP(1, true, 1);
                    if (!alive) break;
                    tc_fence_after();
                    const uint32_t d1 = tmem + cD1 + 64u * (uint32_t)(i & 1);
#pragma unroll
                    for (int kk = 0; kk < K1 / 8; kk++) {
                        tc_mma_ts(d1, tmem + cOh + 8 * kk, dW1h + kk * kStepB, id64, kk > 0);
                        if (SPLIT) {
                            tc_mma_ts(d1, tmem + cOl + 8 * kk, dW1h + kk * kStepB, id64, 1);
                            tc_mma_ts(d1, tmem + cOh + 8 * kk, dW1l + kk * kStepB, id64, 1);
                        }
                    }
                    tc_commit(bar + kBarD1Full + (i & 1));
                    tc_commit(bar + kBarA1Free);
                    WG_WS_STAMP(1, true, 2);
                }
                if (i > 0) {
                    const int t = i - 1, b2 = t % ND2;
                    alive = tc_wait_bar(bar + kBarE1Done, (uint32_t)t & 1u);
                    WG_WS_STAMP(1, true, 3);
                    if (alive && t >= ND2) alive = tc_wait_bar(bar + kBarD2Free + b2, (uint32_t)(t / ND2 - 1) & 1u);
                    WG_WS_STAMP(1, true, 4);
                    if (!alive) break;
                    tc_fence_after();
                    const uint32_t d2 = tmem + cD2 + 64u * (uint32_t)b2;
#pragma unroll
                    for (int kk = 0; kk < kTcKH / 8; kk++) {
                        tc_mma_ts(d2, tmem + cHh + 8 * kk, dW2h + kk * kStepB, id64, kk > 0);
                        if (SPLIT) {
                            tc_mma_ts(d2, tmem + cHl + 8 * kk, dW2h + kk * kStepB, id64, 1);
                            tc_mma_ts(d2, tmem + cHh + 8 * kk, dW2l + kk * kStepB, id64, 1);
                        }
                    }
                    tc_commit(bar + kBarD2Full + b2);
                    tc_commit(bar + kBarHFree);
                    WG_WS_STAMP(1, true, 5);
                }
            }
        }
        alive = __all_sync(0xffffffffu, alive);
    } else if (warp < kWsWarpE2) {
        // ================= E1: tanh(D1) -> layer-2 A planes =================
        const int half = (warp - kWsWarpE1) >> 2;
        for (int i = 0; i < n_my && alive; i++) {
            WG_WS_STAMP(2, tid == 32 * kWsWarpE1, 0);
            alive = tc_wait_bar(bar + kBarD1Full + (i & 1), (uint32_t)(i >> 1) & 1u);
            WG_WS_STAMP(2, tid == 32 * kWsWarpE1, 1);
            if (!alive) break;
            tc_fence_after();
            uint32_t a[16], b[16];
            const uint32_t d1 = t_lane + cD1 + 64u * (uint32_t)(i & 1) + 32 * half;
            tc_ld16(d1, a);
            tc_ld16(d1 + 16, b);
            tc_wait_ld();
            WG_WS_STAMP(2, tid == 32 * kWsWarpE1, 2);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const float y = ws_tanh<SPLIT>(__uint_as_float(c ? b[q] : a[q]));
                    act_split<SPLIT>(y, hi[q], lo[q]);
                }
                if (c == 0) WG_WS_STAMP(2, tid == 32 * kWsWarpE1, 3);
                if (c == 0 && i > 0) {                                  // layer 2 of the previous tile has read the planes
                    alive = tc_wait_bar(bar + kBarHFree, (uint32_t)(i - 1) & 1u);
                    tc_fence_after();
                }
                if (c == 0) WG_WS_STAMP(2, tid == 32 * kWsWarpE1, 4);
                if (alive) {
                    tc_st16(t_lane + cHh + 32 * half + 16 * c, hi);
                    if (SPLIT) tc_st16(t_lane + cHl + 32 * half + 16 * c, lo);
                }
            }
            if (!alive) break;
            if (half) {
                const uint32_t one[8] = { 0x3f800000u, 0u, 0u, 0u, 0u, 0u, 0u, 0u }, zero[8] = { 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u };
                tc_st8(t_lane + cHh + 64, one);
                if (SPLIT) tc_st8(t_lane + cHl + 64, zero);
            }
            tc_wait_st();
            tc_fence_before();
            ws_arrive(bar + kBarE1Done);
            WG_WS_STAMP(2, tid == 32 * kWsWarpE1, 5);
        }
    } else if (warp < kWsWarpO) {
        // ================= E2: tanh(D2) -> partial sums of the heads =================
        const int half = (warp - kWsWarpE2) >> 2;
        for (int i = 0; i < n_my && alive; i++) {
            const int b2 = i % ND2, hb = i & 1;
            WG_WS_STAMP(3, tid == 32 * kWsWarpE2, 0);
            alive = tc_wait_bar(bar + kBarD2Full + b2, (uint32_t)(i / ND2) & 1u);
            WG_WS_STAMP(3, tid == 32 * kWsWarpE2, 1);
            if (!alive) break;
            tc_fence_after();
            uint32_t a[16], b[16];
            tc_ld16(t_lane + cD2 + 64u * (uint32_t)b2 + 32 * half, a);
            tc_ld16(t_lane + cD2 + 64u * (uint32_t)b2 + 32 * half + 16, b);
            tc_wait_ld();
            tc_fence_before();
            ws_arrive(bar + kBarD2Free + b2);                           // D2[b2] is in registers: the tensor core may overwrite it
            WG_WS_STAMP(3, tid == 32 * kWsWarpE2, 2);
            float y[32];
#pragma unroll
            for (int q = 0; q < 16; q++) { y[q] = ws_tanh<SPLIT>(__uint_as_float(a[q])); y[16 + q] = ws_tanh<SPLIT>(__uint_as_float(b[q])); }
            WG_WS_STAMP(3, tid == 32 * kWsWarpE2, 3);
            if (i >= 2) alive = tc_wait_bar(bar + kBarHxFree + hb, (uint32_t)((i >> 1) - 1) & 1u);   // the output warps have read HX[hb]
            if (!alive) break;
            WG_WS_STAMP(3, tid == 32 * kWsWarpE2, 4);
            float* const hx = HX + ((hb * 2 + half) * kTcMaxHeads) * kTcTile + row;
#pragma unroll 1
            for (int n = 0; n <= M; n++) {                              // rolled: one copy of the 32 FMAs in the instruction cache
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
                hx[n * kTcTile] = s0 + s1;
            }
            ws_arrive(bar + kBarHxFull + hb);                           // release: the partial sums are visible to the output warps
            WG_WS_STAMP(3, tid == 32 * kWsWarpE2, 5);
        }
    } else {
        // ================= O: heads -> gaussian sample, log-prob, value -> global memory (one thread per env) =================
        const uint32_t step = A.step_index + (A.step_counter ? __ldg(A.step_counter) : 0u);
        const int n_pairs = (M + 1) / 2;
        const int og = SA::kFused ? (warp - kWsWarpO) >> 2 : 0;        // fused: output group og takes the tiles of its parity
        for (int i = og; i < n_my && alive; i += WsCfg<SA>::kGroupsO) {
            const int hb = i & 1;
            const int64_t e = tile_of(i) * kTcTile + row;
            const bool ev = e < E;
            // fused env step: the env's packed state is requested before anything else (latency under the noise + wait)
            float sv[SA::kStateRegs];
            float act_reg[2] = { 0.0f, 0.0f };
            if constexpr (SA::kFused) { if (ev) S.load(sv, tile_of(i), row); }
            // the first action pair's noise does not depend on the network: evaluated while the tile is still upstream
            float2 z0 = make_float2(0.0f, 0.0f);
            if (A.sample) z0 = pol_normal2(A.seed_lo, A.seed_hi, A.env_offset + (uint32_t)e, step, 0u);
            alive = tc_wait_bar(bar + kBarHxFull + hb, (uint32_t)(i >> 1) & 1u);
            if (!alive) break;
            const float* const h0 = HX + (hb * 2 * kTcMaxHeads) * kTcTile + row;
            const float* const h1 = h0 + kTcMaxHeads * kTcTile;
            float lp = 0.0f;
#pragma unroll 1
            for (int pr = 0; pr < n_pairs; pr++) {
                float2 z = z0;
                if (pr > 0 && A.sample) z = pol_normal2(A.seed_lo, A.seed_hi, A.env_offset + (uint32_t)e, step, (uint32_t)pr);
                const int n0 = 2 * pr, n1 = n0 + 1 < M ? n0 + 1 : n0;    // an odd M's last pair: the second slot repeats the first, not stored
                const float m0 = h0[n0 * kTcTile] + h1[n0 * kTcTile] + HB[n0], m1 = h0[n1 * kTcTile] + h1[n1 * kTcTile] + HB[n1];
                const float l0 = LS[n0], l1 = LS[n1];
                const float a0 = A.sample ? __fmaf_rn(__expf(l0), z.x, m0) : m0, a1 = A.sample ? __fmaf_rn(__expf(l1), z.y, m1) : m1;
                lp += -0.5f * z.x * z.x - l0 - 0.9189385332046727f;
                if (n0 + 1 < M) lp += -0.5f * z.y * z.y - l1 - 0.9189385332046727f;
                if constexpr (SA::kFused) { if (pr == 0) { act_reg[0] = a0; act_reg[1] = a1; } }      // fused bodies have M == 2
                if (ev) {
                    if (A.mean) { A.mean[(int64_t)n0 * E + e] = m0; if (n0 + 1 < M) A.mean[(int64_t)n1 * E + e] = m1; }
                    if (A.action) {
                        A.action[A.act_layout ? (int64_t)n0 * E + e : e * M + n0] = a0;
                        if (n0 + 1 < M) A.action[A.act_layout ? (int64_t)n1 * E + e : e * M + n1] = a1;
                    }
                }
            }
            const float val = h0[M * kTcTile] + h1[M * kTcTile] + HB[M];
            ws_arrive(bar + kBarHxFree + hb);                           // HX[hb] has been read
            if (ev) {
                if (A.value) A.value[e] = val;
                if (A.logp) A.logp[e] = lp;
            }
            if constexpr (SA::kFused) {
                // ---- PhysicsEnv.step of this env with the action just sampled (wg_policy_step.cuh); the next observation
                // rows of the warp's 32 envs leave with one TMA bulk store, like the stand-alone step kernel's ----
                float* const OT = tsm + L::o_ot + og * SA::kObsFloats;
                constexpr int DO = SA::D;
                if (ev) S.step(sv, act_reg, tile_of(i), row, e, OT + row * DO);
                __syncwarp();
                const int64_t ew = tile_of(i) * kTcTile + (row & ~31);                 // first env of this warp
                const int64_t remw = E - ew;
                if (remw > 0 && S.A.obs) {
                    float* wt = OT + (row & ~31) * DO;
                    if (remw >= 32 && ((reinterpret_cast<uintptr_t>(S.A.obs) & 15u) == 0)) {
                        fence_proxy_async();
                        __syncwarp();
                        if ((tid & 31) == 0) { bulk_s2g(S.A.obs + ew * DO, wt, (uint32_t)(32 * DO * 4)); bulk_commit(); bulk_wait_read0(); }
                    } else {
                        const int total = (remw < 32 ? (int)remw : 32) * DO;
                        for (int idx = tid & 31; idx < total; idx += 32) S.A.obs[ew * DO + idx] = wt[idx];
                    }
                }
                __syncwarp();                                           // the tile's rows may be overwritten by the next tile
            }
        }
    }
    if (!alive && error_flag && (tid & 31) == 0) atomicExch(error_flag, 1);
    // ---- teardown: the allocating warp frees the tensor memory once every warp is done with it ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

}  // namespace wg
