// wg_pkg.cuh -- the reference's *package* lineage: Environment.update_physics of
// gym/optimized_walker/env.py:135-184 over Point.forced / anti_forced / resilience / run1 of
// gym/optimized_walker/core.py:81-121,184-200 (DingPoint :259-275), batched over independent copies
// of one point-and-spring system.
//
// A different model from PhysicsEnv.step: physically signed springs with optional one-sided
// ("string") behaviour, gravity as an acceleration ((g*m)/m), multiplicative velocity damping,
// quadratic drag, and a ground that clamps the position and reflects the velocity; no action,
// observation, reward or done.  Because nothing enters or leaves between steps, the kernel keeps the
// state on chip for n_steps consecutive updates: ONE HBM read and ONE HBM write per launch.
//
// One thread owns one env and applies the springs strictly in list order, with the reference's
// separately rounded float32 operations (-fmad=false), so the result is bit-identical to NumPy.
#pragma once
#include "wg_physics.cuh"

namespace wg {

struct PkgArgs {
    float ga[kMaxMass * 3];         // 0.0f + (gravity * float32(m)) / float32(m): what `a` holds after the gravity pass
    float mass_f[kMaxMass];         // float32(m): divisor of every float32 force
    float mass_r[kMaxMass];         // RN(1 / float32(m))
    int32_t mass_kind[kMaxMass];    // div_const kind
    float srest[kMaxSpring], sk[kMaxSpring];
    uint8_t si[kMaxSpring], sj[kMaxSpring];
    uint32_t string_mask[kMaxSpring / 32];
    uint32_t fixed_mask;            // DingPoint bits
    float damping, drag_c, ground_level, restitution, friction, dt, min_dist;
    int32_t ground, n_point, n_spring, n_steps;
    float* pos; float* vel; float* old_a;
    int64_t E;
};

// One spring: Point.resilience (core.py:93-121) = two anti_forced calls (:85-91) with the same f_size.
//   d = p1 - p2; L = |d|; dx = L - x; f_size = -dx * k (0 for a slack string);
//   p1.anti_forced(p2, f_size): dir = p2 - p1, a1 += ((-f_size * dir) / max(|dir|, r)) / m1
//   p2.anti_forced(p1, f_size): dir = p1 - p2, a2 += ((-f_size * dir) / max(|dir|, r)) / m2
// p2 - p1 == -(p1 - p2) exactly and every operation is sign-symmetric, so the force on p1 is the exact
// negation of the force on p2 and the norm is evaluated once.
struct PkgRuntimeTopo {
    const uint8_t* si_; const uint8_t* sj_;
    __device__ __forceinline__ int si(int k) const { return si_[k]; }
    __device__ __forceinline__ int sj(int k) const { return sj_[k]; }
};

// Point.forced: a += f / m for a force that is NOT a raw packed product (the quotients of div3_len)
template <class Store>
__device__ __forceinline__ void pkg_forced(const PkgArgs& A, Store& st, int n, const V3& f) {
    const float m = A.mass_f[n], r = A.mass_r[n]; const int kd = A.mass_kind[n];
    float2 axy = make_float2(st.acc(n, 0), st.acc(n, 1));
    float az = st.acc(n, 2);
    if (kd == 0) { axy = make_float2(axy.x + f.xy.x, axy.y + f.xy.y); az = az + f.z; }
    else if (kd <= 2) { axy = __fadd2_rn(axy, div_smallint2(f.xy, m, r)); az = az + div_smallint(f.z, m, r); }
    else { axy = make_float2(axy.x + div_rn(f.xy.x, m), axy.y + div_rn(f.xy.y, m)); az = az + div_rn(f.z, m); }
    st.acc(n, 0) = axy.x; st.acc(n, 1) = axy.y; st.acc(n, 2) = az;
}

template <class Topo, class Store>
__device__ __forceinline__ void pkg_spring(const Topo& topo, const PkgArgs& A, Store& st, int sp, uint32_t skip_mask) {
    const int i = topo.si(sp), j = topo.sj(sp);
    // x / y halves packed (wg_math.cuh): the same IEEE operations, two instructions per 3-vector operation
    const V3 d = v3_sub(v3(st.pos(i, 0), st.pos(i, 1), st.pos(i, 2)), v3(st.pos(j, 0), st.pos(j, 1), st.pos(j, 2)));
    const float L = np_norm3(d);
    const float dx = L - A.srest[sp];
    const bool slack = (dx < 0.0f) && ((A.string_mask[sp >> 5] >> (sp & 31)) & 1u);
    const float nfs = slack ? 0.0f : -((-dx) * A.sk[sp]);                 // -f_size
    const float dist = (A.min_dist > L) ? A.min_dist : L;                  // python max(norm, Config.r)
    V3 f = v3_scale(d, nfs);                                               // force on p2 (dir = d)
    div3_len<true>(f, dist);                                               // dist > 0 or NaN: always a division
    // skip_mask: points whose accumulator this thread must not touch -- DingPoints (forced() is a no-op) and, in the
    // partitioned kernel, points owned by another lane
    if (!((skip_mask >> i) & 1u)) pkg_forced(A, st, i, v3_neg(f));
    if (!((skip_mask >> j) & 1u)) pkg_forced(A, st, j, f);
}

// Everything update_physics does to one point after the springs: damping (:153-154), quadratic drag
// (:157-161), Point.run1 (core.py:184-200), ground (:167-181).  The passes of the reference are
// per-point independent, so running them back to back for one point keeps its order of operations.
template <class Store>
__device__ __forceinline__ void pkg_point(const PkgArgs& A, Store& st, int n) {
    float vx = st.vel(n, 0), vy = st.vel(n, 1), vz = st.vel(n, 2);
    float ax = st.acc(n, 0), ay = st.acc(n, 1), az = st.acc(n, 2);
    const bool fixed = (A.fixed_mask >> n) & 1u;
    if (!fixed) {
        const float2 vxy = __fmul2_rn(make_float2(vx, vy), bc2(A.damping));
        vx = vxy.x; vy = vxy.y; vz = vz * A.damping;
        const float cs = A.drag_c * np_norm3(v3(vx, vy, vz));
        const float m = A.mass_f[n], r = A.mass_r[n]; const int kd = A.mass_kind[n];
        const float2 dxy = __fmul2_rn(vxy, bc2(cs));                       // cs * v: packed products, then scalar / quotient adds
        ax = ax + div_const(dxy.x, m, r, kd);
        ay = ay + div_const(dxy.y, m, r, kd);
        az = az + div_const(cs * vz, m, r, kd);
        st.acc(n, 0) = ax; st.acc(n, 1) = ay; st.acc(n, 2) = az;
    }
    // run1: every point, DingPoints drift; products packed, additions scalar (wg_math.cuh CAUTION)
    const float2 avt = __fmul2_rn(make_float2(ax, ay), bc2(A.dt));
    vx = vx + avt.x; vy = vy + avt.y; vz = vz + az * A.dt;
    const float2 pvt = __fmul2_rn(make_float2(vx, vy), bc2(A.dt));
    float px = st.pos(n, 0) + pvt.x, py = st.pos(n, 1) + pvt.y, pz = st.pos(n, 2) + vz * A.dt;
    if (!fixed && A.ground && py <= A.ground_level) {
        py = A.ground_level;
        if (vy < 0.0f) { vy = (-vy) * A.restitution; vx = vx * A.friction; vz = vz * A.friction; }
    }
    st.vel(n, 0) = vx; st.vel(n, 1) = vy; st.vel(n, 2) = vz;
    st.pos(n, 0) = px; st.pos(n, 1) = py; st.pos(n, 2) = pz;
}

// Run-time topology; per-env state in a shared-memory tile [row][kBlock + 1] that stays resident for
// all n_steps updates.  Loads and stores are coalesced SoA rows.
__global__ void __launch_bounds__(kBlock)
pkg_update_kernel(const __grid_constant__ PkgArgs A) {
    extern __shared__ float smem[];
    constexpr int PITCH = kBlock + 1;
    const int P = A.n_point, S = A.n_spring;
    const int tid = threadIdx.x;
    const int64_t E = A.E;
    const int64_t e = (int64_t)blockIdx.x * kBlock + tid;
    if (e >= E) return;
    SmemStore st{ smem + tid, PITCH, P };
    const PkgRuntimeTopo topo{ A.si, A.sj };
    for (int r = 0; r < 3 * P; r++) {
        st.base[r * PITCH] = A.pos[(int64_t)r * E + e];
        st.base[(3 * P + r) * PITCH] = A.vel[(int64_t)r * E + e];
    }
    for (int t = 0; t < A.n_steps; t++) {
        for (int n = 0; n < P; n++) {                                      // zero (:141-142) + gravity (:145-146)
            const bool fixed = (A.fixed_mask >> n) & 1u;
            st.acc(n, 0) = fixed ? 0.0f : A.ga[n * 3 + 0];
            st.acc(n, 1) = fixed ? 0.0f : A.ga[n * 3 + 1];
            st.acc(n, 2) = fixed ? 0.0f : A.ga[n * 3 + 2];
        }
        for (int sp = 0; sp < S; sp++) pkg_spring(topo, A, st, sp, A.fixed_mask);  // (:149-150)
        for (int n = 0; n < P; n++) pkg_point(A, st, n);
    }
    for (int r = 0; r < 3 * P; r++) {
        A.pos[(int64_t)r * E + e] = st.base[r * PITCH];
        A.vel[(int64_t)r * E + e] = st.base[(3 * P + r) * PITCH];
        if (A.old_a) A.old_a[(int64_t)r * E + e] = st.base[(6 * P + r) * PITCH];
    }
}

// Larger bodies: LP adjacent lanes share one env ("point partition", the scheme of wg_kernels_part.cuh).  With one
// thread per env a 21-point body needs 756 bytes of shared-memory state per thread, which leaves an SM 9 warps; here
// the points are split into LP parts, lane `part` evaluates -- in list order -- every spring that touches one of its
// points and accumulates only into ITS points (a spring crossing two parts is evaluated by both owners: same inputs,
// same operations, same bits), then damps / integrates / grounds its own points.  Per-point accumulation order is the
// reference's, so the bits do not change, while an env exposes LP-fold parallelism and the SM holds LP times the warps.
struct PkgPartTables {
    uint8_t spring[8][kMaxSpring];    // springs touching part p, ascending (= list order)
    uint8_t point[8][kMaxMass];       // points owned by part p
    uint8_t n_spring[8], n_point[8];
    uint32_t own_mask[8];
};
struct PkgPartArgs { PkgArgs A; PkgPartTables pt; };

template <int LP>
__global__ void __launch_bounds__(kBlock)
pkg_update_part_kernel(const __grid_constant__ PkgPartArgs PA) {
    extern __shared__ float smem[];
    // lanes of one warp work on different springs and points: per-spring constants would be divergent constant-bank
    // reads, so the tables are staged once per block in shared memory
    __shared__ PkgArgs A;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&PA.A);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&A);
        for (int i = threadIdx.x; i < (int)(sizeof(PkgArgs) / 4); i += kBlock) dst[i] = src[i];
    }
    constexpr int EB = kBlock / LP, PITCH = EB + 1;
    const int P = PA.A.n_point;
    const int tid = threadIdx.x, el = tid / LP, part = tid % LP;
    const int64_t E = PA.A.E;
    const int64_t e0 = (int64_t)blockIdx.x * EB;
    const bool valid = e0 + el < E;
    const int64_t rem = E - e0;
    const int nvalid = rem < EB ? (int)rem : EB;
    uint8_t* tab = reinterpret_cast<uint8_t*>(smem + 9 * P * PITCH);
    uint8_t* my_springs = tab + part * kMaxSpring;
    uint8_t* my_points = tab + LP * kMaxSpring + part * kMaxMass;
    for (int i = tid; i < LP * kMaxSpring; i += kBlock) tab[i] = PA.pt.spring[i / kMaxSpring][i % kMaxSpring];
    for (int i = tid; i < LP * kMaxMass; i += kBlock) tab[LP * kMaxSpring + i] = PA.pt.point[i / kMaxMass][i % kMaxMass];
    const int n_my_springs = PA.pt.n_spring[part], n_my_points = PA.pt.n_point[part];
    const uint32_t skip = PA.A.fixed_mask | ~PA.pt.own_mask[part];
    // single HBM read: the block's EB envs of every state row, coalesced
    for (int idx = tid; idx < 3 * P * EB; idx += kBlock) {
        const int r = idx / EB, c = idx - r * EB;
        if (c < nvalid) {
            smem[r * PITCH + c] = PA.A.pos[(int64_t)r * E + e0 + c];
            smem[(3 * P + r) * PITCH + c] = PA.A.vel[(int64_t)r * E + e0 + c];
        }
    }
    __syncthreads();
    SmemStore st{ smem + el, PITCH, P };
    const PkgRuntimeTopo topo{ A.si, A.sj };
    for (int t = 0; t < A.n_steps; t++) {
        if (valid) {
            for (int q = 0; q < n_my_points; q++) {
                const int n = my_points[q];
                const bool fixed = (A.fixed_mask >> n) & 1u;
                st.acc(n, 0) = fixed ? 0.0f : A.ga[n * 3 + 0];
                st.acc(n, 1) = fixed ? 0.0f : A.ga[n * 3 + 1];
                st.acc(n, 2) = fixed ? 0.0f : A.ga[n * 3 + 2];
            }
            for (int q = 0; q < n_my_springs; q++) pkg_spring(topo, A, st, my_springs[q], skip);
        }
        __syncwarp();                              // every lane of the env has read the old positions
        if (valid)
            for (int q = 0; q < n_my_points; q++) pkg_point(A, st, my_points[q]);
        __syncwarp();                              // new positions / velocities visible to the env's other lanes
    }
    __syncthreads();
    for (int idx = tid; idx < 3 * P * EB; idx += kBlock) {
        const int r = idx / EB, c = idx - r * EB;
        if (c < nvalid) {
            PA.A.pos[(int64_t)r * E + e0 + c] = smem[r * PITCH + c];
            PA.A.vel[(int64_t)r * E + e0 + c] = smem[(3 * P + r) * PITCH + c];
            if (PA.A.old_a) PA.A.old_a[(int64_t)r * E + e0 + c] = smem[(6 * P + r) * PITCH + c];
        }
    }
}

// Register-resident specialisation for small bodies the reference ships (compile-time endpoints): every state
// element has a static register, the n_steps loop runs without touching memory.
template <class Topo>
__global__ void __launch_bounds__(kBlock, Topo::N <= 4 ? 6 : 4)
pkg_update_static_kernel(const __grid_constant__ PkgArgs A) {
    constexpr int P = Topo::N, S = Topo::S;
    const int64_t E = A.E;
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= E) return;
    const Topo topo;
    RegStore<P, 0> st;
#pragma unroll
    for (int r = 0; r < 3 * P; r++) {
        st.p_[r / 3][r % 3] = A.pos[(int64_t)r * E + e];
        st.v_[r / 3][r % 3] = A.vel[(int64_t)r * E + e];
    }
    for (int t = 0; t < A.n_steps; t++) {
#pragma unroll
        for (int n = 0; n < P; n++) {
            const bool fixed = (A.fixed_mask >> n) & 1u;
#pragma unroll
            for (int c = 0; c < 3; c++) st.a_[n][c] = fixed ? 0.0f : A.ga[n * 3 + c];
        }
#pragma unroll
        for (int sp = 0; sp < S; sp++) pkg_spring(topo, A, st, sp, A.fixed_mask);
#pragma unroll
        for (int n = 0; n < P; n++) pkg_point(A, st, n);
    }
#pragma unroll
    for (int r = 0; r < 3 * P; r++) {
        A.pos[(int64_t)r * E + e] = st.p_[r / 3][r % 3];
        A.vel[(int64_t)r * E + e] = st.v_[r / 3][r % 3];
        if (A.old_a) A.old_a[(int64_t)r * E + e] = st.a_[r / 3][r % 3];
    }
}

}  // namespace wg
