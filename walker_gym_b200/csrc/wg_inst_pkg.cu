// wg_inst_pkg.cu -- the package lineage's Environment.update_physics (wg_pkg_update_physics).
#include "wg_launch.cuh"
#include "wg_pkg.cuh"
namespace wg {

// bodies of gym/optimized_walker/walker.py with a register-resident kernel: endpoints as (point1, point2) pairs
// (box, 8 points / 12 springs, was measured slower register-resident -- 190 us at 128 registers / 4 CTAs per SM,
// 202 us at 154 registers / 3 CTAs, vs 163 us per 2^20 envs -- so it and the larger bodies use the run-time-topology
// kernel)
// leg2 (:368-393): body, two 3-point legs
WG_STATIC_TOPO(PkgLeg2, 2, 7, 6, 0, 0,1, 1,2, 2,3, 0,4, 4,5, 5,6)
// balance3 (:452-467): pivot + 3 bobs; balance2 / balance1 / test are chains of 3 / 2 / 2 points
WG_STATIC_TOPO(PkgChain4, 3, 4, 3, 0, 0,1, 1,2, 2,3)
WG_STATIC_TOPO(PkgChain3, 4, 3, 2, 0, 0,1, 1,2)
WG_STATIC_TOPO(PkgChain2, 5, 2, 1, 0, 0,1)

template <class Topo>
static bool pkg_matches(const wg_pkg_system* sy) {
    if (sy->n_point != Topo::N || sy->n_spring != Topo::S) return false;
    for (int s = 0; s < Topo::S; s++)
        if (sy->si[s] != Topo::si(s) || sy->sj[s] != Topo::sj(s)) return false;
    return true;
}

int pkg_variant(const wg_pkg_system* sy) {
    if (pkg_matches<PkgLeg2>(sy)) return PkgLeg2::kId;
    if (pkg_matches<PkgChain4>(sy)) return PkgChain4::kId;
    if (pkg_matches<PkgChain3>(sy)) return PkgChain3::kId;
    if (pkg_matches<PkgChain2>(sy)) return PkgChain2::kId;
    return 0;
}

// Breadth-first order of the spring graph (components one after another) cut into `parts` chunks of near-equal size.
static void build_pkg_partition(const wg_pkg_system* sy, int parts, PkgPartTables& pt) {
    const int N = sy->n_point, S = sy->n_spring;
    int order[kMaxMass], n_order = 0, owner[kMaxMass];
    bool seen[kMaxMass] = {};
    for (int root = 0; root < N; root++) {
        if (seen[root]) continue;
        int head = n_order;
        order[n_order++] = root; seen[root] = true;
        while (head < n_order) {
            const int u = order[head++];
            for (int s = 0; s < S; s++) {
                int v = -1;
                if (sy->si[s] == u) v = sy->sj[s]; else if (sy->sj[s] == u) v = sy->si[s];
                if (v >= 0 && !seen[v]) { seen[v] = true; order[n_order++] = v; }
            }
        }
    }
    memset(&pt, 0, sizeof(pt));
    for (int q = 0; q < N; q++) owner[order[q]] = (int)(((int64_t)q * parts) / N);
    for (int n = 0; n < N; n++) {
        const int p = owner[n];
        pt.point[p][pt.n_point[p]++] = (uint8_t)n;
        pt.own_mask[p] |= 1u << n;
    }
    for (int s = 0; s < S; s++) {                       // ascending spring order = list order
        const int pi = owner[sy->si[s]], pj = owner[sy->sj[s]];
        pt.spring[pi][pt.n_spring[pi]++] = (uint8_t)s;
        if (pj != pi) pt.spring[pj][pt.n_spring[pj]++] = (uint8_t)s;
    }
}

template <int LP>
static int launch_pkg_part(const PkgArgs& A, const wg_pkg_system* sy, cudaStream_t s) {
    static thread_local PkgPartArgs PA;
    PA.A = A;
    build_pkg_partition(sy, LP, PA.pt);
    constexpr int EB = kBlock / LP;
    const size_t smem = sizeof(float) * (size_t)(9 * sy->n_point) * (EB + 1) + (size_t)LP * (kMaxSpring + kMaxMass);
    auto kern = pkg_update_part_kernel<LP>;
    if (smem > 40 * 1024) {          // plus ~2.5 KB of static shared memory (staged tables)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<(unsigned)((A.E + EB - 1) / EB), kBlock, smem, s>>>(PA);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "update_physics kernel (partitioned) launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

template <class Topo>
static int launch_pkg_static(const PkgArgs& A, cudaStream_t s) {
    pkg_update_static_kernel<Topo><<<(unsigned)((A.E + kBlock - 1) / kBlock), kBlock, 0, s>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "update_physics kernel (static) launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

int launch_pkg_update(const wg_pkg_system* sy, const wg_pkg_params* p, float* pos, float* vel, float* old_a,
                      int64_t E, int32_t n_steps, bool force_generic, cudaStream_t s) {
    static thread_local PkgArgs A;
    memset(&A, 0, sizeof(A));
    for (int n = 0; n < sy->n_point; n++) {
        const float mf = (float)sy->mass[n];
        const ConstDiv cd = make_const_div(mf);
        A.mass_f[n] = mf; A.mass_r[n] = cd.r; A.mass_kind[n] = cd.kind;
        for (int c = 0; c < 3; c++) {
            // point.forced(self.gravity * point.m) on a zeroed accumulator: float32 product, float32 quotient
            volatile float gm = p->gravity[c] * mf;
            volatile float q = gm / mf;
            A.ga[n * 3 + c] = 0.0f + q;
        }
        if (sy->fixed[n]) A.fixed_mask |= 1u << n;
    }
    for (int k = 0; k < sy->n_spring; k++) {
        A.si[k] = (uint8_t)sy->si[k]; A.sj[k] = (uint8_t)sy->sj[k];
        A.srest[k] = sy->srest[k]; A.sk[k] = sy->sk[k];
        if (sy->sstring[k]) A.string_mask[k >> 5] |= 1u << (k & 31);
    }
    A.damping = p->damping; A.drag_c = p->drag_c; A.ground_level = p->ground_level;
    A.restitution = p->restitution; A.friction = p->friction; A.dt = p->dt; A.min_dist = p->min_dist;
    A.ground = p->ground; A.n_point = sy->n_point; A.n_spring = sy->n_spring; A.n_steps = n_steps;
    A.pos = pos; A.vel = vel; A.old_a = old_a; A.E = E;
    switch (force_generic ? 0 : pkg_variant(sy)) {
        case PkgLeg2::kId:   return launch_pkg_static<PkgLeg2>(A, s);
        case PkgChain4::kId: return launch_pkg_static<PkgChain4>(A, s);
        case PkgChain3::kId: return launch_pkg_static<PkgChain3>(A, s);
        case PkgChain2::kId: return launch_pkg_static<PkgChain2>(A, s);
        default: break;
    }
    // larger bodies: several lanes per env (WG_TUNE_PART: -1 automatic, 0 never, 2 / 4 / 8 forced)
    int parts = tuning(WG_TUNE_PART);
    // measured on the reference's bodies (2^20 envs, us per update): insect (21 points) 769 / 746 / 607 / 961 with
    // 1 / 2 / 4 / 8 lanes per env, humanb (14) 410 / 402 / 459 / 874, box (8) 163 / 208 / 345 / 806
    if (parts < 0) parts = sy->n_point >= 18 ? 4 : (sy->n_point >= 12 ? 2 : 0);
    if (parts > sy->n_point) parts = 0;
    if (!force_generic && parts >= 2) {
        if (parts == 2) return launch_pkg_part<2>(A, sy, s);
        if (parts == 4) return launch_pkg_part<4>(A, sy, s);
        return launch_pkg_part<8>(A, sy, s);
    }
    const size_t smem = sizeof(float) * (size_t)(9 * sy->n_point) * (kBlock + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(pkg_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    pkg_update_kernel<<<(unsigned)((E + kBlock - 1) / kBlock), kBlock, smem, s>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "update_physics kernel launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

}  // namespace wg
