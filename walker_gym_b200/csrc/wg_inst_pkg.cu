// wg_inst_pkg.cu -- the package lineage's Environment.update_physics (wg_pkg_update_physics).
#include "wg_launch.cuh"
#include "wg_pkg.cuh"
namespace wg {

int launch_pkg_update(const wg_pkg_system* sy, const wg_pkg_params* p, float* pos, float* vel, float* old_a,
                      int64_t E, int32_t n_steps, cudaStream_t s) {
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
