// wg_inst_generic.cu -- run-time-topology step kernel, reset kernel (K2), stats reduction (K3).
#include "wg_launch.cuh"
namespace wg {
static size_t generic_smem(const wg_topology* t) {
    return sizeof(float) * (size_t)(11 * t->n_mass + t->n_muscle + 3) * (kBlock + 1);
}

template <bool IN3D, bool ROWMAJOR>
static int launch_generic(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    StepArgs<kMaxMass, kMaxSpring> A;
    fill_args(A, t, p, b, E);
    const size_t smem = generic_smem(t);
    auto kern = step_generic_kernel<IN3D, ROWMAJOR>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    const unsigned grid = (unsigned)((E + kBlock - 1) / kBlock);
    kern<<<grid, kBlock, smem, s>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "generic step kernel launch: %s", cudaGetErrorString(e));
    return WG_OK;
}


int launch_generic_step(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    const bool rm = b->obs_layout == 0;
    if (p->in3d) return rm ? launch_generic<true, true>(t, p, b, E, s) : launch_generic<true, false>(t, p, b, E, s);
    return rm ? launch_generic<false, true>(t, p, b, E, s) : launch_generic<false, false>(t, p, b, E, s);
}

int launch_reset(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int mode, const uint8_t* mask, cudaStream_t s) {
    StepArgs<kMaxMass, kMaxSpring> A;
    fill_args(A, t, p, b, E);
    const unsigned grid = (unsigned)((E + kBlock - 1) / kBlock);
    if (p->in3d) reset_kernel<true><<<grid, kBlock, 0, s>>>(A, mode, mask, b->obs_layout);
    else reset_kernel<false><<<grid, kBlock, 0, s>>>(A, mode, mask, b->obs_layout);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "reset kernel launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

int launch_stats(const float* fin_stats, int64_t E, double* out8, cudaStream_t s) {
    stats_reduce_kernel<1024><<<1, 1024, 0, s>>>(fin_stats, E, out8);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "stats kernel launch: %s", cudaGetErrorString(e));
    return WG_OK;
}
}  // namespace wg
