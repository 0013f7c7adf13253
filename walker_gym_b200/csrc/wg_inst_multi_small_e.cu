// wg_inst_multi_small_e.cu -- instantiates the T-steps-per-launch kernel for walker.py leg.
#include "wg_launch.cuh"
#include "wg_kernels_multi.cuh"
namespace wg {
int launch_leg_multi(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int n_steps, int64_t as, cudaStream_t s) {
    return launch_multi_flags<TopoLeg>(t, p, b, E, n_steps, as, s);
}
}  // namespace wg
