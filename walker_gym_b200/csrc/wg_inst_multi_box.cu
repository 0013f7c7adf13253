// wg_inst_multi_box.cu -- instantiates the T-steps-per-launch kernel for the Box-v0 spring graph.
#include "wg_launch.cuh"
#include "wg_kernels_multi.cuh"
namespace wg {
int launch_box_multi(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int n_steps, int64_t as, cudaStream_t s) {
    return launch_multi_flags<TopoBox>(t, p, b, E, n_steps, as, s);
}
}  // namespace wg
