// wg_inst_multi_box.cu -- instantiates the T-steps-per-launch kernel for the Box-v0 spring graph.
#include "wg_launch.cuh"
#include "wg_kernels_multi.cuh"
namespace wg {
int launch_box_multi(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int n_steps, cudaStream_t s) {
    if (mass_mode(t) == 0)
        return p->in3d ? launch_multi_packed<TopoBox, true, 0>(t, p, b, E, n_steps, s)
                       : launch_multi_packed<TopoBox, false, 0>(t, p, b, E, n_steps, s);
    return p->in3d ? launch_multi_packed<TopoBox, true, 1>(t, p, b, E, n_steps, s)
                   : launch_multi_packed<TopoBox, false, 1>(t, p, b, E, n_steps, s);
}
}  // namespace wg
