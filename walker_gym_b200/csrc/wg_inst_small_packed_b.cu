// wg_inst_small_packed_b.cu -- packed-state step kernels for smaller walker.py bodies.
#include "wg_launch.cuh"
namespace wg {
int launch_hat_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoHat>(t, p, b, E, s);
}
}  // namespace wg
