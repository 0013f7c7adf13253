// wg_inst_small_packed_d.cu -- packed-state step kernels for smaller walker.py bodies.
#include "wg_launch.cuh"
namespace wg {
int launch_box4_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoBox4>(t, p, b, E, s);
}
}  // namespace wg
