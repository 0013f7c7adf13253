// wg_inst_small_packed_e.cu -- packed-state step kernel for a walker.py body.
#include "wg_launch.cuh"
namespace wg {
int launch_leg2_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoLeg2>(t, p, b, E, s);
}
}  // namespace wg
