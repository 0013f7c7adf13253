// wg_inst_small_packed_f.cu -- packed-state step kernel for a walker.py body.
#include "wg_launch.cuh"
namespace wg {
int launch_leg_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoLeg>(t, p, b, E, s);
}
}  // namespace wg
