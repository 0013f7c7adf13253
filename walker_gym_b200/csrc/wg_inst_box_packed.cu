// wg_inst_box_packed.cu -- instantiates the packed-state step kernel for TopoBox.
#include "wg_launch.cuh"
namespace wg {
int launch_box_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoBox>(t, p, b, E, s);
}
}  // namespace wg
