// wg_inst_box.cu -- instantiates the register-resident step kernel for TopoBox.
#include "wg_launch.cuh"
namespace wg {
int launch_box(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int ept, cudaStream_t s) {
    (void)ept;
    return launch_static_flags<TopoBox, 1>(t, p, b, E, s);
}
}  // namespace wg
