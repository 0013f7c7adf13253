// wg_inst_insect.cu -- instantiates the register-resident step kernel for TopoInsect.
#include "wg_launch.cuh"
namespace wg {
int launch_insect(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int ept, cudaStream_t s) {
    (void)ept;
    return launch_static_flags<TopoInsect, 1>(t, p, b, E, s);
}
}  // namespace wg
