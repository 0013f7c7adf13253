// wg_inst_balance_packed.cu -- instantiates the packed-state step kernel for TopoBalance.
#include "wg_launch.cuh"
namespace wg {
int launch_balance_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoBalance, true>(t, p, b, E, s);      // also balance2 (mass 0.1) and balance3 (DingPoint)
}
}  // namespace wg
