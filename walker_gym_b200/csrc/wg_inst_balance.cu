// wg_inst_balance.cu -- instantiates the register-resident step kernel for TopoBalance.
#include "wg_launch.cuh"
namespace wg {
int launch_balance(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int ept, cudaStream_t s) {
    (void)ept;
    return launch_static_flags<TopoBalance, 1>(t, p, b, E, s);
}
}  // namespace wg
