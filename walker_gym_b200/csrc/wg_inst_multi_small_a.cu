// wg_inst_multi_small_a.cu -- instantiates the T-steps-per-launch kernel for walker.py box, test and intrian.
#include "wg_launch.cuh"
#include "wg_kernels_multi.cuh"
namespace wg {
int launch_legacy_box_multi(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int n_steps, int64_t as, cudaStream_t s) {
    return launch_multi_flags<TopoLegacyBox>(t, p, b, E, n_steps, as, s);
}
int launch_test_multi(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int n_steps, int64_t as, cudaStream_t s) {
    return launch_multi_flags<TopoTest>(t, p, b, E, n_steps, as, s);
}
int launch_intrian_multi(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int n_steps, int64_t as, cudaStream_t s) {
    return launch_multi_flags<TopoIntrian>(t, p, b, E, n_steps, as, s);
}
}  // namespace wg
