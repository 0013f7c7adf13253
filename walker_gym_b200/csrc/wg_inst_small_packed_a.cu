// wg_inst_small_packed_a.cu -- packed-state step kernels for smaller walker.py bodies.
#include "wg_launch.cuh"
namespace wg {
int launch_legacy_box_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoLegacyBox>(t, p, b, E, s);
}
int launch_test_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoTest>(t, p, b, E, s);
}
int launch_intrian_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    return launch_packed_flags<TopoIntrian>(t, p, b, E, s);
}
}  // namespace wg
