// wg_inst_multi_balance.cu -- instantiates the T-steps-per-launch kernel for the Balance spring graph.
#include "wg_launch.cuh"
#include "wg_kernels_multi.cuh"
namespace wg {
int launch_balance_multi(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int n_steps, int64_t as, cudaStream_t s) {
    bool general = mass_mode(t) == 2;                 // arbitrary masses (full IEEE division) or DingPoints: mass mode 2
    for (int n = 0; n < t->n_mass; n++) if (t->fixed[n]) general = true;
    if (general)
        return p->in3d ? launch_multi_packed<TopoBalance, true, 2>(t, p, b, E, n_steps, as, s)
                       : launch_multi_packed<TopoBalance, false, 2>(t, p, b, E, n_steps, as, s);
    if (mass_mode(t) == 1 && t->mass[2] == 1.0 && t->mass[0] == t->mass[1] && t->mass[0] != 1.0 && t->mass[3] != 1.0)   // Balance-v0
        return p->in3d ? launch_multi_packed<TopoBalanceV0, true, 3>(t, p, b, E, n_steps, as, s)
                       : launch_multi_packed<TopoBalanceV0, false, 3>(t, p, b, E, n_steps, as, s);
    return launch_multi_flags<TopoBalance>(t, p, b, E, n_steps, as, s);
}
}  // namespace wg
