// wg_inst_units_link.cu -- CONNECTED bodies made of identical units chained by link bones (BASELINE config 4's enlarged
// morphology as one creature: quad_balance_chain): step_units_kernel with LA / LB set -- one lane per unit, the unit's
// physics register-resident for all substeps, the link bones' far endpoints fetched from the neighbouring lanes with
// warp shuffles every substep.
#include "wg_launch.cuh"
namespace wg {

template <class U>
static int unit_spring(int M_total, int u, int ls) {
    return ls < U::M ? U::M * u + ls : M_total + (U::S - U::M) * u + (ls - U::M);
}

// t is R copies of U (same masses and spring constants, no DingPoints) laid out unit after unit, followed in the
// skeleton list by R-1 link bones, link u = (unit u's mass LA, unit u+1's mass LB), all with the same constants
template <class U, int LA, int LB>
static bool chain_match(const wg_topology* t, int R) {
    if (t->n_mass != R * U::N || t->n_spring != R * U::S + (R - 1) || t->n_muscle != R * U::M) return false;
    for (int u = 0; u < R; u++) {
        for (int n = 0; n < U::N; n++)
            if (t->fixed[U::N * u + n] || t->mass[U::N * u + n] != t->mass[n]) return false;
        for (int ls = 0; ls < U::S; ls++) {
            const int g = unit_spring<U>(t->n_muscle, u, ls), g0 = unit_spring<U>(t->n_muscle, 0, ls);
            if (t->si[g] != U::N * u + U::si(ls) || t->sj[g] != U::N * u + U::sj(ls)) return false;
            if (t->sk[g] != t->sk[g0] || t->sdamp[g] != t->sdamp[g0] || t->srest[g] != t->srest[g0]) return false;
            if (ls < U::M && (t->mlo[g] != t->mlo[g0] || t->mhi[g] != t->mhi[g0])) return false;
        }
    }
    const int l0 = R * U::S;
    for (int u = 0; u + 1 < R; u++) {
        const int g = l0 + u;
        if (t->si[g] != U::N * u + LA || t->sj[g] != U::N * (u + 1) + LB) return false;
        if (t->sk[g] != t->sk[l0] || t->sdamp[g] != t->sdamp[l0] || t->srest[g] != t->srest[l0]) return false;
    }
    return true;
}

template <class U, bool IN3D, int R, bool ROWMAJOR, int MM, int LA, int LB>
static int launch_chain_t(const wg_topology* t, const wg_topology* ut, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    static thread_local UnitsArgs<U, MM> UA;
    fill_args(UA.P.A, t, p, b, E);
    static thread_local StepArgs<U::N, U::S> tmp;
    fill_args(tmp, ut, p, b, E);
    UA.ubv = tmp.bv;
    const int l0 = R * U::S;
    UA.link_k = t->sk[l0]; UA.link_damp = t->sdamp[l0]; UA.link_rest = t->srest[l0];
    constexpr int KB = 256, EB = KB / R;
    constexpr int N = R * U::N, D = 3 * (IN3D ? 3 : 2) * N + R * U::M, SCR = 5 * N + 4;
    const size_t smem = sizeof(float) * ((size_t)((EB * SCR + 31) / 32) * 32 + (ROWMAJOR ? (size_t)EB * D : 0));
    auto kern = step_units_kernel<U, IN3D, R, ROWMAJOR, MM, KB, LA, LB>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<(unsigned)((E + EB - 1) / EB), KB, smem, s>>>(UA);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "linked-units step kernel launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

// number of Balance units (2, 4 or 8) chained shoulder to shoulder (unit u's mass 1 -- unit u+1's mass 0), or 0
int balance_chain_units(const wg_topology* t) {
    for (int R : { 2, 4, 8 })
        if (chain_match<TopoBalance, 1, 0>(t, R) && mass_mode(t) <= 1) return R;
    return 0;
}

template <int R>
static int launch_balance_chain_r(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    using U = TopoBalance;
    wg_topology ut;
    memset(&ut, 0, sizeof(ut));
    ut.n_mass = U::N; ut.n_spring = U::S; ut.n_muscle = U::M;
    for (int n = 0; n < U::N; n++) { ut.mass[n] = t->mass[n]; for (int c = 0; c < 3; c++) ut.tmpl_pos[n * 3 + c] = t->tmpl_pos[n * 3 + c]; }
    for (int ls = 0; ls < U::S; ls++) {
        const int g = unit_spring<U>(t->n_muscle, 0, ls);
        ut.si[ls] = U::si(ls); ut.sj[ls] = U::sj(ls);
        ut.sk[ls] = t->sk[g]; ut.sdamp[ls] = t->sdamp[g]; ut.srest[ls] = t->srest[g]; ut.mlo[ls] = t->mlo[g]; ut.mhi[ls] = t->mhi[g];
    }
    const int mm = mass_mode(&ut);
    const bool rm = b->obs_layout == 0;
    const bool v0 = mm == 1 && ut.mass[2] == 1.0 && ut.mass[0] == ut.mass[1] && ut.mass[0] != 1.0 && ut.mass[3] != 1.0;
#define WG_CH(UT, MMV) (p->in3d ? (rm ? launch_chain_t<UT, true, R, true, MMV, 1, 0>(t, &ut, p, b, E, s) : launch_chain_t<UT, true, R, false, MMV, 1, 0>(t, &ut, p, b, E, s)) \
                                : (rm ? launch_chain_t<UT, false, R, true, MMV, 1, 0>(t, &ut, p, b, E, s) : launch_chain_t<UT, false, R, false, MMV, 1, 0>(t, &ut, p, b, E, s)))
    if (v0) return WG_CH(TopoBalanceV0, 3);
    if (mm == 0) return WG_CH(TopoBalance, 0);
    return WG_CH(TopoBalance, 1);
#undef WG_CH
}

int launch_balance_chain(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int R, cudaStream_t s) {
    switch (R) {
        case 2: return launch_balance_chain_r<2>(t, p, b, E, s);
        case 4: return launch_balance_chain_r<4>(t, p, b, E, s);
        case 8: return launch_balance_chain_r<8>(t, p, b, E, s);
        default: return fail(WG_ERR_BAD_ARG, "unit count must be 2, 4 or 8%s");
    }
}

}  // namespace wg
