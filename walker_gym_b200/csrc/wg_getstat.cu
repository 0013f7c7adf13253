// wg_getstat.cu -- Creature.getstat (gym/optimized_walker.py:129-162) with its non-default options for every env of a
// batch: scale factors pk / vk / ak / mk, midform (positions relative to the centroid or absolute), conmid (append the
// centroid).  PhysicsEnv._get_observation uses the defaults and is produced by the step kernel itself; this kernel is the
// accessor for callers that want another view of the same state.  One thread per env, not on the hot path.
#include "wg_launch.cuh"

namespace wg {

struct GetstatArgs {
    const float* pos; const float* vel; const float* mx; const float* old_a; const float* state_packed;
    const float* obs_default;     // old_a source when old_a is null: the acceleration entries of the step kernel's observation
    float* out;
    int64_t E;
    int32_t N, M, d, d_env, obs_layout_default, out_layout, midform, conmid;
    float pk, vk, ak, mk;
    ConstDiv ndiv;
};

__global__ void __launch_bounds__(128) getstat_kernel(const __grid_constant__ GetstatArgs A) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.E) return;
    const int N = A.N, M = A.M, d = A.d;
    const int64_t E = A.E;
    const int r4 = (6 * N + M + 2 + 3) / 4;
    const int D_def = 3 * A.d_env * N + M;
    const int D = 3 * d * N + M + (A.conmid ? 3 : 0);
    auto pos = [&](int n, int c) { return A.state_packed ? A.state_packed[packed_index(e, n * 3 + c, r4)] : A.pos[(int64_t)(n * 3 + c) * E + e]; };
    auto vel = [&](int n, int c) { return A.state_packed ? A.state_packed[packed_index(e, 3 * N + n * 3 + c, r4)] : A.vel[(int64_t)(n * 3 + c) * E + e]; };
    auto mxv = [&](int m) { return A.state_packed ? A.state_packed[packed_index(e, 6 * N + m, r4)] : A.mx[(int64_t)m * E + e]; };
    auto olda = [&](int n, int c) -> float {
        if (A.old_a) return A.old_a[(int64_t)(n * 3 + c) * E + e];
        const int k = n * 3 * A.d_env + 2 * A.d_env + c;          // entry of the default observation (c < d_env checked on the host)
        return A.obs_layout_default == 0 ? A.obs_default[e * D_def + k] : A.obs_default[(int64_t)k * E + e];
    };
    auto emit = [&](int k, float v) { if (A.out_layout == 0) A.out[e * D + k] = v; else A.out[(int64_t)k * E + e] = v; };
    float mid[3] = { 0.0f, 0.0f, 0.0f };                          // mid = np.zeros(3, float32)
    if (A.midform) {                                              // for i in phys: mid += i.pos;  mid /= len(phys)
        for (int n = 0; n < N; n++) { mid[0] = mid[0] + pos(n, 0); mid[1] = mid[1] + pos(n, 1); mid[2] = mid[2] + pos(n, 2); }
        for (int c = 0; c < 3; c++) mid[c] = div_const(mid[c], A.ndiv.m, A.ndiv.r, A.ndiv.kind);
    }
    int k = 0;
    for (int n = 0; n < N; n++) {
        for (int c = 0; c < d; c++) emit(k++, (A.midform ? pos(n, c) - mid[c] : pos(n, c)) * A.pk);
        for (int c = 0; c < d; c++) emit(k++, vel(n, c) * A.vk);
        for (int c = 0; c < d; c++) emit(k++, olda(n, c) * A.ak);
    }
    if (A.conmid) for (int c = 0; c < 3; c++) emit(k++, mid[c]);
    for (int m = 0; m < M; m++) emit(k++, mxv(m) * A.mk);
}

}  // namespace wg

using namespace wg;

extern "C" int wg_getstat(const wg_topology* topo, const wg_buffers* buf, const float* obs_default, int32_t env_in3d,
                          int32_t in3d, float pk, float vk, float ak, float mk, int32_t midform, int32_t conmid,
                          float* out, int32_t out_layout, int64_t n_env, void* cuda_stream) {
    if (!topo || !buf || !out) return fail(WG_ERR_BAD_ARG, "wg_getstat: null argument%s");
    if (topo->n_mass < 1 || topo->n_mass > WG_MAX_MASS || topo->n_muscle < 0 || topo->n_muscle > WG_MAX_SPRING)
        return fail(WG_ERR_BAD_ARG, "wg_getstat: topology out of range%s");
    if (!buf->state_packed && (!buf->pos || !buf->vel || (topo->n_muscle > 0 && !buf->mx)))
        return fail(WG_ERR_BAD_ARG, "wg_getstat: pos / vel / mx (or state_packed) must be set%s");
    if (!buf->old_a && !obs_default) return fail(WG_ERR_BAD_ARG, "wg_getstat: needs buf->old_a or the step kernel's observation as the old_a source%s");
    if (!buf->old_a && in3d && !env_in3d)
        return fail(WG_ERR_BAD_ARG, "wg_getstat: a 2-D env's observation has no z accelerations; keep old_a (buf->old_a) for in3d = 1%s");
    if (out_layout != 0 && out_layout != 1) return fail(WG_ERR_BAD_ARG, "wg_getstat: out_layout must be 0 or 1%s");
    if (n_env < 0 || n_env > ((int64_t)1 << 31) - 1) return fail(WG_ERR_BAD_ARG, "n_env out of range%s");
    if (n_env == 0) return WG_OK;
    GetstatArgs A;
    A.pos = buf->pos; A.vel = buf->vel; A.mx = buf->mx; A.old_a = buf->old_a; A.state_packed = buf->state_packed;
    A.obs_default = obs_default; A.out = out; A.E = n_env;
    A.N = topo->n_mass; A.M = topo->n_muscle; A.d = in3d ? 3 : 2; A.d_env = env_in3d ? 3 : 2;
    A.obs_layout_default = buf->obs_layout; A.out_layout = out_layout; A.midform = midform ? 1 : 0; A.conmid = conmid ? 1 : 0;
    A.pk = pk; A.vk = vk; A.ak = ak; A.mk = mk;
    A.ndiv = make_const_div((float)topo->n_mass);
    getstat_kernel<<<(unsigned)((n_env + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "getstat kernel launch: %s", cudaGetErrorString(e));
    return WG_OK;
}
