// wg_inst_policy.cu -- rollout-side kernels: the fused MLP policy (wg_policy_act) and GAE (wg_gae).
#include "wg_launch.cuh"
#include "wg_policy.cuh"
namespace wg {

#ifndef WG_POLICY_MT
#define WG_POLICY_MT 1
#endif
template <int KT1, bool SPLIT>
static int launch_policy_t(const PolicyArgs& A, cudaStream_t s) {
    constexpr int MT = WG_POLICY_MT;
    auto kern = policy_act_kernel<KT1, SPLIT, MT>;
    const size_t smem = sizeof(uint32_t) * (size_t)PolicySmem<KT1>::words(SPLIT);
    static thread_local int cached_dev = -1, n_sm = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int64_t n_tiles = (A.E + kPolBlock - 1) / kPolBlock;
    const int64_t resident = (int64_t)n_sm * 2;                 // persistent: weights are staged once per CTA
    kern<<<(unsigned)(n_tiles < resident ? n_tiles : resident), 32 * (kPolBlock / (16 * MT)), smem, s>>>(A);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "policy kernel launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

int launch_policy(const PolicyArgs& A, int precision, cudaStream_t s) {
    const int kt = (A.D + 7) / 8;
#define WG_POL(KT) (precision == 0 ? launch_policy_t<KT, true>(A, s) : launch_policy_t<KT, false>(A, s))
    if (kt <= 3) return WG_POL(3);
    if (kt <= 4) return WG_POL(4);
    if (kt <= 5) return WG_POL(5);
    return WG_POL(8);
#undef WG_POL
}

int launch_gae(const float* rewards, const float* values, const uint8_t* dones, float* adv, float* ret, int T, int64_t E,
               float gamma, float lam, float clip, cudaStream_t s) {
    gae_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(rewards, values, dones, adv, ret, T, E, gamma, lam, clip);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "gae kernel launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

}  // namespace wg
