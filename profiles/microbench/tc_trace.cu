// timeline of the tcgen05 policy kernel: phase timestamps of CTA 0, thread 0
#define WG_TC_TRACE
#include <cstdio>
#include <vector>
#include "../../walker_gym_b200/csrc/wg_kernels.cuh"
#include "../../walker_gym_b200/csrc/wg_policy_tc.cuh"
namespace wg { int fail(int, const char*, ...) { return 1; } }
using namespace wg;
template <bool SPLIT> void run(const PolicyArgs& A, int ctas_per_sm) {
    auto kern = policy_act_tc_kernel<40, SPLIT>;
    const size_t smem = TcSmem<40>::bytes(A.D);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int* flag; cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        kern<<<148 * ctas_per_sm, kTcThreads, smem>>>(A, flag);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("SPLIT %d ctas/SM %d: %.1f us (%s)\n", (int)SPLIT, ctas_per_sm, ms * 1e3, cudaGetErrorString(e));
    }
    std::vector<long long> tr(16 * 64);
    cudaMemcpyFromSymbol(tr.data(), g_tc_trace, sizeof(long long) * 16 * 64);
    const char* names[15] = {"top", "tma_ok", "obs_done", "sync", "mma1_issued", "mma1_done", "ep1_done", "sync", "mma2_issued", "noise+mma2_done", "ep2_done", "sync", "heads_computed", "ep2_tanh_done", "heads_done"};
    for (int it = 0; it < 8; it++) {
        printf(" tile %d:", it);
        long long prev = it ? tr[(it - 1) * 16 + 14] : tr[0];
        const int order[15] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 10, 11, 12, 14};
        for (int si = 0; si < 15; si++) { const int s = order[si]; printf(" %s+%lld", names[s], tr[it * 16 + s] - prev); prev = tr[it * 16 + s]; }
        printf("  | total %lld\n", tr[it * 16 + 14] - tr[it * 16]);
    }
}
int main() {
    const int D = 38, M = 2; const int64_t E = 1 << 18;
    std::vector<float> h(64 * D + 64 + 64 * 64 + 64 + M * 64 + M + 64 + 1 + M);
    for (size_t i = 0; i < h.size(); i++) h[i] = 0.01f * (float)((int)(i * 2654435761u >> 20) % 21 - 10);
    float* w; cudaMalloc(&w, h.size() * 4); cudaMemcpy(w, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    std::vector<float> ho(E * D);
    for (size_t i = 0; i < ho.size(); i++) ho[i] = 0.1f * (float)((int)(i * 2246822519u >> 18) % 41 - 20);
    float* obs; cudaMalloc(&obs, ho.size() * 4); cudaMemcpy(obs, ho.data(), ho.size() * 4, cudaMemcpyHostToDevice);
    float *act, *logp, *val; cudaMalloc(&act, E * M * 4); cudaMalloc(&logp, E * 4); cudaMalloc(&val, E * 4);
    PolicyArgs A{};
    float* p = w;
    A.w1 = p; p += 64 * D; A.b1 = p; p += 64; A.w2 = p; p += 64 * 64; A.b2 = p; p += 64; A.w_mu = p; p += M * 64; A.b_mu = p; p += M;
    A.w_v = p; p += 64; A.b_v = p; p += 1; A.log_std = p;
    A.obs = obs; A.action = act; A.logp = logp; A.value = val; A.mean = nullptr; A.step_counter = nullptr;
    A.E = E; A.D = D; A.M = M; A.act_layout = 0; A.sample = 1; A.obs_layout = 0; A.obs_scale = 1.0f; A.obs_clip = 10.0f;
    A.seed_lo = 1; A.seed_hi = 2; A.step_index = 3; A.env_offset = 0;
    run<true>(A, 2); run<true>(A, 1); run<false>(A, 2); run<false>(A, 1);
    return 0;
}
