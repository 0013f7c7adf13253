// timeline of the warp-specialised tcgen05 policy kernel: phase timestamps of one thread per role of CTA 0
#define WG_WS_TRACE
#include <cstdio>
#include <vector>
#include "../../walker_gym_b200/csrc/wg_kernels.cuh"
#include "../../walker_gym_b200/csrc/wg_policy_ws.cuh"
namespace wg { int fail(int, const char*, ...) { return 1; } }
using namespace wg;
template <bool SPLIT> void run(const PolicyArgs& A) {
    auto kern = policy_act_ws_kernel<40, SPLIT>;
    const size_t smem = WsSmem<40>::bytes(A.D);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int* flag; cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        kern<<<148, kWsThreads, smem>>>(A, flag, NoStep{});
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("SPLIT %d: %.1f us (%s)\n", (int)SPLIT, ms * 1e3, cudaGetErrorString(e));
    }
    std::vector<long long> tr(4 * 16 * 8);
    cudaMemcpyFromSymbol(tr.data(), g_ws_trace, sizeof(long long) * tr.size());
    const long long t0 = tr[0];
    const char* roles[4] = {"P  ", "MMA", "E1 ", "E2 "};
    const int nslot[4] = {5, 6, 6, 6};
    const char* names[4][8] = {{"top", "obs_full", "a1_free", "stored", "gsync"}, {"top", "a1_ready", "L1_issued", "e1_done", "d2_free", "L2_issued"},
                               {"top", "d1_full", "loaded", "chunk0", "h_free", "done"}, {"top", "d2_full", "loaded", "tanh_done", "hx_free", "done"}};
    for (int it = 0; it < 15; it++)
        for (int r = 0; r < 4; r++) {
            printf(" tile %2d %s:", it, roles[r]);
            for (int s = 0; s < nslot[r]; s++) printf(" %s@%lld", names[r][s], tr[(r * 16 + it) * 8 + s] - t0);
            printf("\n");
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
    run<true>(A); run<false>(A);
    return 0;
}
