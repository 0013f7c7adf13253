// wg_jit.cu -- run-time specialisation: a body without an ahead-of-time kernel gets the packed-state step kernel
// compiled for ITS spring graph (NVRTC, ~1 s, once per process and variant), so user-built creatures run the same
// register-resident code as the in-tree bodies instead of the run-time-topology kernel (2.5-3x slower on small bodies).
// NVRTC is loaded with dlopen on first use; if it is missing the caller keeps the run-time-topology kernel.
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "wg_launch.cuh"
#include "wg_kernels_multi.cuh"

namespace wg {

namespace {
typedef void* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    int (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : { "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so" }) {
            n.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (!n.h) return;
#define WG_SYM(F) *(void**)(&n.F) = dlsym(n.h, "nvrtc" #F); if (!n.F) return;
        WG_SYM(CreateProgram) WG_SYM(DestroyProgram) WG_SYM(AddNameExpression) WG_SYM(CompileProgram) WG_SYM(GetLoweredName)
        WG_SYM(GetCUBINSize) WG_SYM(GetCUBIN) WG_SYM(GetProgramLogSize) WG_SYM(GetProgramLog)
#undef WG_SYM
        n.ok = true;
    });
    return n;
}

std::string include_dir() {
    if (const char* e = getenv("WG_JIT_INCLUDE")) return e;
    Dl_info info;
    if (dladdr((void*)&include_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t slash = p.rfind('/');
        return (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/csrc";
    }
    return "csrc";
}

struct JitKernel { cudaKernel_t kernel = nullptr; int rc = WG_OK; std::string err; };
std::mutex g_mu;
std::map<std::string, JitKernel> g_cache;

std::string key_of(const wg_topology* t, int in3d, int obs_rm, int mm, int kind) {      // kind: 0 SoA, 1 packed, 2 packed T-steps-per-launch
    std::string k = std::string(kind == 3 ? "G" : kind == 2 ? "M" : kind ? "P" : "S") + std::to_string(t->n_mass) + "," + std::to_string(t->n_spring) + "," + std::to_string(t->n_muscle) + ":";
    for (int s = 0; s < t->n_spring; s++) k += std::to_string(t->si[s]) + "-" + std::to_string(t->sj[s]) + ",";
    k += "|" + std::to_string(in3d) + std::to_string(obs_rm) + std::to_string(mm);
    if (mm == 1) {                       // mass mode 3 bakes the mass pattern (which masses are 1 / equal) into the code
        k += "|";
        for (int n = 0; n < t->n_mass; n++) {
            int c = 0;
            if (t->mass[n] != 1.0) { c = n + 1; for (int q = 0; q < n; q++) if (t->mass[q] == t->mass[n]) { c = q + 1; break; } }
            k += std::to_string(c) + ",";
        }
    }
    return k;
}

JitKernel compile(const wg_topology* t, int in3d, int obs_rm, int mm, int kind) {
    const bool packed = kind != 0;
    JitKernel out;
    Nvrtc& nv = nvrtc();
    if (!nv.ok) { out.rc = WG_ERR_UNSUPPORTED; out.err = "libnvrtc not available"; return out; }
    std::string src = std::string(kind >= 2 ? "#include \"wg_kernels_multi.cuh\"\n" : "#include \"wg_kernels_packed.cuh\"\n") + "namespace wg { WG_STATIC_TOPO(TopoJit, 99, " + std::to_string(t->n_mass) + ", " +
                      std::to_string(t->n_spring) + ", " + std::to_string(t->n_muscle);
    for (int s = 0; s < t->n_spring; s++) src += ", " + std::to_string(t->si[s]) + "," + std::to_string(t->sj[s]);
    src += ")\n";
    // unit / small-integer masses: bake the mass pattern in as well (mass mode 3: unit masses need no division,
    // endpoints of equal mass share their quotients); class ids, not values -- the values stay kernel arguments
    std::string topo_name = "wg::TopoJit";
    if (mm == 1) {
        std::string cls;
        // simple encoding: class[n] = index of the first mass equal to mass[n] (+1), 0 for unit masses
        for (int n = 0; n < t->n_mass; n++) {
            int c = 0;
            if (t->mass[n] != 1.0) { c = n + 1; for (int q = 0; q < n; q++) if (t->mass[q] == t->mass[n]) { c = q + 1; break; } }
            cls += (n ? ", " : "") + std::to_string(c);
        }
        src += "struct TopoJitP : TopoJit {\n"
               "    __host__ __device__ static constexpr int cls(int n) { constexpr int c[" + std::to_string(t->n_mass) + "] = { " + cls + " }; return c[n]; }\n"
               "    __host__ __device__ static constexpr bool unit(int n) { return cls(n) == 0; }\n"
               "    __host__ __device__ static constexpr bool same(int i, int j) { return cls(i) != 0 && cls(i) == cls(j); }\n};\n";
        topo_name = "wg::TopoJitP";
    }
    src += "}\n";
    const std::string mm_s = std::to_string(mm == 1 ? 3 : mm), i3 = in3d ? "true" : "false";
    const std::string name = kind >= 2          // kind 3: the instance with the in-kernel action sources compiled in
        ? "&wg::step_multi_packed_kernel<" + topo_name + ", " + i3 + ", " + mm_s + ", wg::StepArgs<wg::kMaxMass, wg::kMaxSpring>, " +
          (kind == 3 ? "true" : "false") + ">"
        : std::string(packed ? "&wg::step_static_packed_kernel<" : "&wg::step_static_kernel<") + topo_name + ", " + i3 + ", " +
          std::to_string(obs_rm) + (packed ? ", " : ", 1, ") + mm_s + ", wg::StepArgs<wg::kMaxMass, wg::kMaxSpring>>";
    nvrtcProgram prog = nullptr;
    if (nv.CreateProgram(&prog, src.c_str(), "wg_jit.cu", 0, nullptr, nullptr) != 0) { out.rc = WG_ERR_CUDA; out.err = "nvrtcCreateProgram failed"; return out; }
    nv.AddNameExpression(prog, name.c_str());
    const std::string inc = "-I" + include_dir();
    const char* opts[] = { "--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "-lineinfo", inc.c_str() };
    const int crc = nv.CompileProgram(prog, 5, opts);
    if (crc != 0) {
        size_t n = 0;
        nv.GetProgramLogSize(prog, &n);
        std::vector<char> log(n + 1, 0);
        nv.GetProgramLog(prog, log.data());
        out.rc = WG_ERR_CUDA; out.err = std::string("nvrtc: ") + std::string(log.data()).substr(0, 400);
        nv.DestroyProgram(&prog);
        return out;
    }
    const char* lowered = nullptr;
    size_t n = 0;
    nv.GetLoweredName(prog, name.c_str(), &lowered);
    nv.GetCUBINSize(prog, &n);
    std::vector<char> cubin(n);
    nv.GetCUBIN(prog, cubin.data());
    cudaLibrary_t lib = nullptr;
    cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&out.kernel, lib, lowered ? lowered : "");
    if (e != cudaSuccess) { out.rc = WG_ERR_CUDA; out.err = std::string("loading the compiled kernel: ") + cudaGetErrorString(e); out.kernel = nullptr; }
    nv.DestroyProgram(&prog);
    return out;
}
}  // namespace

// bodies the run-time specialisation accepts: small enough for a register-resident kernel (packed state: the
// float4 layout pays up to 8 masses; SoA state: one thread per env holds up to 16 masses in registers, as the
// ahead-of-time insect / quad kernels do)
bool jit_eligible(const wg_topology* t) { return t->n_mass >= 1 && t->n_mass <= 8 && t->n_spring >= 1 && t->n_spring <= 16; }
bool jit_eligible_soa(const wg_topology* t) { return t->n_mass >= 1 && t->n_mass <= 16 && t->n_spring >= 1 && t->n_spring <= 32; }
bool jit_runtime_available() { return nvrtc().ok; }

static int mode_of(const wg_topology* t) {
    int mm = mass_mode(t);
    for (int n = 0; n < t->n_mass; n++) if (t->fixed[n]) mm = 2;
    return mm;
}

// compile (or fetch) the kernel for this body / variant; WG_OK or an error with the compiler log in the error string
int jit_prepare(const wg_topology* t, int in3d, int obs_layout, cudaKernel_t* kernel, int kind) {
    const int obs_rm = obs_layout == 0 ? 1 : 0, mm = mode_of(t);
    const std::string key = key_of(t, in3d ? 1 : 0, obs_rm, mm, kind);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) it = g_cache.emplace(key, compile(t, in3d ? 1 : 0, obs_rm, mm, kind)).first;
    if (it->second.rc != WG_OK) return fail(it->second.rc, "run-time specialisation failed: %s", it->second.err.c_str());
    if (kernel) *kernel = it->second.kernel;
    return WG_OK;
}

int launch_jit_packed(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    cudaKernel_t kernel = nullptr;
    int rc = jit_prepare(t, p->in3d, b->obs_layout, &kernel, true);
    if (rc != WG_OK) return rc;
    static thread_local StepArgs<kMaxMass, kMaxSpring> A;
    fill_args(A, t, p, b, E);
    const int D = 3 * (p->in3d ? 3 : 2) * t->n_mass + t->n_muscle;
    const bool rm = b->obs_layout == 0;
    const bool bulk = rm && gcd_c(D, 32) <= 2;                                 // mirrors the kernel's OBS_BULK
    const size_t smem = (rm && b->obs) ? sizeof(float) * kPackedBlock * (bulk ? D : (D | 1)) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFuncSetAttribute (jit): %s", cudaGetErrorString(e));
    }
    void* args[] = { &A };
    cudaError_t e = cudaLaunchKernel((const void*)kernel, dim3((unsigned)((E + kPackedBlock - 1) / kPackedBlock)), dim3(kPackedBlock),
                                     args, smem, s);
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "step kernel (jit) launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

// T env-steps per launch (wg_step_multi) for a body without an ahead-of-time kernel
int launch_jit_multi(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, int n_steps, int64_t act_stride,
                     cudaStream_t s) {
    cudaKernel_t kernel = nullptr;
    static thread_local ActionGen G;
    fill_gen(G, b);
    int rc = jit_prepare(t, p->in3d, 0, &kernel, G.mode ? 3 : 2);
    if (rc != WG_OK) return rc;
    static thread_local StepArgs<kMaxMass, kMaxSpring> A;
    fill_args(A, t, p, b, E);
    const int D = 3 * (p->in3d ? 3 : 2) * t->n_mass + t->n_muscle;
    const size_t smem = b->obs ? sizeof(float) * kMultiBlock * (gcd_c(D, 32) <= 2 ? D : (D | 1)) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFuncSetAttribute (jit): %s", cudaGetErrorString(e));
    }
    void* args[] = { &A, &n_steps, &act_stride, &G };
    cudaError_t e = cudaLaunchKernel((const void*)kernel, dim3((unsigned)((E + kMultiBlock - 1) / kMultiBlock)), dim3(kMultiBlock),
                                     args, smem, s);
    if (e != cudaSuccess) return fail(WG_ERR_CUDA, "step kernel (jit, multi) launch: %s", cudaGetErrorString(e));
    return WG_OK;
}

// SoA state: step_static_kernel<TopoJit, ..., EPT = 1, ...> compiled for this body (up to 16 masses)
int launch_jit_soa(const wg_topology* t, const wg_params* p, const wg_buffers* b, int64_t E, cudaStream_t s) {
    cudaKernel_t kernel = nullptr;
    int rc = jit_prepare(t, p->in3d, b->obs_layout, &kernel, false);
    if (rc != WG_OK) return rc;
    static thread_local StepArgs<kMaxMass, kMaxSpring> A;
    fill_args(A, t, p, b, E);
    const int D = 3 * (p->in3d ? 3 : 2) * t->n_mass + t->n_muscle;
    const bool rm = b->obs_layout == 0;
    const bool bulk = rm && gcd_c(D, 32) <= 2;                                 // mirrors launch_static
    const size_t smem = (rm && b->obs) ? sizeof(float) * kBlock * (bulk ? D : (D | 1)) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(WG_ERR_CUDA, "cudaFuncSetAttribute (jit): %s", cudaGetErrorString(e));
    }
    void* args[] = { &A };
    cudaError_t e = cudaLaunchKernel((const void*)kernel, dim3((unsigned)((E + kBlock - 1) / kBlock)), dim3(kBlock), args, smem, s);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(WG_ERR_CUDA, "step kernel (jit, SoA) launch: %s", cudaGetErrorString(e)); }
    return WG_OK;
}

}  // namespace wg
