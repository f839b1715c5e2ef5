// Host side of the specialised kernels: assembles the translation unit of one homotopy (hc_jitgen.h writes its
// evaluate / Jacobian / Taylor code), compiles it for sm_100a with NVRTC, loads the cubin (cudaLibraryLoadData) and
// keeps it in a two-level cache (per process by source hash; on disk next to the library, so that a second process
// -- the next rank, the next solve() of the same system -- skips the compiler).
//
// This is what `compile = true` is in the reference (src/model_kit/compiled_system_homotopy.jl:178-243: Julia
// generates and compiles the straight-line code of a system on first use); like there, small batches keep using
// the interpreter (HC_B200_JIT_MIN_PATHS) because compiling costs seconds.
//
// With HC_HOST_SIM (tests/host_sim) the very same unit is compiled by g++ into a shared object and run
// sequentially on the CPU -- test infrastructure for the generated code, never reachable from libhc_b200.so.
#pragma once
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "hc_jitgen.h"
#include "hc_kernel.h"

#ifndef HC_HOST_SIM
#include <cuda_runtime.h>
#include <nvrtc.h>
#endif

namespace hc {
namespace jit {

struct Module {
    int n = 0, block = 128, lu_smem = 0;
    size_t hot = 0, cold = 0, slab = 0;  // lane state: bytes in the hot / cold part, local slab of the kernel
    double compile_ms = 0;
    bool from_cache = false;
    size_t cubin_bytes = 0;
#ifndef HC_HOST_SIM
    cudaLibrary_t lib = nullptr;
    cudaKernel_t track = nullptr, hook = nullptr;
    bool attr_set = false;
#else
    void* dl = nullptr;
    void (*track)(const KArgs*, unsigned char*, unsigned char*) = nullptr;
    void (*hook)(const KArgs*, unsigned char*, unsigned char*, int, int, const cx*, cx, const double*, cx*, cx*) = nullptr;
#endif
};

inline uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ULL) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
    return h;
}

inline std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::string("specialised kernel: cannot read ") + path;
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// directory of the device headers: next to the shared library (the package ships csrc/), or HC_B200_CSRC
inline std::string csrc_dir() {
    if (const char* e = getenv("HC_B200_CSRC")) return e;
#ifdef HC_CSRC_DIR
    return HC_CSRC_DIR;
#else
    Dl_info info;
    if (dladdr((void*)&fnv1a, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.rfind('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/csrc";
    }
    return "csrc";
#endif
}

inline const std::vector<std::string>& header_names() {
    static const std::vector<std::string> names = {"hc_common.h", "hc_coop.h", "hc_tape.h", "hc_path.h", "hc_lane.h", "hc_kernel.h", "hc_jit_unit.h"};
    return names;
}

#ifndef HC_HOST_SIM
struct Nvrtc {
    void* dl = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcGetErrorString) errstr = nullptr;
    decltype(&nvrtcVersion) version = nullptr;
    static Nvrtc& get() {
        static Nvrtc N;
        if (N.dl) return N;
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) {
            N.dl = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (N.dl) break;
        }
        if (!N.dl) throw std::string("specialised kernel: libnvrtc not found (") + dlerror() + ")";
#define HC_SYM(field, sym) N.field = (decltype(N.field))dlsym(N.dl, #sym); if (!N.field) throw std::string("libnvrtc lacks " #sym)
        HC_SYM(create, nvrtcCreateProgram); HC_SYM(compile, nvrtcCompileProgram); HC_SYM(destroy, nvrtcDestroyProgram);
        HC_SYM(cubin_size, nvrtcGetCUBINSize); HC_SYM(cubin, nvrtcGetCUBIN); HC_SYM(log_size, nvrtcGetProgramLogSize);
        HC_SYM(log, nvrtcGetProgramLog); HC_SYM(errstr, nvrtcGetErrorString); HC_SYM(version, nvrtcVersion);
#undef HC_SYM
        return N;
    }
};

// source -> cubin (sm_100a); `gen` is the generated include
inline std::string nvrtc_compile(const std::string& unit, const std::string& gen, const std::vector<std::string>& hdr_text) {
    Nvrtc& N = Nvrtc::get();
    std::vector<const char*> names, texts;
    const std::vector<std::string>& hn = header_names();
    for (size_t i = 0; i < hn.size(); ++i) { names.push_back(hn[i].c_str()); texts.push_back(hdr_text[i].c_str()); }
    names.push_back("hc_jit_gen.inc"); texts.push_back(gen.c_str());
    nvrtcProgram prog;
    nvrtcResult r = N.create(&prog, unit.c_str(), "hc_jit_unit.cu", (int)names.size(), texts.data(), names.data());
    if (r != NVRTC_SUCCESS) throw std::string("nvrtcCreateProgram: ") + N.errstr(r);
    // no FMA contraction by default (HC_B200_FMAD=1 turns it on: + 4 ... 7 % paths/s, results then differ from the
    // CPU build of the same code in the last bits, and rounding-sensitive paths may end differently)
    const bool fmad = getenv("HC_B200_FMAD") && atoi(getenv("HC_B200_FMAD")) != 0;
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DHC_JIT=1", fmad ? "--fmad=true" : "--fmad=false"};
    std::string extra;
    if (const char* e = getenv("HC_B200_JIT_FLAGS")) extra = e;
    std::vector<std::string> extra_opts;
    { std::stringstream ss(extra); std::string tok; while (ss >> tok) extra_opts.push_back(tok); }
    for (const std::string& o : extra_opts) opts.push_back(o.c_str());
    r = N.compile(prog, (int)opts.size(), opts.data());
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        N.log_size(prog, &ls);
        std::string log(ls, '\0');
        if (ls) N.log(prog, &log[0]);
        N.destroy(&prog);
        if (log.size() > 4000) log.resize(4000);
        throw std::string("specialised kernel failed to compile: ") + N.errstr(r) + "\n" + log;
    }
    size_t sz = 0;
    N.cubin_size(prog, &sz);
    std::string cubin(sz, '\0');
    N.cubin(prog, &cubin[0]);
    N.destroy(&prog);
    return cubin;
}
#endif

inline double wall_ms() {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

inline std::string cache_dir() {
    if (const char* e = getenv("HC_B200_JIT_CACHE")) return e;
#ifdef HC_HOST_SIM
    return "/tmp/hc_b200_simcache";
#else
    std::string d = csrc_dir();
    const size_t k = d.rfind('/');
    return (k == std::string::npos ? std::string(".") : d.substr(0, k)) + "/_jitcache";
#endif
}

// Builds (or fetches) the specialised module of one homotopy.  load = false: generate + compile only (no CUDA device
// needed: NVRTC targets sm_100a offline); the cubin lands in the disk cache.
inline std::shared_ptr<Module> build_module(const GenInput& in, int P, int tape_cx, int block, bool load = true, int sync = 1, int lu_smem = 0) {
    static std::map<uint64_t, std::shared_ptr<Module>> cache;
    const double t0 = wall_ms();
    PathMem<0> dummy;
    const SlabSizes ss = carve(dummy, in.n, P, tape_cx, nullptr, nullptr, 1 | (lu_smem ? 2 : 0));
    const size_t slab = (ss.hot + ss.cold + 255) & ~(size_t)255;
    const std::string gen = generate_members(in);
    std::string unit;
    unit += "#define HC_JIT_N " + std::to_string(in.n) + "\n";
    unit += "#define HC_JIT_SLAB " + std::to_string(slab) + "\n";
    unit += "#define HC_JIT_BLOCK " + std::to_string(block) + "\n";
    unit += "#define HC_JIT_SYNC " + std::to_string(sync) + "\n";
#ifndef HC_HOST_SIM
    unit += "#define HC_JIT_LU_SMEM " + std::to_string(lu_smem) + "\n";
    unit += std::string("#define HC_JIT_LU_UNROLL_MAX_N ") + (getenv("HC_B200_JIT_LU_UNROLL_MAX_N") ? getenv("HC_B200_JIT_LU_UNROLL_MAX_N") : "0") + "\n";
    unit += std::string("#define HC_JIT_PREFETCH ") + (getenv("HC_B200_JIT_PREFETCH") ? getenv("HC_B200_JIT_PREFETCH") : "0") + "\n";
#endif
    unit += "#define HC_JIT_GEN \"hc_jit_gen.inc\"\n";
    unit += "#include \"hc_jit_unit.h\"\n";
    const std::string dir = csrc_dir();
    std::vector<std::string> hdr_text;
    uint64_t h = fnv1a(unit);
    h = fnv1a(gen, h);
    for (const std::string& nme : header_names()) { hdr_text.push_back(read_file(dir + "/" + nme)); h = fnv1a(hdr_text.back(), h); }
    if (const char* e = getenv("HC_B200_JIT_FLAGS")) h = fnv1a(e, h);
    if (const char* e = getenv("HC_B200_FMAD")) h = fnv1a(std::string("fmad") + e, h);
#ifdef HC_HOST_SIM
    h = fnv1a("host-sim", h);
#endif
    auto it = cache.find(h);
    if (it != cache.end() && load) return it->second;
    auto M = std::make_shared<Module>();
    M->n = in.n; M->block = block; M->hot = ss.hot; M->cold = ss.cold; M->slab = slab; M->lu_smem = lu_smem;
    char hex[32];
    snprintf(hex, sizeof hex, "%016llx", (unsigned long long)h);
    const std::string cdir = cache_dir();
    mkdir(cdir.c_str(), 0777);
    if (getenv("HC_B200_JIT_DUMP")) {
        std::ofstream(std::string(getenv("HC_B200_JIT_DUMP")) + "/hc_jit_" + hex + "_gen.inc") << gen;
        std::ofstream(std::string(getenv("HC_B200_JIT_DUMP")) + "/hc_jit_" + hex + "_unit.cu") << unit;
    }
#ifndef HC_HOST_SIM
    const std::string cpath = cdir + "/" + hex + ".cubin";
    std::string cubin;
    {
        std::ifstream f(cpath, std::ios::binary);
        if (f && !getenv("HC_B200_JIT_NOCACHE")) { std::stringstream s; s << f.rdbuf(); cubin = s.str(); M->from_cache = !cubin.empty(); }
    }
    if (cubin.empty()) {
        cubin = nvrtc_compile(unit, gen, hdr_text);
        const std::string tmp = cpath + "." + std::to_string((long)getpid()) + ".tmp";
        std::ofstream o(tmp, std::ios::binary);
        if (o) { o.write(cubin.data(), (std::streamsize)cubin.size()); o.close(); rename(tmp.c_str(), cpath.c_str()); }
    }
    M->cubin_bytes = cubin.size();
    if (!load) { M->compile_ms = wall_ms() - t0; return M; }
    cudaError_t e = cudaLibraryLoadData(&M->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) throw std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e);
    e = cudaLibraryGetKernel(&M->track, M->lib, "hc_jit_track");
    if (e != cudaSuccess) throw std::string("cudaLibraryGetKernel(hc_jit_track): ") + cudaGetErrorString(e);
    e = cudaLibraryGetKernel(&M->hook, M->lib, "hc_jit_hook");
    if (e != cudaSuccess) throw std::string("cudaLibraryGetKernel(hc_jit_hook): ") + cudaGetErrorString(e);
#else
    const std::string so = cdir + "/" + hex + ".so";
    struct stat st;
    if (stat(so.c_str(), &st) != 0 || getenv("HC_B200_JIT_NOCACHE")) {
        const std::string base = cdir + "/" + hex + "." + std::to_string((long)getpid());
        mkdir(base.c_str(), 0777);
        std::ofstream(base + "/hc_jit_gen.inc") << gen;
        std::ofstream(base + "/unit.cpp") << unit;
        const std::string cmd = "g++ -std=c++17 -O2 -fPIC -shared -DHC_HOST_SIM -DHC_JIT=1 -Wno-unused-function -Wno-unknown-pragmas -I" + base + " -I" + dir +
                                " -o " + base + "/unit.so " + base + "/unit.cpp 2> " + base + "/log.txt";
        if (system(cmd.c_str()) != 0) throw std::string("specialised unit failed to compile (g++):\n") + read_file(base + "/log.txt").substr(0, 4000);
        rename((base + "/unit.so").c_str(), so.c_str());
    } else M->from_cache = true;
    M->dl = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!M->dl) throw std::string("dlopen of the specialised unit failed: ") + dlerror();
    M->track = (decltype(M->track))dlsym(M->dl, "hc_jit_sim_track");
    M->hook = (decltype(M->hook))dlsym(M->dl, "hc_jit_sim_hook");
    if (!M->track || !M->hook) throw std::string("specialised unit lacks its entry points");
#endif
    M->compile_ms = wall_ms() - t0;
    if (getenv("HC_B200_VERBOSE"))
        fprintf(stderr, "[hc_b200] specialised kernel %s: n = %d, lane state %zu + %zu B, %s in %.0f ms\n", hex, in.n, ss.hot, ss.cold,
                M->from_cache ? "loaded from cache" : "compiled", M->compile_ms);
    cache[h] = M;
    return M;
}

}  // namespace jit
}  // namespace hc
