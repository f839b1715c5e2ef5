// libhc_b200: kernels + C ABI (include/hc_b200.h).
//
// One persistent kernel per batch: grid = #SMs x resident CTAs, one solution path per group of
// G lanes (state in shared memory), all groups pull path indices from a device-side atomic queue (the `next_k` work counter of
// threaded_solve, reference src/solve.jl:641, 660-667, moved onto the device) and write their
// PathResult by path index (src/solve.jl:637, 670).
//
// With -DHC_HOST_SIM the same host logic and the same device headers are compiled by g++ into
// tests/host_sim/libhc_sim.so, which runs the lane state machine sequentially on the CPU.  That
// library exists only so that kernel logic can be unit-tested in a container without a GPU; it is
// never linked into, loaded by, or reachable from libhc_b200.so.
#include "../../include/hc_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "hc_kernel.h"
#include "hc_lower.h"
#include "hc_jit.h"

#ifndef HC_HOST_SIM
#include <cuda_runtime.h>
#endif

// The tracker kernels live in hc_kern.cu (one translation unit per instantiation, so that they compile in parallel);
// this file only needs their entry points.
namespace hc {
const void* hc_kernel_tpl(int slab);    // hc_track_tpl_kernel<slab>, slab = 12288 | 24576 | 49152
const void* hc_kernel_group(int G);     // hc_track_kernel<G>, G = 8 | 32
const void* hc_kernel_group_sync(int G);  // hc_track_kernel_sync<G>: its lockstep variant (second pass)
}

using namespace hc;

namespace {

thread_local std::string g_err;
thread_local hc_timing g_timing;
std::mutex g_mutex;

int fail(const std::string& msg) { g_err = msg; return -1; }

// ------------------------------------------------------------------ backend
#ifndef HC_HOST_SIM
#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) throw std::string(#call) + ": " + cudaGetErrorString(e_);            \
    } while (0)
// The devices this process drives (hc_init: one; hc_init_devices: a list).  Every device has its own non-blocking
// stream; handles keep one copy of their programs per device, a batch call splits its path index range over the
// devices (reference: one solve() drives all workers and stores results by path index, src/solve.jl:628-709).
// Device memory comes from the stream-ordered pool of the device (cudaMallocAsync on the device's stream) with the
// release threshold set to "never", so the ~30 buffers of a batch are recycled by the next call instead of being
// unmapped and mapped again (cudaFree of a 100 MB buffer synchronises the device and costs milliseconds).
// HC_B200_POOL=0 falls back to cudaMalloc / cudaFree.
struct Slot { int device = 0; cudaStream_t stream = nullptr; };
std::vector<Slot> g_slots;
int g_cur = 0;  // slot that dev_alloc / h2d / d2h / launches address (set by use_slot, under g_mutex)
bool g_pool = true;
int* g_cancel = nullptr;  // host-mapped flag polled by the kernels before they pull new paths (hc_request_cancel)
// HC_B200_NO_DEVICE=1: handles are built in host memory so that hc_jit_prepare can generate and compile the specialised
// kernel of a system on a machine without a GPU (NVRTC targets sm_100a offline; the driver's build check and the
// "not gpu" tests use this).  Every compute entry point fails in this mode.
bool nodev() { static const bool v = getenv("HC_B200_NO_DEVICE") && atoi(getenv("HC_B200_NO_DEVICE")) != 0; return v; }
int n_slots() { return g_slots.empty() ? 1 : (int)g_slots.size(); }
cudaStream_t cur_stream() { return g_slots.empty() ? (cudaStream_t)0 : g_slots[(size_t)g_cur].stream; }
// the CUDA current device is per host thread: every entry point selects its device again (a Julia task may run on
// any thread), under g_mutex
void use_slot(int s) {
    g_cur = s;
    if (nodev() || g_slots.empty()) return;
    CK(cudaSetDevice(g_slots[(size_t)s].device));
}
void dev_sync() { if (!nodev()) CK(cudaStreamSynchronize(cur_stream())); }
void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    if (nodev()) return calloc(bytes ? bytes : 16, 1);
    if (g_pool) CK(cudaMallocAsync(&p, bytes ? bytes : 16, cur_stream())); else CK(cudaMalloc(&p, bytes ? bytes : 16));
    return p;
}
void dev_free(void* p) { if (p) { if (nodev()) free(p); else if (g_pool) cudaFreeAsync(p, cur_stream()); else cudaFree(p); } }
// copies are stream-ordered: from pageable host memory the call returns once the source is staged, from page-locked
// memory (hc_host_register) it is a DMA transfer that the next dev_sync waits for -- caller-owned arrays outlive the call
void h2d(void* d, const void* h, size_t bytes) { if (bytes) { if (nodev()) memcpy(d, h, bytes); else CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, cur_stream())); } }
void d2h(void* h, const void* d, size_t bytes) { if (bytes) { if (nodev()) memcpy(h, d, bytes); else CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, cur_stream())); } }
void dev_zero(void* d, size_t bytes) { if (bytes) { if (nodev()) memset(d, 0, bytes); else CK(cudaMemsetAsync(d, 0, bytes, cur_stream())); } }
#else
int g_cur = 0;
int* g_cancel = nullptr;
int n_slots() { return 1; }
void use_slot(int s) { g_cur = s; }
void dev_sync() {}
void* dev_alloc(size_t bytes) { return calloc(bytes ? bytes : 16, 1); }
void dev_free(void* p) { free(p); }
void h2d(void* d, const void* h, size_t bytes) { if (bytes) memcpy(d, h, bytes); }
void d2h(void* h, const void* d, size_t bytes) { if (bytes) memcpy(h, d, bytes); }
void dev_zero(void* d, size_t bytes) { if (bytes) memset(d, 0, bytes); }
#endif

template <class T>
T* to_dev(const std::vector<T>& v) {
    T* d = (T*)dev_alloc(v.size() * sizeof(T));
    h2d(d, v.data(), v.size() * sizeof(T));
    return d;
}

// ------------------------------------------------------------------ handles
// device allocations of a handle: (slot, pointer)
struct Owned {
    std::vector<std::pair<int, void*>> v;
    void add(void* p) { v.push_back({g_cur, p}); }
    ~Owned() { const int keep = g_cur; for (auto& e : v) { try { use_slot(e.first); } catch (...) {} dev_free(e.second); } g_cur = keep; }
};
struct ProgramH {
    DevProgram dev;                // the copy on slot 0 (its sizes are the same on every device)
    std::vector<DevProgram> devs;  // pointers into the memory of each device of the process
    LoweredProgram low;
    bool tpp = false;   // segment-scheduled (thread-per-path engines) rather than level-scheduled (lane groups)
    jit::ProgCopy ref;  // the reference tape as handed over (input of the code generator, hc_jitgen.h)
    Owned owned;
};
struct SystemH {
    ProgramH eval, jac;
    // level-scheduled copies for the lane-group engine of a system whose programs were segment-scheduled for the
    // thread-per-path engines (n <= 14): built on first use by the second pass of a two-pass batch (group_programs)
    std::unique_ptr<ProgramH> eval_g, jac_g;
    int m = 0, n = 0, P = 0;
};
struct HomotopyH {
    DevHomotopy dev;                // slot 0
    std::vector<DevHomotopy> devs;  // per device
    SystemH* F = nullptr; SystemH* G = nullptr;
    Owned owned;
    std::vector<double> tw_hook;  // weights set through hc_toric_set_weights (test hook)
    std::shared_ptr<jit::Module> jit_mod[4];  // specialised kernels: [polyhedral driver ? 1 : 0] + 2 * [per-path parameter rows]
};

int env_int(const char* name, int def) { const char* v = getenv(name); return v ? atoi(v) : def; }

// lanes per path: enough lanes for the vectors / tape rounds of the system, never more than a warp
int group_size_for(int n) {
    int G = n > 8 ? 32 : 8;
    G = env_int("HC_B200_GROUP", G);
    if (G != 8 && G != 32) throw std::string("HC_B200_GROUP must be 8 or 32");
    return G;
}

int engine_for(int n);

// Lower the reference's 24-byte, 1-based Instruction stream (hc_lower.h) and upload it.
void build_program(ProgramH& H, const hc_program_desc* d, bool is_jac, bool force_group = false) {
    H.ref.assign(d);
#ifdef HC_HOST_SIM
    const bool tpp = true;
    (void)force_group;
#else
    const bool tpp = !force_group && engine_for(d->n_vars) != 0;
#endif
    H.tpp = tpp;
    // thread per path: segment scheduling (window of tape-order ops) feeds the segment loops of hc_tape.h;
    // lane groups: rounds of at most `cap` independent ops
    const int cap = tpp ? 0 : env_int("HC_B200_ROUND_CAP", 2 * group_size_for(d->n_vars));
    // Segment-scheduling window: longer segments (fewer dispatches, more independent ops per trip) cost tape
    // slots.  Take the largest window whose tape stays within 10 % of the window-32 tape: cyclic-7's Jacobian
    // goes from 65 to 28 segments for 152 -> 158 slots (+3-6 % paths/s), katsura(8) stays at 32 (a full window
    // would double its tape).  HC_B200_SEG_WINDOW / _SEG_WINDOW_EVAL pin the window instead.
    const bool prio = env_int("HC_B200_PRIO_HEIGHT", 1) != 0;
    const char* pin = getenv(is_jac ? "HC_B200_SEG_WINDOW" : "HC_B200_SEG_WINDOW_EVAL");
    if (!tpp) H.low = lower_program(d, cap, prio, false, 0);
    else if (pin) H.low = lower_program(d, 0, prio, false, atoi(pin));
    else {
        H.low = lower_program(d, 0, prio, false, 32);
        const int limit = H.low.W + H.low.W / 10;
        for (int w : {1000, 128, 64}) {
            LoweredProgram cand = lower_program(d, 0, prio, false, w);
            if (cand.W <= limit) { H.low = std::move(cand); break; }
        }
    }
    const LoweredProgram& L = H.low;
    H.devs.assign((size_t)n_slots(), DevProgram());
    for (int sl = 0; sl < n_slots(); ++sl) {
        use_slot(sl);
        DevProgram& P = H.devs[(size_t)sl];
        P.ops = to_dev(L.ops); P.level_end = to_dev(L.level_end); P.consts = to_dev(L.consts);
        P.u_assign = to_dev(L.u_assign); P.U_assign = to_dev(L.U_assign);
        P.fops = to_dev(L.fops); P.segs = to_dev(L.segs); P.n_segs = (int)L.segs.size(); P.n_fops = (int)L.fops.size();
        for (void* q : {(void*)P.ops, (void*)P.level_end, (void*)P.consts, (void*)P.u_assign, (void*)P.U_assign, (void*)P.fops, (void*)P.segs}) H.owned.add(q);
        P.n_levels = (int)L.level_end.size(); P.n_ops = (int)L.ops.size();
        P.C = (int)L.consts.size(); P.param_off = L.param_off; P.P = L.P; P.t_slot = L.t_slot;
        P.var_off = L.var_off; P.n = L.n; P.out_dim = L.out_dim; P.W = L.W;
        P.nu = (int)L.u_assign.size(); P.nU = (int)L.U_assign.size();
        dev_sync();  // the staging vectors of to_dev are temporaries
    }
    use_slot(0);
    H.dev = H.devs[0];
    const DevProgram& P = H.dev;
    if (getenv("HC_B200_VERBOSE"))
        fprintf(stderr, "[hc_b200] program: %d reference instructions -> %d micro-ops in %d levels (max width %d), %d segments, tape %d -> %d slots\n",
                d->n_instructions, P.n_ops, P.n_levels, L.max_width, P.n_segs, d->tape_space, P.W);
}

cx* cvec_dev(const double* p, int n, Owned& owned) {
    std::vector<cx> v(n);
    for (int i = 0; i < n; ++i) v[i] = mk(p[2 * i], p[2 * i + 1]);
    cx* d = to_dev(v);
    owned.add(d);
    dev_sync();
    return d;
}

DevOptions to_dev_options(const hc_options* o) {
    DevOptions D;
    static_assert(sizeof(DevOptions) == sizeof(hc_options), "hc_options / DevOptions layout mismatch");
    memcpy(&D, o, sizeof(D));
    return D;
}

// ------------------------------------------------------------------ kernels (KArgs, staging, tpp_loop: hc_kernel.h)
#ifndef HC_HOST_SIM
// single-path operator-API hooks (one thread: valid for levelised and for sequential programs)
__global__ void hc_hook_kernel(const KArgs A, int what, int K, const cx* x, const cx* xlo, cx t, const double* tw, cx* u, cx* U) {
    __shared__ KArgs sA;
    sA = A;
    Lane<1, 0> L;
    L.g.init();
    L.H = &sA.H; L.O = &sA.O; L.n = sA.H.n; L.pidx = L.prow = 0; L.kind = sA.H.kind;
    const size_t b = blockIdx.x;  // hc_evaluate_batch: one point per block (one thread each), its own slab and scratch
    carve(L.M, sA.H.n, sA.H.P, sA.H.tape_cx, hc_smem, A.cold + b * (size_t)A.cold_bytes);
    L.n_evaljac = L.n_eval = L.n_evaldd = L.n_tay1 = L.n_tay2 = L.n_tay3 = 0; L.tape_prog = L.tay_prog = nullptr; L.a_in_lu = L.rs_raw = false;
    const int n = sA.H.n;
    x += b * (size_t)(what == 3 ? K * n : n); u += b * (size_t)n;
    if (xlo) xlo += b * (size_t)n;
    if (U) U += b * (size_t)n * n;
    if (tw) for (int i = 0; i < sA.H.P; ++i) L.M.tw[i] = tw[i];
    if (what == 3) { for (int i = 0; i < K * n; ++i) L.M.tx[i] = x[i]; }
    else for (int i = 0; i < n; ++i) L.M.x[i] = x[i];
    if (what == 1) for (int i = 0; i < n; ++i) L.M.xhat[i] = xlo[i];
    if (what == 0) L.eval_f64(L.M.u, nullptr, L.M.x, t);
    else if (what == 1) L.eval_dd(L.M.u, L.M.x, &L.M.xhat, t);
    else if (what == 2) L.eval_f64(L.M.u, &L.M.A, L.M.x, t);
    else {
        if (K == 1) L.template taylor<1>(L.M.u, L.M.tx, t);
        else if (K == 2) L.template taylor<2>(L.M.u, L.M.tx, t);
        else if (K == 3) L.template taylor<3>(L.M.u, L.M.tx, t);
        else L.template taylor<4>(L.M.u, L.M.tx, t);
    }
    for (int i = 0; i < n; ++i) u[i] = L.M.u[i];
    if (what == 2) for (int i = 0; i < n * n; ++i) U[i] = L.M.A[i];
}

// Duplicate filter of the monodromy driver (reference: UniquePoints lookup in add_tracked_result!, src/monodromy.jl:1176-1200,
// src/unique_points.jl:247-285): candidate i -> index of the first known point within its radius, or -1.  One thread per
// candidate; the known points are read by all threads of a warp at the same address (broadcast).
__global__ void hc_unique_filter_kernel(int n, long long M, const cx* known, long long N, const cx* cand, double atol, double rtol, long long* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const cx* c = cand + i * n;
    double nrm2 = 0;
    for (int j = 0; j < n; ++j) nrm2 += abs2(c[j]);
    double rad = rtol * sqrt(nrm2);
    if (rad < atol) rad = atol;
    const double rad2 = rad * rad;
    long long hit = -1;
    for (long long k = 0; k < M && hit < 0; ++k) {
        double d = 0;
        for (int j = 0; j < n; ++j) d += abs2(known[k * n + j] - c[j]);
        if (d <= rad2) hit = k;
    }
    out[i] = hit;
}

// Device-side post-processing of a many-parameter solve (reference: many_solve's `transform_result`,
// src/solve.jl:422-430, 815-881, with a reduction such as nreal / nsolutions): per parameter point the counts
// [nonsingular, singular, real, at infinity, failed] over its S paths (is_real: max |imag| < tol, src/path_result.jl:280-298;
// no multiplicity clustering) -- 20 bytes per point leave the device instead of the PathResults of its paths.
__global__ void hc_sweep_counts_kernel(DevResults R, int n, long long S, long long M, double tol, int* counts) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    int ns = 0, sg = 0, re = 0, inf = 0, fl = 0;
    for (long long s = 0; s < S; ++s) {
        const long long k = j * S + s;
        const int rc = R.return_code[k];
        if (rc == EG_success) {
            if (R.singular[k]) ++sg; else ++ns;
            double m = 0.0;
            for (int i = 0; i < n; ++i) m = fmax(m, fabs(R.solution[k * n + i].im));
            if (m < tol) ++re;
        } else if (rc == EG_at_infinity || rc == EG_at_zero) ++inf;
        else ++fl;
    }
    int* c = counts + 5 * j;
    c[0] = ns; c[1] = sg; c[2] = re; c[3] = inf; c[4] = fl;
}

__global__ void hc_dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
#endif  // !HC_HOST_SIM

// ------------------------------------------------------------------ launch planning
struct Plan {
    int engine;  // 0 = lane group per path (state in shared memory), 1 = thread per path (state in local memory),
                 // 2 = thread per path, kernel specialised for the system at run time (hc_jit.h)
    int grid, block, group, paths_per_block;
    size_t smem, slab, cold, stage_bytes;
    int stage;
    long long lanes;
    std::shared_ptr<jit::Module> jit;
    int jit_tape_cx;
};

const size_t kSmemMax = 227 * 1024 - 2048;  // dynamic shared memory per CTA (the kernels keep ~1 KB static)

// ------------------------------------------------------------------ specialised kernels: policy + module
// HC_B200_JIT = 0: never; 1: always; auto (default): batches of at least HC_B200_JIT_MIN_PATHS paths -- compiling a
// system costs seconds, which only a large batch repays (the reference's `compile = :mixed` makes the same trade).
bool jit_wanted(const HomotopyH& H, long long N, bool complex_t = false) {
    if (complex_t && H.dev.kind != H_STRAIGHT_LINE) return false;  // the generated parameter code assumes real t
    const char* e = getenv("HC_B200_JIT");
    if (e && !strcmp(e, "0")) return false;
    if (H.dev.n > env_int("HC_B200_JIT_MAX_N", 20)) return false;  // the register-blocked LU of the specialised kernels goes up to n = 20
    if (e && !strcmp(e, "1")) return true;
#ifdef HC_HOST_SIM
    return false;
#else
    if (getenv("HC_B200_ENGINE")) return false;  // an engine was pinned explicitly
    return N >= env_int("HC_B200_JIT_MIN_PATHS", 8192);
#endif
}
size_t jit_stage_bytes(const HomotopyH& H) {  // what hc_jit_track stages: DoubleDouble interpreter tables + parameters
    auto r16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    auto sb = [&](const ProgramH& P) {
        return r16((size_t)P.dev.n_levels * sizeof(int)) + r16((size_t)P.dev.C * sizeof(cx)) + r16((size_t)P.dev.nu * sizeof(int2));
    };
    return sb(H.F->eval) + (H.dev.kind == H_STRAIGHT_LINE ? sb(H.G->eval) : 0) + 2 * 16 * (size_t)std::max(1, std::max(H.dev.P, H.G ? H.G->P : 0));
}
int jit_tape_cx(const HomotopyH& H) {  // tape of the DoubleDouble interpreter (2 cx per slot)
    int We = H.F->eval.dev.W;
    if (H.G && H.G->eval.dev.W > We) We = H.G->eval.dev.W;
    return std::max(2 * We, 2 * H.dev.P);  // ... which also holds the factors of the parameter series during a predictor update (jit_fill_pfac)
}
std::shared_ptr<jit::Module> jit_module(HomotopyH& H, bool poly, bool load = true, bool path_params = false) {
    const int slot = (poly ? 1 : 0) + (path_params ? 2 : 0);
    if (H.jit_mod[slot] && load) return H.jit_mod[slot];
    jit::GenInput in;
    in.kind = H.dev.kind; in.poly = poly; in.n = H.dev.n; in.path_params = path_params;
    in.hoist = env_int("HC_B200_JIT_HOIST", 1000);
    in.Fe = &H.F->eval.ref; in.Fj = &H.F->jac.ref;
    if (H.G) { in.Ge = &H.G->eval.ref; in.Gj = &H.G->jac.ref; }
    // Lanes per SM and where the LU factors live.  The factors are the most re-read piece of lane state; if
    // block x n^2 x 16 B fit next to the staged programs they go to shared memory (thread-interleaved), and the block
    // shrinks (in warps) until they do, down to HC_B200_JIT_MIN_BLOCK; otherwise they stay in local memory.
    int block = env_int("HC_B200_JIT_BLOCK", 256);
    int lu_smem = env_int("HC_B200_JIT_LU_SMEM", 0);  // measured: L1-cached local memory beats it (the rest of the lane state needs the L1)
#ifdef HC_HOST_SIM
    lu_smem = 0;
#else
    if (lu_smem) {
        const size_t per_lane = (size_t)H.dev.n * H.dev.n * 16, avail = kSmemMax - jit_stage_bytes(H);
        int b = block;
        while (b > 32 && (size_t)b * per_lane > avail) b -= 32;
        if ((size_t)b * per_lane <= avail && b >= env_int("HC_B200_JIT_MIN_BLOCK", 128)) block = b; else lu_smem = 0;
    }
#endif
    std::shared_ptr<jit::Module> M = jit::build_module(in, H.dev.P, jit_tape_cx(H), block, load, env_int("HC_B200_JIT_SYNC", 1), lu_smem);
    if (load) H.jit_mod[slot] = M;
    return M;
}

size_t program_stage_bytes(const ProgramH& P, bool fast) {
    auto r16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    return (fast ? r16(P.low.fops.size() * sizeof(FOp)) + r16(P.low.segs.size() * sizeof(int2)) : r16((size_t)P.dev.n_ops * sizeof(MOp))) + r16((size_t)P.dev.n_levels * sizeof(int)) + r16((size_t)P.dev.C * sizeof(cx)) +
           r16((size_t)P.dev.nu * sizeof(int2)) + r16((size_t)P.dev.nU * sizeof(int2));
}

const size_t kLocalSlabMax = 48 * 1024;      // largest instantiated local-memory slab (hc_track_tpl_kernel)

// Which engine tracks a system of n variables: one thread per path keeps every lane busy but its
// state (O(n^2) per path) lives in L2/HBM; a lane group per path keeps the state in shared memory
// but only pays off once the vectors are long enough to fill the group.
int engine_for(int n) {
    const char* e = getenv("HC_B200_ENGINE");
    if (e) {
        if (!strcmp(e, "tpp") || !strcmp(e, "local")) return 1;
        if (!strcmp(e, "group")) return 0;
        throw std::string("HC_B200_ENGINE must be tpp or group");
    }
    return n <= env_int("HC_B200_TPP_MAX_N", 14) ? 1 : 0;
}

Plan make_plan(HomotopyH& H, long long N, int mode, bool complex_t = false, bool path_params = false) {
    Plan p;
    p.engine = 0; p.grid = p.block = p.group = p.paths_per_block = 0; p.smem = p.slab = p.cold = p.stage_bytes = 0; p.stage = 0; p.lanes = 0;
    p.jit_tape_cx = 0;
    PathMem<0> dummy;
    if (jit_wanted(H, N, complex_t)) {
        p.jit = jit_module(H, mode == MODE_POLYHEDRAL, true, path_params);
        p.engine = 2; p.group = 1; p.stage = 1;
        p.jit_tape_cx = jit_tape_cx(H);
        p.slab = p.jit->hot; p.cold = p.jit->cold;
        p.stage_bytes = jit_stage_bytes(H);
        p.block = p.jit->block;
        int sms = 148;
#ifndef HC_HOST_SIM
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#endif
        const int per_sm = env_int("HC_B200_BLOCKS_PER_SM", 1);
        long long want_lanes = N;
        const long long lanes_cap = (long long)sms * per_sm * p.block;
        if (want_lanes > lanes_cap) want_lanes = lanes_cap;
        long long want = (want_lanes + p.block - 1) / p.block;
        p.grid = (int)(want < 1 ? 1 : want);
        p.paths_per_block = p.block;
        p.smem = p.stage_bytes + (p.jit->lu_smem ? (size_t)p.block * H.dev.n * H.dev.n * 16 : 0);
        p.lanes = (long long)p.grid * p.block;
        return p;
    }
    SlabSizes ss = carve(dummy, H.dev.n, H.dev.P, H.dev.tape_cx, nullptr, nullptr);
    p.slab = ss.hot; p.cold = ss.cold;
#ifndef HC_HOST_SIM
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    p.engine = engine_for(H.dev.n);
    auto stage_plan = [&](bool fast) {
        p.stage_bytes = program_stage_bytes(H.F->eval, fast) + program_stage_bytes(H.F->jac, fast);
        if (H.dev.kind == H_STRAIGHT_LINE) p.stage_bytes += program_stage_bytes(H.G->eval, fast) + program_stage_bytes(H.G->jac, fast);
        p.stage_bytes += 2 * 16 * (size_t)std::max(1, std::max(H.dev.P, H.G ? H.G->P : 0));  // stage_params
        p.stage = (p.stage_bytes <= (size_t)env_int("HC_B200_STAGE_MAX", 96 * 1024)) && env_int("HC_B200_STAGE", 1);
        if (!p.stage) p.stage_bytes = 0;
    };
    stage_plan(p.engine != 0);
    // the thread-per-path kernel reads its programs with ld.shared and its state with ld.local: both must fit
    if (p.engine == 1 && (p.slab + p.cold > kLocalSlabMax || !p.stage)) {
        if (getenv("HC_B200_ENGINE")) throw std::string("thread-per-path engine: state or programs of this system do not fit (use the group engine)");
        p.engine = 0;  // (the segment-scheduled programs are valid, if narrow, level programs for a lane group)
        stage_plan(false);
    }
    if (p.engine != 0) {
        p.group = 1;
        // One CTA per SM: 4 warps (8 for tiny systems).  The lane state is cached in L1, so a single copy of
        // the staged programs (smallest shared-memory carve-out) and few lanes beat more warps; 255
        // registers per thread keep the unrolled segment loops and the LU columns out of local memory
        // (profiles/r01_sweep.md).
        p.block = env_int("HC_B200_BLOCK", H.dev.n <= 4 ? 256 : 128);
        if (p.block % 32 || p.block < 32 || p.block > 256) throw std::string("HC_B200_BLOCK must be a multiple of 32 in [32, 256]");
        int per_sm = env_int("HC_B200_BLOCKS_PER_SM", 1);
        // lanes: at most a fraction of the paths, so that finished lanes have work to refill with
        long long lanes_cap = (long long)sms * per_sm * p.block;
        long long want_lanes = (N + env_int("HC_B200_PATHS_PER_LANE", 1) - 1) / env_int("HC_B200_PATHS_PER_LANE", 1);
        if (want_lanes > lanes_cap) want_lanes = lanes_cap;
        long long want = (want_lanes + p.block - 1) / p.block;
        if (want < sms && p.block > 32) { p.block = 32; want = (want_lanes + 31) / 32; }  // few paths: spread over the SMs
        p.grid = (int)(want < 1 ? 1 : want);
        p.paths_per_block = p.block;
        p.smem = p.stage_bytes;
        p.lanes = (long long)p.grid * p.block;
        return p;
    }
    const int G = group_size_for(H.dev.n);
    p.group = G;
    if (p.stage_bytes + p.slab > kSmemMax) {
        p.stage = 0; p.stage_bytes = 0;
        if (p.slab > kSmemMax) throw std::string("system too large: one path needs ") + std::to_string(p.slab) + " bytes of shared memory";
    }
    // block: up to 256 threads; shrink until the slabs of its paths fit next to the staged programs
    int block = env_int("HC_B200_BLOCK", 256);
    if (block % 32 || block < 32 || block > 256) throw std::string("HC_B200_BLOCK must be a multiple of 32 in [32, 256]");
    long long small = (N * G + sms - 1) / sms;            // few paths: spread them over the SMs
    while (block > 32 && block / 2 >= small) block /= 2;
    if (block < G) block = G;
    while (block > G && p.stage_bytes + (size_t)(block / G) * p.slab > kSmemMax) block -= 32;
    p.block = block; p.paths_per_block = block / G;
    p.smem = p.stage_bytes + (size_t)p.paths_per_block * p.slab;
    int per_sm = (int)((228 * 1024) / (p.smem + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * block > 512) per_sm = 512 / block;       // register file: 128 registers per thread
    per_sm = env_int("HC_B200_BLOCKS_PER_SM", per_sm);
    long long want = (N + p.paths_per_block - 1) / p.paths_per_block;
    long long cap = (long long)sms * per_sm;
    p.grid = (int)(want < cap ? want : cap);
    if (p.grid < 1) p.grid = 1;
    p.lanes = (long long)p.grid * p.paths_per_block;
#else
    (void)N;
    p.block = 1; p.grid = 1; p.smem = p.slab; p.stage = 0; p.stage_bytes = 0; p.group = 1; p.paths_per_block = 1; p.lanes = 1;
#endif
    return p;
}

struct DeviceBatch {  // device-resident inputs and outputs of one batch
    HomotopyH* H = nullptr;
    int mode = 0; long long N = 0; int n = 0;
    KArgs A;
    Plan plan;
    int slot = 0;          // device of the batch
    long long first = 0;   // index of its first path in the caller's arrays
    std::vector<void*> owned;
    int64_t h2d_bytes = 0;
    // two-pass batches: paths the thread-per-path pass handed over are tracked again by the lane-group engine
    long long handoff_paths = 0; double handoff_ms = 0;
    unsigned char* cold2 = nullptr; size_t cold2_bytes = 0; long long* map2 = nullptr; size_t map2_count = 0;
#ifndef HC_HOST_SIM
    cudaEvent_t e0 = nullptr, e1 = nullptr, f0 = nullptr, f1 = nullptr;
    bool second_running = false;
#endif
    ~DeviceBatch() {
        const int keep = g_cur;
        try { use_slot(slot); } catch (...) {}
        for (void* p : owned) dev_free(p);
#ifndef HC_HOST_SIM
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (f0) cudaEventDestroy(f0);
        if (f1) cudaEventDestroy(f1);
#endif
        g_cur = keep;
    }
    template <class T> T* alloc(size_t count) { T* p = (T*)dev_alloc(count * sizeof(T)); owned.push_back(p); return p; }
    template <class T> T* upload(const T* h, size_t count) {
        T* p = alloc<T>(count); h2d(p, h, count * sizeof(T)); h2d_bytes += (int64_t)(count * sizeof(T)); return p;
    }
};

// How the start solutions (and the rows of the per-path parameters) of a batch are produced.
struct SweepCounts { int32_t* counts = nullptr; long long S = 0; double real_tol = 1e-6; };  // hc_track_sweep_counts

struct StartGen {
    long long start_rows = 0;           // > 0: `starts` holds this many rows, path k starts from row k % start_rows
    long long param_div = 0;            // > 0: the per-path parameter arrays hold one row per param_div consecutive paths
    const int32_t* degrees = nullptr;   // total degree: no `starts` at all
    long long td_first = 0;             // index of the batch's first path in the start system (total degree and binomial starts)
    // polyhedral start solutions made on the device: per mixed cell its volume, the Hermite normal form of its binomial
    // system, the transformed angles and the moduli (hc_polyhedral_track_cells)
    const int64_t* bin_volume = nullptr;
    const int64_t* bin_H = nullptr;
    const double* bin_mu = nullptr;
    const double* bin_r = nullptr;
};

void setup_batch(DeviceBatch& D, HomotopyH* H, const hc_options* o, int mode, long long N, const double* starts, const double* t1,
                 const double* t0, const double* path_p, const double* path_q, const double* omega_mu, const int32_t* cell_index,
                 const double* cell_weights, int ncells, const StartGen& sg = StartGen(), int slot = 0) {
    D.H = H; D.mode = mode; D.N = N; D.n = H->dev.n; D.slot = slot;
    use_slot(slot);
    const int n = D.n, P = H->dev.P;
    memset(&D.A, 0, sizeof(D.A));
    D.A.H = H->devs[(size_t)slot];
    D.A.H.N = N;
    D.A.O = to_dev_options(o);
    BatchIn& B = D.A.B;
    B.mode = mode; B.N = N;
    if (sg.degrees) {
        // roots of unity per variable, exactly the values the host iterator yields (cis(2 pi j / d), total_degree.jl:258)
        std::vector<double> roots; std::vector<int> deg(n);
        double total = 1.0;
        for (int i = 0; i < n; ++i) {
            const int d = sg.degrees[i];
            if (d < 1) throw std::string("total degree start system: degrees must be >= 1");
            deg[i] = d; total *= d;
            for (int j = 0; j < d; ++j) { const double a = 2.0 * M_PI * j / d; roots.push_back(cos(a)); roots.push_back(sin(a)); }
        }
        if (sg.td_first < 0 || (double)sg.td_first + (double)N > total) throw std::string("total degree start system: path range exceeds prod(degrees)");
        B.td_roots = (const cx*)D.upload<double>(roots.data(), roots.size());
        B.td_degrees = D.upload<int>(deg.data(), (size_t)n);
        B.td_first = sg.td_first;
    } else if (sg.bin_H) {
        if (mode != MODE_POLYHEDRAL || ncells <= 0 || !sg.bin_volume || !sg.bin_mu || !sg.bin_r) throw std::string("binomial start solutions: cell data missing");
        std::vector<long long> first((size_t)ncells + 1, 0);
        for (int c = 0; c < ncells; ++c) {
            long long det = 1;
            for (int i = 0; i < n; ++i) {
                const long long d = sg.bin_H[((size_t)c * n + i) * n + i];
                if (d < 1) throw std::string("binomial start solutions: the diagonal of a Hermite normal form must be positive");
                det *= d;
            }
            if (det != sg.bin_volume[c]) throw std::string("binomial start solutions: volume of cell ") + std::to_string(c) + " is not det(H)";
            first[(size_t)c + 1] = first[(size_t)c] + det;
        }
        if (sg.td_first < 0) throw std::string("binomial start solutions: negative first path index");
        B.ncells = ncells;
        B.cell_first = D.upload<long long>(first.data(), first.size());
        B.bin_H = D.upload<long long>((const long long*)sg.bin_H, (size_t)ncells * n * n);
        B.bin_mu = D.upload<double>(sg.bin_mu, (size_t)ncells * n);
        B.bin_r = D.upload<double>(sg.bin_r, (size_t)ncells * n);
        B.k_first = sg.td_first;
    } else {
        if (!starts) throw std::string("start solutions missing");
        B.start_mod = sg.start_rows;
        B.starts = (const cx*)D.upload<double>(starts, (size_t)2 * n * (sg.start_rows > 0 ? sg.start_rows : N));
    }
    B.param_div = sg.param_div;
    const long long prow = sg.param_div > 0 ? (N + sg.param_div - 1) / sg.param_div : N;
    B.t1 = t1 ? mk(t1[0], t1[1]) : mk(1.0); B.t0 = t0 ? mk(t0[0], t0[1]) : mk(0.0);
    B.omega_mu = omega_mu ? D.upload<double>(omega_mu, (size_t)2 * N) : nullptr;
    // per-path parameters stay path-major (P values per path, read once per step and lane)
    D.A.H.path_p = path_p ? (const cx*)D.upload<double>(path_p, (size_t)2 * P * prow) : nullptr;
    D.A.H.path_q = path_q ? (const cx*)D.upload<double>(path_q, (size_t)2 * P * prow) : nullptr;
    if (mode == MODE_POLYHEDRAL) {
        if ((!cell_index && !sg.bin_H) || !cell_weights || ncells <= 0) throw std::string("polyhedral batch: cell_index / cell_weights missing");
        if (!sg.bin_H) {
            for (long long k = 0; k < N; ++k)
                if (cell_index[k] < 0 || cell_index[k] >= ncells) throw std::string("polyhedral batch: cell_index[") + std::to_string(k) + "] is outside [0, ncells)";
            B.cell_index = D.upload<int32_t>(cell_index, (size_t)N);
        }
        B.cell_weights = D.upload<double>(cell_weights, (size_t)ncells * P);
    }
    DevResults& R = D.A.R;
    R.return_code = D.alloc<int>(N); R.solution = D.alloc<cx>((size_t)n * N); R.t = D.alloc<double>(N);
    R.accuracy = D.alloc<double>(N); R.residual = D.alloc<double>(N); R.singular = D.alloc<unsigned char>(N);
    R.condition_jacobian = D.alloc<double>(N); R.winding_number = D.alloc<int>(N);
    R.extended_precision = D.alloc<unsigned char>(N); R.last_point = D.alloc<cx>((size_t)n * N); R.last_t = D.alloc<double>(N);
    R.valuation = D.alloc<double>((size_t)n * N); R.has_valuation = D.alloc<unsigned char>(N); R.omega = D.alloc<double>(N);
    R.mu = D.alloc<double>(N); R.accepted_steps = D.alloc<int>(N); R.rejected_steps = D.alloc<int>(N);
    R.steps_eg = D.alloc<int>(N); R.extended_precision_used = D.alloc<unsigned char>(N);
    R.counters = D.alloc<long long>((size_t)8 * N);
    dev_zero(R.return_code, (size_t)N * sizeof(int));  // a cancelled batch leaves the paths it never started at 0 (= tracking)
    D.A.cancel = g_cancel;
    D.plan = make_plan(*H, N, mode, B.t1.im != 0.0 || B.t0.im != 0.0, path_p != nullptr || path_q != nullptr);
    const Plan& pl = D.plan;
    D.A.queue = D.alloc<unsigned long long>(1);
    D.A.stage = pl.stage;
    D.A.slab_bytes = (int)pl.slab;
    D.A.cold_bytes = (int)pl.cold;
    if (pl.engine == 1) D.A.refill_min = env_int("HC_B200_REFILL_MIN", 8);
    else if (pl.engine == 0) D.A.cold = D.alloc<unsigned char>((size_t)pl.grid * pl.paths_per_block * pl.cold);
    D.A.stage_bytes = (int)pl.stage_bytes;
    if (pl.engine == 2) {
        D.A.H.tape_cx = pl.jit_tape_cx;
        D.A.refill_min = env_int("HC_B200_REFILL_MIN", 0);   // 0: adaptive (tpp_loop_sync), > 0: that many parked lanes per warp
        const char* rk = getenv("HC_B200_REFILL_K");
        D.A.refill_k = rk ? (float)atof(rk) : 1.0f;
    }
    D.A.sync_cta = env_int("HC_B200_SYNC_CTA", 0);
#ifndef HC_HOST_SIM
    // Two-pass batch: the specialised thread-per-path kernel gives up paths beyond HC_B200_HANDOFF_EG_STEPS endgame steps
    // (typical paths need 10 - 80; the 99th percentile of the heavy-tailed configs is 210 - 230), beyond
    // HC_B200_HANDOFF_STEPS steps in total, or in extended precision; finish_batch tracks them again on the lane-group
    // engine (second_pass).
    if (pl.engine == 2 && mode != MODE_TRACKER && env_int("HC_B200_HANDOFF", 1)) {
        B.handoff_steps = env_int("HC_B200_HANDOFF_STEPS", 1000);
        B.handoff_eg_steps = env_int("HC_B200_HANDOFF_EG_STEPS", 120);
        B.handoff_ext = env_int("HC_B200_HANDOFF_EXT", 1);
    }
#endif
}

#ifndef HC_HOST_SIM
// function attributes and the stack limit are per device: remember what was set where
bool first_use(const void* kern) {
    static std::vector<std::pair<const void*, int>> seen;
    for (auto& e : seen) if (e.first == kern && e.second == g_cur) return false;
    seen.push_back({kern, g_cur});
    return true;
}
void launch_tpl(const DeviceBatch& D, int SLAB) {
    const void* kern = hc_kernel_tpl(SLAB);
    if (first_use(kern)) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
        size_t cur = 0;
        CK(cudaDeviceGetLimit(&cur, cudaLimitStackSize));
        if (cur < (size_t)SLAB + 8192) CK(cudaDeviceSetLimit(cudaLimitStackSize, (size_t)SLAB + 8192));
    }
    // the lane state lives in L1 (local memory): ask for the smallest shared-memory carve-out that holds
    // the staged programs of the CTAs resident on one SM
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = (D.plan.grid + sms - 1) / sms;
        int pct = env_int("HC_B200_CARVEOUT", (int)((per_sm * (D.plan.smem + 3072) * 100 + 228 * 1024 - 1) / (228 * 1024)));
        if (pct > 100) pct = 100;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    void* args[] = {(void*)&D.A};
    CK(cudaLaunchKernel(kern, dim3((unsigned)D.plan.grid), dim3((unsigned)D.plan.block), args, D.plan.smem, cur_stream()));
}
void launch_track(const DeviceBatch& D, int G) {
    const void* kern = hc_kernel_group(G);
    if (first_use(kern)) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    void* args[] = {(void*)&D.A};
    CK(cudaLaunchKernel(kern, dim3((unsigned)D.plan.grid), dim3((unsigned)D.plan.block), args, D.plan.smem, cur_stream()));
}
#endif

#ifndef HC_HOST_SIM
void launch_jit(const DeviceBatch& D) {
    jit::Module& M = *D.plan.jit;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (first_use((const void*)M.track)) {
        CK(cudaFuncSetAttribute((const void*)M.track, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
        size_t cur = 0;
        CK(cudaDeviceGetLimit(&cur, cudaLimitStackSize));
        if (cur < M.slab + 8192) CK(cudaDeviceSetLimit(cudaLimitStackSize, M.slab + 8192));
    }
    const int per_sm = (D.plan.grid + sms - 1) / sms;
    int pct = env_int("HC_B200_CARVEOUT", (int)((per_sm * (D.plan.smem + 3072) * 100 + 228 * 1024 - 1) / (228 * 1024)));
    if (pct > 100) pct = 100;
    CK(cudaFuncSetAttribute((const void*)M.track, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    void* args[] = {(void*)&D.A};
    CK(cudaLaunchKernel((const void*)M.track, dim3((unsigned)D.plan.grid), dim3((unsigned)D.plan.block), args, D.plan.smem, cur_stream()));
}
#endif

// launches the batch on its device's stream (asynchronous); finish_batch waits for it and returns kernel milliseconds
void launch_batch(DeviceBatch& D) {
    use_slot(D.slot);
    dev_zero(D.A.queue, sizeof(unsigned long long));
#ifndef HC_HOST_SIM
    if (!D.e0) { CK(cudaEventCreate(&D.e0)); CK(cudaEventCreate(&D.e1)); }
    CK(cudaEventRecord(D.e0, cur_stream()));
    if (D.plan.engine == 2) launch_jit(D);
    else if (D.plan.engine == 1) {
        const size_t need = D.plan.slab + D.plan.cold;
        launch_tpl(D, need <= 12 * 1024 ? 12 * 1024 : (need <= 24 * 1024 ? 24 * 1024 : 48 * 1024));
    } else launch_track(D, D.plan.group == 32 ? 32 : 8);
    CK(cudaEventRecord(D.e1, cur_stream()));
    CK(cudaGetLastError());
#endif
}
#ifndef HC_HOST_SIM
hc_program_desc desc_of(const jit::ProgCopy& r) {
    hc_program_desc d;
    d.instructions = r.instr.data(); d.n_instructions = (int32_t)(r.instr.size() / 6);
    d.constants = r.consts.data(); d.n_constants = (int32_t)(r.consts.size() / 2);
    d.param_offset = r.param_offset; d.n_params = r.n_params; d.t_index = r.t_index; d.var_offset = r.var_offset; d.n_vars = r.n_vars;
    d.u_assign = r.u_assign.data(); d.n_u = (int32_t)(r.u_assign.size() / 2);
    d.U_assign = r.U_assign.data(); d.n_U = (int32_t)(r.U_assign.size() / 2);
    d.out_dim = r.out_dim; d.tape_space = r.tape_space;
    return d;
}
// the programs of a system as the lane-group engine wants them (level-scheduled)
void group_programs(SystemH& S, ProgramH*& e, ProgramH*& j) {
    if (!S.eval.tpp) { e = &S.eval; j = &S.jac; return; }
    if (!S.eval_g) {
        const int keep = g_cur;
        std::unique_ptr<ProgramH> pe(new ProgramH()), pj(new ProgramH());
        const jit::ProgCopy re = S.eval.ref, rj = S.jac.ref;  // (build_program re-assigns ref from the descriptor)
        hc_program_desc de = desc_of(re), dj = desc_of(rj);
        build_program(*pe, &de, false, true);
        build_program(*pj, &dj, true, true);
        S.eval_g = std::move(pe); S.jac_g = std::move(pj);
        use_slot(keep);
    }
    e = S.eval_g.get(); j = S.jac_g.get();
}

// Second pass of a two-pass batch: the paths the thread-per-path kernel gave up (return_code == RC_HANDOFF) are tracked
// again, from their start solutions, by the lane-group engine -- a group of lanes walks one long path several times
// faster than a single lane, and nothing waits for it.  Results land at the paths' own indices.  The launch is
// asynchronous (finish_second waits for it), so the second passes of several devices overlap.
void second_pass(DeviceBatch& D) {
    const long long N = D.N;
    std::vector<int> rc((size_t)N);
    d2h(rc.data(), D.A.R.return_code, (size_t)N * 4);
    dev_sync();
    // the paths handed over for their step count first: they are the ones that can run to max_endgame_steps, and the
    // kernel ends when the last of them ends -- so they must not start behind the (many, short) extended-precision paths
    std::vector<long long> idx;
    for (long long k = 0; k < N; ++k) if (rc[(size_t)k] == RC_HANDOFF_STEPS) idx.push_back(k);
    for (long long k = 0; k < N; ++k) if (rc[(size_t)k] == RC_HANDOFF) idx.push_back(k);
    D.handoff_paths = (long long)idx.size();
    if (idx.empty()) return;
    HomotopyH& H = *D.H;
    KArgs A2 = D.A;
    ProgramH *Fe, *Fj, *Ge = nullptr, *Gj = nullptr;
    group_programs(*H.F, Fe, Fj);
    if (H.dev.kind == H_STRAIGHT_LINE) group_programs(*H.G, Ge, Gj);
    use_slot(D.slot);
    A2.H.Fe = Fe->devs[(size_t)D.slot]; A2.H.Fj = Fj->devs[(size_t)D.slot];
    int W = Fj->dev.W, We = Fe->dev.W;
    if (Ge) { A2.H.Ge = Ge->devs[(size_t)D.slot]; A2.H.Gj = Gj->devs[(size_t)D.slot]; W = std::max(W, Gj->dev.W); We = std::max(We, Ge->dev.W); }
    A2.H.tape_cx = std::max(W, 4 * We);
    // launch shape: as make_plan's lane-group branch
    PathMem<0> dummy;
    const SlabSizes ss = carve(dummy, H.dev.n, H.dev.P, A2.H.tape_cx, nullptr, nullptr);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t stage_bytes = program_stage_bytes(*Fe, false) + program_stage_bytes(*Fj, false) + (Ge ? program_stage_bytes(*Ge, false) + program_stage_bytes(*Gj, false) : 0) +
                         2 * 16 * (size_t)std::max(1, std::max(H.dev.P, H.G ? H.G->P : 0));
    int stage = stage_bytes <= (size_t)env_int("HC_B200_STAGE_MAX", 96 * 1024) && env_int("HC_B200_STAGE", 1);
    if (!stage || stage_bytes + ss.hot > kSmemMax) { stage = 0; stage_bytes = 0; }
    if (ss.hot > kSmemMax) throw std::string("system too large for the lane-group engine");
    const long long N2 = (long long)idx.size();
    // A warp per path (32 lanes).  If the handed-over paths fit the chip at once the kernel is the latency of its longest
    // path and the warps run free (cyclooctane: 1.0 s; in lockstep 1.1 - 1.3 s); if they come in several waves the warps
    // of a CTA run in lockstep and share what they fetch (tritangents, 6 735 paths, 12 warps per SM walking different
    // code: 1.95 s; in lockstep 1.19 s; 8 lanes per path, unsynchronised: 1.35 s).  HC_B200_HANDOFF_SYNC / _GROUP pin either.
    const int sync_env = env_int("HC_B200_HANDOFF_SYNC", -1);
    const bool sync2 = sync_env >= 0 ? sync_env != 0 : N2 * 32 > (long long)sms * 512;
    int G = env_int("HC_B200_HANDOFF_GROUP", 0);
    if (G != 8 && G != 32) G = 32;
    int block = 256;
    const long long small = (N2 * G + sms - 1) / sms;
    while (block > 32 && block / 2 >= small) block /= 2;
    if (block < G) block = G;
    while (block > G && stage_bytes + (size_t)(block / G) * ss.hot > kSmemMax) block -= 32;
    const int ppb = block / G;
    const size_t smem = stage_bytes + (size_t)ppb * ss.hot;
    int per_sm = (int)((228 * 1024) / (smem + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * block > 512) per_sm = 512 / block;                  // register file: 128 registers per thread
    const long long want = (N2 + ppb - 1) / ppb, cap = (long long)sms * per_sm;
    const int grid = (int)std::max(1LL, std::min(want, cap));
    const size_t cold_need = (size_t)grid * ppb * ss.cold;
    if (cold_need > D.cold2_bytes) { D.cold2 = D.alloc<unsigned char>(cold_need); D.cold2_bytes = cold_need; }
    if (idx.size() > D.map2_count) { D.map2 = D.alloc<long long>(idx.size()); D.map2_count = idx.size(); }
    h2d(D.map2, idx.data(), idx.size() * sizeof(long long));
    A2.B.N = N2; A2.B.index_map = D.map2; A2.B.handoff_steps = 0; A2.B.handoff_eg_steps = 0; A2.B.handoff_ext = 0;
    A2.stage = stage; A2.stage_bytes = (int)stage_bytes; A2.slab_bytes = (int)ss.hot; A2.cold_bytes = (int)ss.cold; A2.cold = D.cold2;
    dev_zero(A2.queue, sizeof(unsigned long long));
    const void* kern = sync2 ? hc_kernel_group_sync(G) : hc_kernel_group(G);
    if (first_use(kern)) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    if (!D.f0) { CK(cudaEventCreate(&D.f0)); CK(cudaEventCreate(&D.f1)); }
    CK(cudaEventRecord(D.f0, cur_stream()));
    void* args[] = {(void*)&A2};
    CK(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3((unsigned)block), args, smem, cur_stream()));
    CK(cudaEventRecord(D.f1, cur_stream()));
    D.second_running = true;
    if (env_int("HC_B200_VERBOSE", 0) >= 1)
        fprintf(stderr, "[hc_b200] second pass: %lld of %lld paths on the lane-group engine (G = %d, grid %d x %d)\n", N2, N, G, grid, block);
}
#endif
// waits for the second pass of the batch, if one was launched; returns its milliseconds
double finish_second(DeviceBatch& D) {
#ifndef HC_HOST_SIM
    if (!D.second_running) return 0.0;
    use_slot(D.slot);
    D.second_running = false;
    CK(cudaEventSynchronize(D.f1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, D.f0, D.f1));
    D.handoff_ms = ms;
    return ms;
#else
    (void)D;
    return 0.0;
#endif
}

double finish_batch(DeviceBatch& D) {
    use_slot(D.slot);
#ifndef HC_HOST_SIM
    CK(cudaEventSynchronize(D.e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, D.e0, D.e1));
    D.handoff_paths = 0; D.handoff_ms = 0;
    if (D.A.B.handoff_steps > 0) second_pass(D);   // asynchronous: finish_second
    return ms;
#else
    if (D.plan.engine == 2) {
        std::vector<unsigned char> hot(D.plan.slab + 16), cold(D.plan.cold + 16);
        D.plan.jit->track(&D.A, (unsigned char*)(((uintptr_t)hot.data() + 15) & ~(uintptr_t)15), (unsigned char*)(((uintptr_t)cold.data() + 15) & ~(uintptr_t)15));
        return 0.0;
    }
    std::vector<unsigned char> slab(D.plan.slab + 16);
    Lane<1, 0> L;
    L.g.init();
    L.H = &D.A.H; L.O = &D.A.O; L.n = D.A.H.n;
    unsigned char* base = (unsigned char*)(((uintptr_t)slab.data() + 15) & ~(uintptr_t)15);
    carve(L.M, D.A.H.n, D.A.H.P, D.A.H.tape_cx, base, D.A.cold);
    sim_loop(L, D.A);
    return 0.0;
#endif
}
double run_batch(DeviceBatch& D) { launch_batch(D); const double ms = finish_batch(D); return ms + finish_second(D); }

// copies the PathResult arrays of the batch into the caller's arrays at the batch's path offset (stream-ordered:
// dev_sync() before the host reads them)
void fetch_results(DeviceBatch& D, hc_results* out) {
    use_slot(D.slot);
    const DevResults& R = D.A.R;
    const size_t N = (size_t)D.N, n = (size_t)D.n, f = (size_t)D.first;
    d2h(out->return_code + f, R.return_code, N * 4); d2h(out->solution + 2 * n * f, R.solution, n * N * 16); d2h(out->t + f, R.t, N * 8);
    d2h(out->accuracy + f, R.accuracy, N * 8); d2h(out->residual + f, R.residual, N * 8); d2h(out->singular + f, R.singular, N);
    d2h(out->condition_jacobian + f, R.condition_jacobian, N * 8); d2h(out->winding_number + f, R.winding_number, N * 4);
    d2h(out->extended_precision + f, R.extended_precision, N); d2h(out->last_point + 2 * n * f, R.last_point, n * N * 16);
    d2h(out->last_t + f, R.last_t, N * 8); d2h(out->valuation + n * f, R.valuation, n * N * 8); d2h(out->has_valuation + f, R.has_valuation, N);
    d2h(out->omega + f, R.omega, N * 8); d2h(out->mu + f, R.mu, N * 8); d2h(out->accepted_steps + f, R.accepted_steps, N * 4);
    d2h(out->rejected_steps + f, R.rejected_steps, N * 4); d2h(out->steps_eg + f, R.steps_eg, N * 4);
    d2h(out->extended_precision_used + f, R.extended_precision_used, N);
    if (out->counters) d2h(out->counters + 8 * f, R.counters, N * 64);
}
int64_t result_bytes(long long N, int n, bool counters) {
    return (int64_t)N * (4 + 16 * n + 8 + 8 + 8 + 1 + 8 + 4 + 1 + 16 * n + 8 + 8 * n + 1 + 8 + 8 + 4 + 4 + 4 + 1 + (counters ? 64 : 0));
}

double now_ms() {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// A polyhedral batch runs on a merged homotopy: the toric stage and the coefficient stage share F;
// p = start coefficients (toric system coefficients), q = target coefficients.
HomotopyH* merged_polyhedral(HomotopyH* toric, HomotopyH* coeff) {
    if (toric->F != coeff->F) throw std::string("toric and coefficient homotopy must share the system handle");
    if (toric->dev.kind != H_TORIC || coeff->dev.kind != H_COEFFICIENT) throw std::string("expected (toric, coefficient) homotopies");
    return coeff;  // coeff.p are the start coefficients == toric system coefficients (src/polyhedral.jl:397-406)
}

void ensure_init();

int track_impl(HomotopyH* H, const hc_options* o, int mode, long long N, const double* starts, const double* t1, const double* t0,
               const double* path_p, const double* path_q, const double* omega_mu, const int32_t* cell_index,
               const double* cell_weights, int ncells, hc_results* out, const StartGen& sg = StartGen(), const SweepCounts* sc = nullptr) {
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        if (N <= 0) return 0;
        if (!H || !o || (!out && !sc)) throw std::string("null handle / options / results");
#ifndef HC_HOST_SIM
        if (nodev()) throw std::string("HC_B200_NO_DEVICE is set: this process can only build kernels, not track (there is no CPU fallback)");
#endif
        ensure_init();
        if ((int)H->devs.size() != n_slots()) throw std::string("handle was created before hc_init_devices changed the device list");
        if (g_cancel) *g_cancel = 0;
        const double tA = now_ms();
        // Shards: contiguous path index ranges, one per device (sweeps: whole parameter points).  Small batches stay on
        // one device -- a device wants thousands of paths per launch.
        const int n = H->dev.n, P = H->dev.P;
        const long long unit = sg.start_rows > 0 ? sg.start_rows : 1;
        int nd = n_slots();
        const long long min_per_dev = env_int("HC_B200_MIN_PATHS_PER_DEVICE", 2048);
        while (nd > 1 && N / nd < min_per_dev) --nd;
        std::vector<std::unique_ptr<DeviceBatch>> Ds;
        long long lo = 0;
        for (int sl = 0; sl < nd; ++sl) {
            long long hi = sl == nd - 1 ? N : ((N / unit) * (sl + 1) / nd) * unit;
            if (hi <= lo) continue;
            const long long prow0 = sg.param_div > 0 ? lo / sg.param_div : lo;
            StartGen g2 = sg;
            g2.td_first = sg.td_first + lo;
            Ds.emplace_back(new DeviceBatch());
            DeviceBatch& D = *Ds.back();
            D.first = lo;
            setup_batch(D, H, o, mode, hi - lo, (starts && sg.start_rows == 0) ? starts + (size_t)2 * n * lo : starts, t1, t0,
                        path_p ? path_p + (size_t)2 * P * prow0 : nullptr, path_q ? path_q + (size_t)2 * P * prow0 : nullptr,
                        omega_mu ? omega_mu + 2 * lo : nullptr, cell_index ? cell_index + lo : nullptr, cell_weights, ncells, g2, sl);
            launch_batch(D);  // asynchronous: the next device is set up while this one tracks
            lo = hi;
        }
        const double tB = now_ms();
        double kms = 0;
        int64_t h2d_bytes = 0;
        long long ho_paths = 0; double ho_ms = 0;
        std::vector<double> ms1;
        for (auto& D : Ds) { ms1.push_back(finish_batch(*D)); h2d_bytes += D->h2d_bytes; }   // first passes done, second passes launched
        for (size_t i = 0; i < Ds.size(); ++i) {
            DeviceBatch& D = *Ds[i];
            kms = std::max(kms, ms1[i] + finish_second(D));
            ho_paths += D.handoff_paths; ho_ms = std::max(ho_ms, D.handoff_ms);
        }
        const double tC = now_ms();
        if (sc) {
#ifndef HC_HOST_SIM
            for (auto& D : Ds) {   // per-point counts of the shard, 20 bytes per point back to the host
                use_slot(D->slot);
                const long long Mloc = D->N / sc->S, j0 = D->first / sc->S;
                int* dc = D->alloc<int>((size_t)5 * Mloc);
                hc_sweep_counts_kernel<<<(unsigned)((Mloc + 127) / 128), 128, 0, cur_stream()>>>(D->A.R, n, sc->S, Mloc, sc->real_tol, dc);
                CK(cudaGetLastError());
                d2h(sc->counts + 5 * j0, dc, (size_t)5 * Mloc * sizeof(int));
            }
#else
            throw std::string("hc_track_sweep_counts is an entry point of the CUDA build");
#endif
        } else
        for (auto& D : Ds) fetch_results(*D, out);
        for (auto& D : Ds) { use_slot(D->slot); dev_sync(); }
        const double tD = now_ms();
        const DeviceBatch& D0 = *Ds[0];
        g_timing.h2d_ms = tB - tA; g_timing.kernel_ms = kms > 0 ? kms : tC - tB; g_timing.d2h_ms = tD - tC;
        g_timing.h2d_bytes = h2d_bytes; g_timing.d2h_bytes = sc ? (int64_t)(N / sc->S) * 20 : result_bytes(N, n, out->counters != nullptr);
        g_timing.grid = D0.plan.grid; g_timing.block = D0.plan.block; g_timing.lanes = D0.plan.group;
        g_timing.slab_bytes = D0.plan.engine != 0 ? (int64_t)D0.plan.lanes * (int64_t)(D0.plan.slab + D0.plan.cold) : (int64_t)D0.plan.slab;
        g_timing.devices = (int32_t)Ds.size(); g_timing.engine = D0.plan.engine;
        g_timing.handoff_paths = ho_paths; g_timing.handoff_ms = ho_ms;
        Ds.clear();
        use_slot(0);
        if (env_int("HC_B200_VERBOSE", 0) >= 2)
            fprintf(stderr, "[hc_b200] batch of %lld paths on %d device(s): setup+h2d+launch %.2f ms, wait %.2f ms (kernel %.2f ms by events), d2h %.2f ms\n",
                    N, (int)g_timing.devices, tB - tA, tC - tB, kms, tD - tC);
    } catch (const std::string& e) { return fail(e); }
    catch (const std::exception& e) { return fail(std::string("internal error: ") + e.what()); }
    catch (...) { return fail("internal error"); }
    return 0;
}

void ensure_init() {
#ifndef HC_HOST_SIM
    if (g_slots.empty() && !nodev()) throw std::string("hc_init / hc_init_devices has not been called");
#endif
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

const char* hc_last_error(void) { return g_err.c_str(); }

int32_t hc_init_devices(const int32_t* devices, int32_t n_devices) {
    std::lock_guard<std::mutex> lock(g_mutex);
#ifndef HC_HOST_SIM
    try {
        if (nodev()) return 0;
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) return fail("no CUDA device available (libhc_b200 has no CPU fallback)");
        if (!devices || n_devices < 1) return fail("device list missing");
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= count) return fail("invalid device index");
            for (int j = 0; j < i; ++j) if (devices[j] == devices[i]) return fail("device listed twice");
        }
        for (Slot& sl : g_slots) { cudaSetDevice(sl.device); if (sl.stream) cudaStreamDestroy(sl.stream); }
        g_slots.clear();
        g_pool = env_int("HC_B200_POOL", 1) != 0;
        for (int i = 0; i < n_devices; ++i) {
            Slot sl; sl.device = devices[i];
            CK(cudaSetDevice(sl.device));
            CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            cudaDeviceSetLimit(cudaLimitStackSize, (size_t)env_int("HC_B200_STACK", 8192));
            if (g_pool) {
                cudaMemPool_t mp;
                unsigned long long keep = ~0ull;
                if (cudaDeviceGetDefaultMemPool(&mp, sl.device) != cudaSuccess ||
                    cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) { cudaGetLastError(); g_pool = false; }
            }
            g_slots.push_back(sl);
        }
        if (!g_cancel) {
            CK(cudaHostAlloc((void**)&g_cancel, sizeof(int), cudaHostAllocPortable | cudaHostAllocMapped));
            *g_cancel = 0;
        }
        use_slot(0);
    } catch (const std::string& e) { return fail(e); }
#else
    (void)devices; (void)n_devices;
#endif
    return 0;
}

int32_t hc_init(int32_t device) { return hc_init_devices(&device, 1); }

int32_t hc_device_count(void) { return (int32_t)n_slots(); }

void hc_request_cancel(int32_t on) { if (g_cancel) *(volatile int*)g_cancel = on ? 1 : 0; }

void hc_options_default(hc_options* o) {
    // TrackerOptions / TrackerParameters (src/tracker.jl:45-62, 105-115)
    o->max_steps = 10000; o->max_step_size = HC_INF; o->max_initial_step_size = HC_INF; o->extended_precision = 1;
    o->min_step_size = 1e-48; o->min_rel_step_size = 0.0;
    o->a = 0.125; o->beta_a = 1.0; o->beta_omega_p = 3.0; o->beta_tau = 0.4; o->strict_beta_tau = 0.3; o->min_newton_iters = 2;
    // EndgameOptions (src/endgame_tracker.jl:47-72)
    o->endgame_start = 0.1; o->max_endgame_steps = 2000; o->max_endgame_extended_steps = 400;
    o->min_cond = 1e6; o->min_cond_growth = 1e4; o->min_coord_growth = 100.0;
    o->zero_is_at_infinity = 0; o->at_infinity_check = 1; o->only_nonsingular = 0;
    o->singular_min_accuracy = 1e-6; o->max_winding_number = 6;
    o->val_finite_tol = 0.05; o->val_at_infinity_tol = 0.01; o->sing_cond = 1e14; o->sing_accuracy = 1e-12;
    o->scaling_threshold = -30.0; o->refine_steps = 3;
    // WeightedNormOptions (src/norm.jl:36-40)
    o->scale_min = 1e-4; o->scale_abs_min = 1e-6; o->scale_max = 6.703903964971299e153;
}

void* hc_system_create(const hc_program_desc* eval, const hc_program_desc* jac) {
    std::lock_guard<std::mutex> lock(g_mutex);
    SystemH* S = nullptr;
    try {
        if (!eval || !jac) throw std::string("program descriptors missing");
        ensure_init();
        if (eval->n_vars != jac->n_vars || eval->out_dim != jac->out_dim || eval->n_params != jac->n_params)
            throw std::string("eval and Jacobian tapes disagree on dimensions");
        if (eval->out_dim != eval->n_vars) throw std::string("only square systems are supported");
        S = new SystemH();
        build_program(S->eval, eval, false);
        build_program(S->jac, jac, true);
        S->m = eval->out_dim; S->n = eval->n_vars; S->P = eval->n_params;
    } catch (const std::string& e) { delete S; fail(e); return nullptr; }
    catch (...) { delete S; fail("internal error"); return nullptr; }
    return S;
}
void hc_system_destroy(void* s) { std::lock_guard<std::mutex> lock(g_mutex); delete (SystemH*)s; }

void* hc_homotopy_create(const hc_homotopy_desc* d) {
    std::lock_guard<std::mutex> lock(g_mutex);
    HomotopyH* H = nullptr;
    try {
        if (!d || !d->F) throw std::string("homotopy needs a system F");
        if (d->kind < HC_STRAIGHT_LINE || d->kind > HC_TORIC) throw std::string("unknown homotopy kind");
        H = new HomotopyH();
        H->F = (SystemH*)d->F; H->G = (SystemH*)d->G;
        if ((int)H->F->eval.devs.size() != n_slots()) throw std::string("system was created before hc_init_devices changed the device list");
        int W = H->F->jac.dev.W, We = H->F->eval.dev.W;
        if (d->kind == HC_STRAIGHT_LINE) {
            if (!H->G) throw std::string("straight-line homotopy needs a start system G");
            if (H->G->n != H->F->n) throw std::string("G and F have different sizes");
            if (d->n_G_params != H->G->P || d->n_F_params != H->F->P) throw std::string("wrong number of fixed parameters");
            if ((d->n_G_params > 0 && !d->G_params) || (d->n_F_params > 0 && !d->F_params)) throw std::string("fixed parameters missing");
            if (H->G->jac.dev.W > W) W = H->G->jac.dev.W;
            if (H->G->eval.dev.W > We) We = H->G->eval.dev.W;
        } else {
            if (d->n_pq != H->F->P) throw std::string("wrong number of parameters");
            if (d->n_pq > 0 && !d->p) throw std::string("start parameters p missing");
            if (d->n_pq > 0 && d->kind != HC_TORIC && !d->q) throw std::string("target parameters q missing");
        }
        static const double dummy[2] = {0, 0};
        H->devs.assign((size_t)n_slots(), DevHomotopy());
        for (int sl = 0; sl < n_slots(); ++sl) {
            use_slot(sl);
            DevHomotopy& D = H->devs[(size_t)sl];
            memset(&D, 0, sizeof(D));
            D.kind = d->kind; D.n = H->F->n; D.P = H->F->P;
            D.Fe = H->F->eval.devs[(size_t)sl]; D.Fj = H->F->jac.devs[(size_t)sl];
            D.gamma = mk(d->gamma[0], d->gamma[1]);
            if (d->kind == HC_STRAIGHT_LINE) {
                D.Ge = H->G->eval.devs[(size_t)sl]; D.Gj = H->G->jac.devs[(size_t)sl];
                D.G_params = cvec_dev(d->n_G_params ? d->G_params : dummy, d->n_G_params ? d->n_G_params : 1, H->owned);
                D.F_params = cvec_dev(d->n_F_params ? d->F_params : dummy, d->n_F_params ? d->n_F_params : 1, H->owned);
            } else {
                D.p = cvec_dev(d->n_pq ? d->p : dummy, d->n_pq ? d->n_pq : 1, H->owned);
                if (d->kind != HC_TORIC) D.q = cvec_dev(d->n_pq ? d->q : dummy, d->n_pq ? d->n_pq : 1, H->owned);
            }
            // tape region (cx units): Jacobian tape, DD eval tape (2x), order-3 Taylor eval tape (4x)
            int need = W;
            if (4 * We > need) need = 4 * We;
            D.tape_cx = need;
        }
        use_slot(0);
        H->dev = H->devs[0];
    } catch (const std::string& e) { delete H; fail(e); return nullptr; }
    catch (...) { delete H; fail("internal error"); return nullptr; }
    return H;
}
void hc_homotopy_destroy(void* h) { std::lock_guard<std::mutex> lock(g_mutex); delete (HomotopyH*)h; }

int32_t hc_homotopy_set_parameters(void* Hv, const double* p, const double* q) {
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        HomotopyH* H = (HomotopyH*)Hv;
        if (!H) throw std::string("null homotopy handle");
        if (H->dev.kind == H_STRAIGHT_LINE) throw std::string("a straight-line homotopy has no start / target parameters");
        if (q && H->dev.kind == H_TORIC) throw std::string("a toric homotopy has no target parameters");
        const size_t bytes = (size_t)H->dev.P * 16;
        for (int sl = 0; sl < (int)H->devs.size(); ++sl) {
            use_slot(sl);
            if (p) h2d((void*)H->devs[(size_t)sl].p, p, bytes);
            if (q) h2d((void*)H->devs[(size_t)sl].q, q, bytes);
            dev_sync();
        }
        use_slot(0);
    } catch (const std::string& e) { return fail(e); }
    return 0;
}

int32_t hc_track_batch(void* H, const hc_options* o, int32_t mode, int64_t N, const double* starts, const double* t1,
                       const double* t0, const double* path_p, const double* path_q, const double* omega_mu, hc_results* out,
                       int32_t) {
    if (mode != MODE_ENDGAME && mode != MODE_TRACKER) return fail("mode must be 0 (endgame tracker) or 1 (tracker)");
    HomotopyH* h = (HomotopyH*)H;
    if (!h) return fail("null homotopy handle");
    if (h->dev.kind == H_TORIC) return fail("toric homotopies are tracked through hc_polyhedral_track_batch");
    return track_impl(h, o, mode, N, starts, t1, t0, path_p, path_q, omega_mu, nullptr, nullptr, 0, out);
}

int32_t hc_polyhedral_track_batch(void* Htoric, void* Hcoeff, const hc_options* o, int64_t N, const double* starts,
                                  const int32_t* cell_index, const double* cell_weights, int32_t ncells, hc_results* out, int32_t) {
    try {
        if (!Htoric || !Hcoeff) throw std::string("null homotopy handle");
        HomotopyH* h = merged_polyhedral((HomotopyH*)Htoric, (HomotopyH*)Hcoeff);
        return track_impl(h, o, MODE_POLYHEDRAL, N, starts, nullptr, nullptr, nullptr, nullptr, nullptr, cell_index, cell_weights, ncells, out);
    } catch (const std::string& e) { return fail(e); }
}

int32_t hc_polyhedral_track_cells(void* Htoric, void* Hcoeff, const hc_options* o, int64_t first, int64_t N, int32_t ncells,
                                  const int64_t* cell_volume, const int64_t* bin_H, const double* bin_mu, const double* bin_r,
                                  const double* cell_weights, hc_results* out) {
    try {
        if (!Htoric || !Hcoeff) return fail("null homotopy handle");
        if (!cell_volume || !bin_H || !bin_mu || !bin_r || !cell_weights || ncells <= 0) return fail("hc_polyhedral_track_cells: cell data missing");
        HomotopyH* h = merged_polyhedral((HomotopyH*)Htoric, (HomotopyH*)Hcoeff);
        StartGen sg; sg.td_first = first; sg.bin_volume = cell_volume; sg.bin_H = bin_H; sg.bin_mu = bin_mu; sg.bin_r = bin_r;
        return track_impl(h, o, MODE_POLYHEDRAL, N, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, cell_weights, ncells, out, sg);
    } catch (const std::string& e) { return fail(e); }
}

int32_t hc_track_total_degree(void* H, const hc_options* o, const int32_t* degrees, int64_t first, int64_t N, hc_results* out) {
    HomotopyH* h = (HomotopyH*)H;
    if (!degrees) return fail("degrees missing");
    if (h->dev.kind == H_TORIC) return fail("toric homotopies are tracked through hc_polyhedral_track_batch");
    StartGen sg; sg.degrees = degrees; sg.td_first = first;
    return track_impl(h, o, MODE_ENDGAME, N, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, out, sg);
}

int32_t hc_track_sweep(void* H, const hc_options* o, int64_t S, const double* starts, int64_t M, const double* target_params,
                       hc_results* out) {
    HomotopyH* h = (HomotopyH*)H;
    if (h->dev.kind != H_PARAMETER) return fail("hc_track_sweep needs a parameter homotopy");
    if (S <= 0 || M < 0 || !starts || !target_params) return fail("hc_track_sweep: starts and target parameters are required");
    StartGen sg; sg.start_rows = S; sg.param_div = S;
    return track_impl(h, o, MODE_ENDGAME, S * M, starts, nullptr, nullptr, nullptr, target_params, nullptr, nullptr, nullptr, 0, out, sg);
}

int32_t hc_track_sweep_counts(void* H, const hc_options* o, int64_t S, const double* starts, int64_t M, const double* target_params,
                              double real_tol, int32_t* counts) {
    HomotopyH* h = (HomotopyH*)H;
    if (!h || h->dev.kind != H_PARAMETER) return fail("hc_track_sweep_counts needs a parameter homotopy");
    if (S <= 0 || M < 0 || !starts || !target_params || !counts) return fail("hc_track_sweep_counts: starts, target parameters and counts are required");
    StartGen sg; sg.start_rows = S; sg.param_div = S;
    SweepCounts sc; sc.counts = counts; sc.S = S; sc.real_tol = real_tol > 0 ? real_tol : 1e-6;
    return track_impl(h, o, MODE_ENDGAME, S * M, starts, nullptr, nullptr, nullptr, target_params, nullptr, nullptr, nullptr, 0, nullptr, sg, &sc);
}

int32_t hc_unique_points_filter(int32_t n, int64_t M, const double* known, int64_t N, const double* cand, double atol, double rtol,
                                int64_t* match) {
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        if (n <= 0 || M < 0 || N < 0 || (M > 0 && !known) || (N > 0 && (!cand || !match))) throw std::string("hc_unique_points_filter: bad arguments");
        if (N == 0) return 0;
#ifndef HC_HOST_SIM
        if (nodev()) throw std::string("HC_B200_NO_DEVICE is set: no compute calls");
        ensure_init();
        use_slot(0);
        cx* dk = (cx*)dev_alloc((size_t)std::max<int64_t>(M, 1) * n * 16);
        cx* dc = (cx*)dev_alloc((size_t)N * n * 16);
        long long* dm = (long long*)dev_alloc((size_t)N * 8);
        h2d(dk, known, (size_t)M * n * 16); h2d(dc, cand, (size_t)N * n * 16);
        hc_unique_filter_kernel<<<(unsigned)((N + 127) / 128), 128, 0, cur_stream()>>>(n, M, dk, N, dc, atol, rtol, dm);
        cudaError_t le = cudaGetLastError();
        d2h(match, dm, (size_t)N * 8);
        dev_sync();
        dev_free(dk); dev_free(dc); dev_free(dm);
        CK(le);
#else
        for (int64_t i = 0; i < N; ++i) {
            const double* c = cand + 2 * n * i;
            double nrm2 = 0;
            for (int j = 0; j < 2 * n; ++j) nrm2 += c[j] * c[j];
            double rad = rtol * sqrt(nrm2);
            if (rad < atol) rad = atol;
            match[i] = -1;
            for (int64_t k = 0; k < M && match[i] < 0; ++k) {
                double d = 0;
                for (int j = 0; j < 2 * n; ++j) { const double e = known[2 * n * k + j] - c[j]; d += e * e; }
                if (d <= rad * rad) match[i] = k;
            }
        }
#endif
    } catch (const std::string& e) { return fail(e); }
    catch (...) { return fail("internal error"); }
    return 0;
}

void hc_get_timing(hc_timing* t) { *t = g_timing; }

void* hc_resident_create(void* H, void* Hcoeff, const hc_options* o, int32_t mode, int64_t N, const double* starts, const double* t1,
                         const double* t0, const double* path_p, const double* path_q, const int32_t* cell_index,
                         const double* cell_weights, int32_t ncells) {
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceBatch* D = nullptr;
    try {
        if (!H || !o) throw std::string("null handle / options");
        ensure_init();
        HomotopyH* h = (HomotopyH*)H;
        if (mode == MODE_POLYHEDRAL) h = merged_polyhedral((HomotopyH*)H, (HomotopyH*)Hcoeff);
        if (g_cancel) *g_cancel = 0;
        D = new DeviceBatch();
        setup_batch(*D, h, o, mode, N, starts, t1, t0, path_p, path_q, nullptr, cell_index, cell_weights, ncells);
        dev_sync();
    } catch (const std::string& e) { delete D; fail(e); return nullptr; }
    catch (...) { delete D; fail("internal error"); return nullptr; }
    return D;
}
void* hc_resident_create_cells(void* Htoric, void* Hcoeff, const hc_options* o, int64_t first, int64_t N, int32_t ncells,
                               const int64_t* cell_volume, const int64_t* bin_H, const double* bin_mu, const double* bin_r,
                               const double* cell_weights) {
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceBatch* D = nullptr;
    try {
        if (!Htoric || !Hcoeff || !o) throw std::string("null handle / options");
        if (!cell_volume || !bin_H || !bin_mu || !bin_r || !cell_weights || ncells <= 0) throw std::string("hc_resident_create_cells: cell data missing");
        ensure_init();
        HomotopyH* h = merged_polyhedral((HomotopyH*)Htoric, (HomotopyH*)Hcoeff);
        if (g_cancel) *g_cancel = 0;
        StartGen sg; sg.td_first = first; sg.bin_volume = cell_volume; sg.bin_H = bin_H; sg.bin_mu = bin_mu; sg.bin_r = bin_r;
        D = new DeviceBatch();
        setup_batch(*D, h, o, MODE_POLYHEDRAL, N, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, cell_weights, ncells, sg);
        dev_sync();
    } catch (const std::string& e) { delete D; fail(e); return nullptr; }
    catch (...) { delete D; fail("internal error"); return nullptr; }
    return D;
}
int32_t hc_resident_run(void* r, double* kernel_ms) {
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        DeviceBatch& D = *(DeviceBatch*)r;
        double ms = run_batch(D);
        if (kernel_ms) *kernel_ms = ms;
        g_timing.handoff_paths = D.handoff_paths; g_timing.handoff_ms = D.handoff_ms;
    }
    catch (const std::string& e) { return fail(e); }
    catch (...) { return fail("internal error"); }
    return 0;
}
int32_t hc_resident_fetch(void* r, hc_results* out) {
    std::lock_guard<std::mutex> lock(g_mutex);
    try { fetch_results(*(DeviceBatch*)r, out); dev_sync(); } catch (const std::string& e) { return fail(e); }
    catch (...) { return fail("internal error"); }
    return 0;
}
void hc_resident_destroy(void* r) { std::lock_guard<std::mutex> lock(g_mutex); delete (DeviceBatch*)r; }

// ------------------------------------------------------------------ operator API hooks
static int hook(void* Hv, int what, int K, const double* x, const double* xlo, const double* t, double* u, double* U) {
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        HomotopyH* H = (HomotopyH*)Hv;
        if (!H || !x || !t || !u) throw std::string("null argument");
        ensure_init();
        use_slot(0);
        const int n = H->dev.n, P = H->dev.P;
        hc_options o; hc_options_default(&o);
        KArgs A; memset(&A, 0, sizeof(A));
        A.H = H->dev; A.H.N = 1; A.O = to_dev_options(&o);
        PathMem<0> dummy;
        A.H.tape_cx = 5 * (H->F->eval.dev.W > (H->G ? H->G->eval.dev.W : 0) ? H->F->eval.dev.W : H->G->eval.dev.W);  // order-4 series
        if (A.H.tape_cx < H->dev.tape_cx) A.H.tape_cx = H->dev.tape_cx;
        const SlabSizes ss = carve(dummy, n, P, A.H.tape_cx, nullptr, nullptr);
        const size_t slab = ss.hot;
        std::vector<void*> owned;
        auto A_ = [&](size_t b) { void* p = dev_alloc(b); owned.push_back(p); return p; };
        A.cold = (unsigned char*)A_(ss.cold);
        const int nx = what == 3 ? K * n : n;
        cx* dx = (cx*)A_((size_t)nx * 16); h2d(dx, x, (size_t)nx * 16);
        cx* dlo = nullptr;
        if (xlo) { dlo = (cx*)A_((size_t)n * 16); h2d(dlo, xlo, (size_t)n * 16); }
        double* dtw = nullptr;
        if (!H->tw_hook.empty()) { dtw = (double*)A_((size_t)P * 8); h2d(dtw, H->tw_hook.data(), (size_t)P * 8); }
        cx* du = (cx*)A_((size_t)n * 16);
        cx* dU = (cx*)A_((size_t)n * n * 16);
        cx tt = mk(t[0], t[1]);
        // HC_B200_JIT=1: the operator API runs on the generated code of the specialised kernels (what the tracker
        // executes there); evaluate_dd and order-4 series only exist in the interpreter
        const char* jenv = getenv("HC_B200_JIT");
        if (jenv && !strcmp(jenv, "1") && what != 1 && !(what == 3 && K > 3) && jit_wanted(*H, 1, t[1] != 0.0)) {
            std::shared_ptr<jit::Module> M = jit_module(*H, false);
            A.H.tape_cx = jit_tape_cx(*H);
            unsigned char* jcold = (unsigned char*)A_(M->cold);
            A.cold = jcold;
#ifndef HC_HOST_SIM
            if (M->hot > kSmemMax) throw std::string("system too large for one shared-memory slab");
            CK(cudaFuncSetAttribute((const void*)M->hook, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
            void* args[] = {(void*)&A, (void*)&what, (void*)&K, (void*)&dx, (void*)&tt, (void*)&dtw, (void*)&du, (void*)&dU};
            CK(cudaLaunchKernel((const void*)M->hook, dim3(1), dim3(1), args, M->hot, cur_stream()));
            CK(cudaGetLastError());
            dev_sync();
#else
            std::vector<unsigned char> hot(M->hot + 16);
            M->hook(&A, (unsigned char*)(((uintptr_t)hot.data() + 15) & ~(uintptr_t)15), jcold, what, K, dx, tt, dtw, du, dU);
#endif
            d2h(u, du, (size_t)n * 16);
            if (U) d2h(U, dU, (size_t)n * n * 16);
            dev_sync();
            for (void* p : owned) dev_free(p);
            return 0;
        }
#ifndef HC_HOST_SIM
        if (slab > kSmemMax) throw std::string("system too large for one shared-memory slab");
        static bool attr_set = false;
        if (!attr_set) { CK(cudaFuncSetAttribute(hc_hook_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)); attr_set = true; }
        hc_hook_kernel<<<1, 1, slab, cur_stream()>>>(A, what, K, dx, dlo, tt, dtw, du, dU);
        CK(cudaGetLastError());
        dev_sync();
#else
        std::vector<unsigned char> mem(slab + 16);
        Lane<1, 0> L;
        L.g.init();
        L.H = &A.H; L.O = &A.O; L.n = n; L.pidx = L.prow = 0; L.kind = A.H.kind;
        carve(L.M, n, P, A.H.tape_cx, (unsigned char*)(((uintptr_t)mem.data() + 15) & ~(uintptr_t)15), A.cold);
        L.n_evaljac = L.n_eval = L.n_evaldd = L.n_tay1 = L.n_tay2 = L.n_tay3 = 0; L.tape_prog = L.tay_prog = nullptr; L.a_in_lu = L.rs_raw = false;
        if (dtw) for (int i = 0; i < P; ++i) L.M.tw[i] = dtw[i];
        if (what == 3) for (int i = 0; i < K * n; ++i) L.M.tx[i] = dx[i]; else for (int i = 0; i < n; ++i) L.M.x[i] = dx[i];
        if (what == 0) L.eval_f64(L.M.u, nullptr, L.M.x, tt);
        else if (what == 1) { for (int i = 0; i < n; ++i) L.M.xhat[i] = dlo[i]; L.eval_dd(L.M.u, L.M.x, &L.M.xhat, tt); }
        else if (what == 2) L.eval_f64(L.M.u, &L.M.A, L.M.x, tt);
        else if (K == 1) L.taylor<1>(L.M.u, L.M.tx, tt);
        else if (K == 2) L.taylor<2>(L.M.u, L.M.tx, tt);
        else if (K == 3) L.taylor<3>(L.M.u, L.M.tx, tt);
        else L.taylor<4>(L.M.u, L.M.tx, tt);
        for (int i = 0; i < n; ++i) du[i] = L.M.u[i];
        if (what == 2) for (int i = 0; i < n * n; ++i) dU[i] = L.M.A[i];
#endif
        d2h(u, du, (size_t)n * 16);
        if (U) d2h(U, dU, (size_t)n * n * 16);
        dev_sync();
        for (void* p : owned) dev_free(p);
    } catch (const std::string& e) { return fail(e); }
    catch (...) { return fail("internal error"); }
    return 0;
}
int32_t hc_evaluate(void* H, const double* x, const double* t, double* u) { return hook(H, 0, 0, x, nullptr, t, u, nullptr); }
int32_t hc_evaluate_dd(void* H, const double* x_hi, const double* x_lo, const double* t, double* u) { return hook(H, 1, 0, x_hi, x_lo, t, u, nullptr); }
int32_t hc_evaluate_and_jacobian(void* H, const double* x, const double* t, double* u, double* U) { return hook(H, 2, 0, x, nullptr, t, u, U); }
int32_t hc_taylor(void* H, int32_t K, const double* tx, const double* t, double* u) {
    if (K < 1 || K > 4) return fail("taylor order must be 1..4");
    return hook(H, 3, K, tx, nullptr, t, u, nullptr);
}
// evaluate! / evaluate_and_jacobian! at N points at once (the batched form of the operator API, SURVEY.md section 7 "minimum
// slice"; what a host uses to check the residuals of many endpoints, e.g. the excess-solution check of an overdetermined
// solve): x is n x N, u n x N, U (optional) n x n x N, one t for all points.
int32_t hc_evaluate_batch(void* Hv, int64_t N, const double* x, const double* t, double* u, double* U) {
    if (N < 0 || !Hv || !x || !t || !u) return fail("hc_evaluate_batch: bad arguments");
    HomotopyH* H = (HomotopyH*)Hv;
    const int n = H->dev.n;
#ifdef HC_HOST_SIM
    for (int64_t i = 0; i < N; ++i) {
        const int rc = hook(Hv, U ? 2 : 0, 0, x + 2 * (size_t)n * i, nullptr, t, u + 2 * (size_t)n * i, U ? U + 2 * (size_t)n * n * i : nullptr);
        if (rc) return rc;
    }
    return 0;
#else
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        if (nodev()) throw std::string("HC_B200_NO_DEVICE is set: no compute calls");
        ensure_init();
        use_slot(0);
        const int P = H->dev.P;
        hc_options o; hc_options_default(&o);
        KArgs A; memset(&A, 0, sizeof(A));
        A.H = H->dev; A.H.N = 1; A.O = to_dev_options(&o);
        PathMem<0> dummy;
        const SlabSizes ss = carve(dummy, n, P, A.H.tape_cx, nullptr, nullptr);
        if (ss.hot > kSmemMax) throw std::string("system too large for one shared-memory slab");
        if (first_use((const void*)hc_hook_kernel)) CK(cudaFuncSetAttribute(hc_hook_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
        const int64_t chunk = 8192;
        const int64_t cap = N < chunk ? N : chunk;
        if (cap == 0) return 0;
        unsigned char* cold = (unsigned char*)dev_alloc((size_t)cap * ss.cold);
        cx* dx = (cx*)dev_alloc((size_t)cap * n * 16);
        cx* du = (cx*)dev_alloc((size_t)cap * n * 16);
        cx* dU = U ? (cx*)dev_alloc((size_t)cap * n * n * 16) : nullptr;
        A.cold = cold; A.cold_bytes = (int)ss.cold;
        const cx tt = mk(t[0], t[1]);
        cudaError_t le = cudaSuccess;
        for (int64_t lo = 0; lo < N && le == cudaSuccess; lo += chunk) {
            const int64_t cnt = N - lo < chunk ? N - lo : chunk;
            h2d(dx, x + 2 * (size_t)n * lo, (size_t)cnt * n * 16);
            hc_hook_kernel<<<(unsigned)cnt, 1, ss.hot, cur_stream()>>>(A, U ? 2 : 0, 0, dx, nullptr, tt, nullptr, du, dU);
            le = cudaGetLastError();
            d2h(u + 2 * (size_t)n * lo, du, (size_t)cnt * n * 16);
            if (U) d2h(U + 2 * (size_t)n * n * lo, dU, (size_t)cnt * n * n * 16);
            dev_sync();
        }
        dev_free(cold); dev_free(dx); dev_free(du); if (dU) dev_free(dU);
        CK(le);
    } catch (const std::string& e) { return fail(e); }
    catch (...) { return fail("internal error"); }
    return 0;
#endif
}
int32_t hc_toric_set_weights(void* Hv, const double* w) {
    HomotopyH* H = (HomotopyH*)Hv;
    H->tw_hook.assign(w, w + H->dev.P);
    return 0;
}

int32_t hc_jit_prepare(void* Hv, int32_t flags, double* info) {
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        HomotopyH* H = (HomotopyH*)Hv;
        if (!H) throw std::string("null homotopy handle");
        std::shared_ptr<jit::Module> M = jit_module(*H, (flags & 1) != 0, false, (flags & 2) != 0);
        if (info) { info[0] = (double)M->cubin_bytes; info[1] = M->compile_ms; info[2] = (double)M->hot; info[3] = (double)M->cold; info[4] = M->from_cache ? 1.0 : 0.0; }
    } catch (const std::string& e) { return fail(e); }
    return 0;
}

int32_t hc_host_register(void* p, int64_t bytes) {
#ifndef HC_HOST_SIM
    if (!p || bytes <= 0) return 0;
    cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
#else
    (void)p; (void)bytes;
#endif
    return 0;
}
int32_t hc_host_unregister(void* p) {
#ifndef HC_HOST_SIM
    if (!p) return 0;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); }
#else
    (void)p;
#endif
    return 0;
}

#ifdef HC_HOST_SIM
// Test hook (tests/host_sim only): lowers a tape with the given segment window and checks the invariants the
// segment interpreters rely on.  Returns 0 if they hold, a positive code otherwise; stats = {ops, segments, W}.
int32_t hc_sim_check_lowering(const hc_program_desc* d, int32_t window, int32_t* stats) {
    try {
        const LoweredProgram L = lower_program(d, 0, true, false, window);
        const int Nin = L.var_off + L.n;
        std::vector<char> written((size_t)L.W, 0);
        for (int i = 0; i < Nin; ++i) written[i] = 1;
        size_t f = 0, op = 0;
        long long total = 0;
        for (const int2& sg : L.segs) {
            const int cls = sg.x & 7;
            std::vector<uint32_t> outs;
            for (int k = 0; k < sg.y; ++k, ++op) {
                if (op >= L.ops.size() || f >= L.fops.size()) return 1;
                const MOp& m = L.ops[op];
                const FOp& F = L.fops[f];
                if ((int)((m.w0 >> 16) & 31) != sg.x) return 2;                       // segment key == op key
                if (F.out != (m.w0 & 0xffff) || F.a != (m.w1 & 0xffff)) return 3;     // fast format mirrors the packed one
                const bool useB = cls == MC_MM || cls == MC_MA || cls == MC_M || cls == MC_DIV;
                const bool useC = cls == MC_MM || cls == MC_MA || cls == MC_AA;
                std::vector<uint32_t> ins{F.a};
                if (useB) { if (F.b != (m.w1 >> 16)) return 3; ins.push_back(F.b); }
                if (useC) { if (F.c != (m.w2 & 0xffff)) return 3; ins.push_back(F.c); }
                if (cls == MC_MM) { if (L.fops[f + 1].a != (m.w2 >> 16)) return 3; ins.push_back(L.fops[f + 1].a); }
                for (uint32_t s_ : ins) {
                    if (s_ >= (uint32_t)L.W || !written[s_]) return 4;                // operands exist before the segment...
                    for (uint32_t o : outs) if (o == s_) return 5;                    // ...and are not produced inside it
                }
                if (F.out < (uint32_t)Nin || F.out >= (uint32_t)L.W) return 6;        // never writes the input block
                for (uint32_t o : outs) if (o == F.out) return 7;                     // no two ops of a segment share a slot
                outs.push_back(F.out);
                f += cls == MC_MM ? 2 : 1;
            }
            for (uint32_t o : outs) written[o] = 1;
            total += sg.y;
        }
        if (op != L.ops.size() || f != L.fops.size() || total != (long long)L.ops.size()) return 8;
        for (const int2& a : L.u_assign) if (a.y < 0 || a.y >= L.W || !written[a.y]) return 9;
        for (const int2& a : L.U_assign) if (a.y < 0 || a.y >= L.W || !written[a.y]) return 9;
        if (stats) { stats[0] = (int32_t)L.ops.size(); stats[1] = (int32_t)L.segs.size(); stats[2] = L.W; }
        return 0;
    } catch (const std::string& e) { fail(e); return -1; }
}
#endif

double hc_dfma_peak(int32_t iters) {
#ifndef HC_HOST_SIM
    try {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int block = 256, grid = sms * 8;
        double* out = (double*)dev_alloc((size_t)grid * block * 8);
        hc_dfma_kernel<<<grid, block>>>(out, 1000);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, 0));
        hc_dfma_kernel<<<grid, block>>>(out, iters);
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        dev_free(out);
        return 2.0 * 8.0 * (double)iters * grid * block / (ms * 1e-3) / 1e9;
    } catch (const std::string& e) { fail(e); return -1.0; }
#else
    (void)iters;
    return -1.0;
#endif
}

}  // extern "C"
