// Body of a per-system specialised translation unit.  hc_jit.h assembles
//
//     #define HC_JIT_N 7              // number of variables: every loop over them gets a constant trip count
//     #define HC_JIT_SLAB 7424        // bytes of lane state (hot + cold) in local memory
//     #define HC_JIT_BLOCK 128
//     #define HC_JIT_SYNC 1           // lockstep rounds (tpp_loop_sync)
//     #define HC_JIT_LU_SMEM 1        // LU factors thread-interleaved in shared memory (SV<cx, 3>)
//     #define HC_JIT_GEN "hc_jit_gen.inc"   // evaluate / evaluate_and_jacobian / taylor of THIS system (hc_jitgen.h)
//     #include "hc_jit_unit.h"
//
// and compiles it for sm_100a with NVRTC (or, in tests/host_sim, with g++).  The tracker logic is the one of
// hc_path.h / hc_lane.h; what changes is that the homotopy is straight-line code with its tape slots in registers,
// so a lane keeps no fp64 / Taylor tape in memory: the state a step touches shrinks from ~9 KB to ~3 KB per lane.
#pragma once
#include "hc_kernel.h"

namespace hc {

#if defined(__CUDACC__)
// Thread-per-path engine, one lane = one path, lane state in local memory (see hc_track_tpl_kernel in hc_api.cu).
extern "C" __global__ void __launch_bounds__(HC_JIT_BLOCK, 1) hc_jit_track(const __grid_constant__ KArgs A) {
    __shared__ KArgs sA;
    if (threadIdx.x == 0) sA = A;
    __syncthreads();
    {
        DevHomotopy h = A.H;  // every thread computes the same pointers
        unsigned char* cur = hc_smem;
        stage_program(h.Fe, cur, 2);
        if (h.kind == H_STRAIGHT_LINE) stage_program(h.Ge, cur, 2);
        stage_params(h, cur);
        __syncthreads();
        if (threadIdx.x == 0) sA.H = h;
        __syncthreads();
    }
    __align__(16) unsigned char slab[HC_JIT_SLAB];
    Lane<1, 2> L;
    L.g.init();
    L.H = &sA.H; L.O = &sA.O;
#if HC_JIT_LU_SMEM
    carve(L.M, HC_JIT_N, sA.H.P, sA.H.tape_cx, slab, slab + A.slab_bytes, 3);
    L.M.LU.p = (unsigned)__cvta_generic_to_shared(hc_smem + A.stage_bytes) + 16u * threadIdx.x;  // thread-interleaved (SV<cx, 3>)
#else
    carve(L.M, HC_JIT_N, sA.H.P, sA.H.tape_cx, slab, slab + A.slab_bytes, 1);
#endif
#if HC_JIT_SYNC
    tpp_loop_sync(L, A, sA);
#else
    tpp_loop(L, A, sA);
#endif
}

// operator API of the generated code at a single point (test hook): what = 0 evaluate, 2 evaluate_and_jacobian,
// 3 taylor of order K; one thread, state behind generic pointers
extern "C" __global__ void hc_jit_hook(const KArgs A, int what, int K, const cx* x, cx t, const double* tw, cx* u, cx* U) {
    __shared__ KArgs sA;
    sA = A;
    Lane<1, 0> L;
    L.g.init();
    L.H = &sA.H; L.O = &sA.O; L.pidx = L.prow = 0; L.kind = sA.H.kind;
    carve(L.M, HC_JIT_N, sA.H.P, sA.H.tape_cx, hc_smem, A.cold, true);
    L.n_evaljac = L.n_eval = L.n_evaldd = L.n_tay1 = L.n_tay2 = L.n_tay3 = 0; L.a_in_lu = L.rs_raw = false; L.tape_prog = L.tay_prog = nullptr; L.pv_kind = L.ps_kind = -1; L.jit_bind_params();
    const int n = HC_JIT_N;
    if (tw) for (int i = 0; i < sA.H.P; ++i) L.M.tw[i] = tw[i];
    if (what == 3) { for (int i = 0; i < K * n; ++i) L.M.tx[i] = x[i]; }
    else for (int i = 0; i < n; ++i) L.M.x[i] = x[i];
    if (what == 0) L.eval_f64(L.M.u, nullptr, L.M.x, t);
    else if (what == 2) L.eval_f64(L.M.u, &L.M.A, L.M.x, t);
    else if (K == 1) L.template taylor<1>(L.M.u, L.M.tx, t);
    else if (K == 2) L.template taylor<2>(L.M.u, L.M.tx, t);
    else L.template taylor<3>(L.M.u, L.M.tx, t);
    for (int i = 0; i < n; ++i) u[i] = L.M.u[i];
    if (what == 2) for (int i = 0; i < n * n; ++i) U[i] = L.M.A[i];
}
#else
// host-compiled build of the same unit (tests/host_sim): sequential driver
extern "C" void hc_jit_sim_track(const KArgs* A, unsigned char* hot, unsigned char* cold) {
    Lane<1, 0> L;
    L.g.init();
    L.H = &A->H; L.O = &A->O;
    carve(L.M, HC_JIT_N, A->H.P, A->H.tape_cx, hot, cold, true);
    sim_loop(L, *A);
}
extern "C" void hc_jit_sim_hook(const KArgs* A, unsigned char* hot, unsigned char* cold, int what, int K, const cx* x, cx t, const double* tw, cx* u, cx* U) {
    Lane<1, 0> L;
    L.g.init();
    L.H = &A->H; L.O = &A->O; L.pidx = L.prow = 0; L.kind = A->H.kind;
    carve(L.M, HC_JIT_N, A->H.P, A->H.tape_cx, hot, cold, true);
    L.n_evaljac = L.n_eval = L.n_evaldd = L.n_tay1 = L.n_tay2 = L.n_tay3 = 0; L.a_in_lu = L.rs_raw = false; L.tape_prog = L.tay_prog = nullptr; L.pv_kind = L.ps_kind = -1; L.jit_bind_params();
    const int n = HC_JIT_N;
    if (tw) for (int i = 0; i < A->H.P; ++i) L.M.tw[i] = tw[i];
    if (what == 3) { for (int i = 0; i < K * n; ++i) L.M.tx[i] = x[i]; }
    else for (int i = 0; i < n; ++i) L.M.x[i] = x[i];
    if (what == 0) L.eval_f64(L.M.u, nullptr, L.M.x, t);
    else if (what == 2) L.eval_f64(L.M.u, &L.M.A, L.M.x, t);
    else if (K == 1) L.taylor<1>(L.M.u, L.M.tx, t);
    else if (K == 2) L.taylor<2>(L.M.u, L.M.tx, t);
    else L.taylor<3>(L.M.u, L.M.tx, t);
    for (int i = 0; i < n; ++i) u[i] = L.M.u[i];
    if (what == 2) for (int i = 0; i < n * n; ++i) U[i] = L.M.A[i];
}
#endif

}  // namespace hc
