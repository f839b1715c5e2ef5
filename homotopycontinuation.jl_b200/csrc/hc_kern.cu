// The ahead-of-time tracker kernels (interpreter engines).  One translation unit per instantiation: hc_kern.cu is
// compiled with -DHC_KERN_TPL=<slab bytes> or -DHC_KERN_GROUP=<lanes per path>, so that the instantiations build in
// parallel (each is 0.2 - 0.4 MB of SASS); hc_api.cu reaches them through hc_kernel_tpl / hc_kernel_group.
#include "hc_kernel.h"

namespace hc {

// Persistent tracker: each group of G lanes owns one shared-memory slab, pulls path indices from
// the device-side queue (the `next_k` counter of threaded_solve, src/solve.jl:641, 660-667) and
// writes its PathResult by path index (src/solve.jl:637, 670).
template <int G>
__global__ void __launch_bounds__(256, 2) hc_track_kernel(const __grid_constant__ KArgs A) {
    __shared__ KArgs sA;
    if (threadIdx.x == 0) sA = A;
    __syncthreads();
    if (A.stage) {
        DevHomotopy h = A.H;  // every thread computes the same pointers
        unsigned char* cur = hc_smem;
        stage_program(h.Fe, cur);
        stage_program(h.Fj, cur);
        if (h.kind == H_STRAIGHT_LINE) { stage_program(h.Ge, cur); stage_program(h.Gj, cur); }
        stage_params(h, cur);
        __syncthreads();
        if (threadIdx.x == 0) sA.H = h;
        __syncthreads();
    }
    Lane<G, 0> L;
    L.g.init();
    L.H = &sA.H; L.O = &sA.O; L.n = sA.H.n;
    carve(L.M, sA.H.n, sA.H.P, sA.H.tape_cx, hc_smem + A.stage_bytes + (size_t)(threadIdx.x / G) * A.slab_bytes,
          A.cold + ((size_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G) * A.cold_bytes);
    L.phase = PH_IDLE; L.ev = EV_START; L.stop_pending = false;
    const long long N = sA.B.N;
    while (true) {
        if (L.ev != EV_NONE) {  // a lane group handles its events at once (all its lanes are in the same state)
            L.event_finish(sA.R);
            long long k = -1;
            if (L.ev == EV_START) {
                if (L.g.lane == 0) k = (long long)atomicAdd(A.queue, 1ULL);
                k = L.g.bcast(k, 0);
                if (k >= N) break;
            }
            L.event_begin(k, k >= 0, sA.B, sA.R);
        }
        if (L.ev == EV_NONE && L.phase != PH_IDLE) L.iterate(sA.B, sA.R);
    }
}

// Lockstep variant of the lane-group kernel (second pass of a two-pass batch): the warps of a CTA go through event
// handling and one tracker step per round together (CTA-wide barriers inside tracker_step_t<true>), so that they fetch
// the same code at the same time -- the unsynchronised kernel spends 55 % of its stall samples waiting for instructions
// (profiles/r02b_ncu_pass2_group32_cyclooctane*.txt).  A group without work keeps the barriers company until the CTA is done.
#ifndef HC_KERN_SYNC_MIN_CTAS
#define HC_KERN_SYNC_MIN_CTAS 2   // (1 = 190 registers instead of 128: measured, no gain)
#endif
template <int G>
__global__ void __launch_bounds__(256, HC_KERN_SYNC_MIN_CTAS) hc_track_kernel_sync(const __grid_constant__ KArgs A) {
    __shared__ KArgs sA;
    if (threadIdx.x == 0) sA = A;
    __syncthreads();
    if (A.stage) {
        DevHomotopy h = A.H;
        unsigned char* cur = hc_smem;
        stage_program(h.Fe, cur);
        stage_program(h.Fj, cur);
        if (h.kind == H_STRAIGHT_LINE) { stage_program(h.Ge, cur); stage_program(h.Gj, cur); }
        stage_params(h, cur);
        __syncthreads();
        if (threadIdx.x == 0) sA.H = h;
        __syncthreads();
    }
    Lane<G, 0> L;
    L.g.init();
    L.H = &sA.H; L.O = &sA.O; L.n = sA.H.n;
    carve(L.M, sA.H.n, sA.H.P, sA.H.tape_cx, hc_smem + A.stage_bytes + (size_t)(threadIdx.x / G) * A.slab_bytes,
          A.cold + ((size_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G) * A.cold_bytes);
    L.phase = PH_IDLE; L.ev = EV_START; L.stop_pending = false;
    const long long N = sA.B.N;
    bool drained = false;  // group-uniform
    while (true) {
        const int npend = __syncthreads_count(L.ev != EV_NONE);
        const int nact = __syncthreads_count(L.ev == EV_NONE && L.phase != PH_IDLE);
        if (npend == 0 && nact == 0) break;
        if (npend * 4 >= (int)blockDim.x || nact == 0) {  // a quarter of the groups wait, or nobody can step
            if (L.ev != EV_NONE) {
                L.event_finish(sA.R);
                long long k = -1;
                if (L.ev == EV_START && !drained) {
                    if (L.g.lane == 0) k = (long long)atomicAdd(A.queue, 1ULL);
                    k = L.g.bcast(k, 0);
                    if (k >= N) { drained = true; k = -1; }
                }
                L.event_begin(k, k >= 0, sA.B, sA.R);
            }
        }
        L.template iterate_t<true>(L.ev == EV_NONE && L.phase != PH_IDLE);
    }
}

// Thread-per-path engine (n <= 14): every lane tracks its own path, the programs sit in shared memory,
// the lane state in LOCAL memory: the hardware interleaves the lanes of
// a warp (a warp access to element i is one contiguous 512 B segment, as in the explicit slabs
// above), element addresses are base + immediate (no per-access stride multiply), and L1 keeps
// local lines write-back, so the state that a step re-reads stays on the SM.
template <int SLAB>
__global__ void __launch_bounds__(256) hc_track_tpl_kernel(const __grid_constant__ KArgs A) {
    __shared__ KArgs sA;
    if (threadIdx.x == 0) sA = A;
    __syncthreads();
    if (A.stage) {
        DevHomotopy h = A.H;
        unsigned char* cur = hc_smem;
        stage_program(h.Fe, cur, 1);
        stage_program(h.Fj, cur, 1);
        if (h.kind == H_STRAIGHT_LINE) { stage_program(h.Ge, cur, 1); stage_program(h.Gj, cur, 1); }
        stage_params(h, cur);
        __syncthreads();
        if (threadIdx.x == 0) sA.H = h;
        __syncthreads();
    }
    __align__(16) unsigned char slab[SLAB];
    Lane<1, 2> L;
    L.g.init();
    L.H = &sA.H; L.O = &sA.O; L.n = sA.H.n;
    carve(L.M, sA.H.n, sA.H.P, sA.H.tape_cx, slab, slab + A.slab_bytes);
    tpp_loop(L, A, sA);
}


#if defined(HC_KERN_TPL)
#define HC_CAT2(a, b) a##b
#define HC_CAT(a, b) HC_CAT2(a, b)
const void* HC_CAT(hc_kernel_tpl_, HC_KERN_TPL)() { return (const void*)hc_track_tpl_kernel<HC_KERN_TPL>; }
#elif defined(HC_KERN_GROUP)
#define HC_CAT2(a, b) a##b
#define HC_CAT(a, b) HC_CAT2(a, b)
const void* HC_CAT(hc_kernel_group_, HC_KERN_GROUP)() { return (const void*)hc_track_kernel<HC_KERN_GROUP>; }
#elif defined(HC_KERN_GROUP_SYNC)
#define HC_CAT2(a, b) a##b
#define HC_CAT(a, b) HC_CAT2(a, b)
const void* HC_CAT(hc_kernel_group_sync_, HC_KERN_GROUP_SYNC)() { return (const void*)hc_track_kernel_sync<HC_KERN_GROUP_SYNC>; }
#else
// dispatcher unit
const void* hc_kernel_tpl_12288(); const void* hc_kernel_tpl_24576(); const void* hc_kernel_tpl_49152();
const void* hc_kernel_group_8(); const void* hc_kernel_group_32();
const void* hc_kernel_group_sync_8(); const void* hc_kernel_group_sync_32();
const void* hc_kernel_group_sync(int G) { return G == 8 ? hc_kernel_group_sync_8() : hc_kernel_group_sync_32(); }
const void* hc_kernel_tpl(int slab) { return slab == 12288 ? hc_kernel_tpl_12288() : (slab == 24576 ? hc_kernel_tpl_24576() : hc_kernel_tpl_49152()); }
const void* hc_kernel_group(int G) { return G == 8 ? hc_kernel_group_8() : hc_kernel_group_32(); }
#endif

}  // namespace hc
