// Kernel-side plumbing shared by the ahead-of-time kernels of hc_api.cu and the per-system specialised
// kernels that hc_jit.h compiles at run time (NVRTC): launch arguments, staging of the programs into shared
// memory, and the main loop of the thread-per-path engine.
//
// The loop is the device-side form of the `next_k` work counter of threaded_solve (reference
// src/solve.jl:641, 660-667): lanes pull path indices from an atomic queue and write their PathResult by
// path index (src/solve.jl:637, 670).
#pragma once
#include "hc_lane.h"

namespace hc {

struct KArgs {
    DevHomotopy H;
    DevOptions O;
    BatchIn B;
    DevResults R;
    unsigned long long* queue;
    int stage;            // 1: copy the programs into shared memory
    int slab_bytes;       // per-path shared-memory slab (lane groups) / hot part of the local slab (thread per path)
    int cold_bytes;       // per-group scratch in global memory
    unsigned char* cold;
    int refill_min;       // thread-per-path engine: idle lanes of a warp refill together once this many wait
    int stage_bytes;      // staged programs (0 when !stage)
    const int* cancel;    // optional host-mapped flag: non-zero = stop handing out new paths (src/solve.jl:685-707)
    int sync_cta;         // thread-per-path engines: the warps of a CTA start every round together (instruction-cache sharing)
    float refill_k;       // lockstep kernels, refill_min <= 0 (adaptive refill threshold, tpp_loop_sync): scale of the cost model
};

#if defined(__CUDACC__)
extern __shared__ __align__(16) unsigned char hc_smem[];

template <class T>
__device__ const T* stage_array(const T* src, int count, unsigned char*& cur) {
    size_t bytes = ((size_t)count * sizeof(T) + 15) & ~(size_t)15;
    T* dst = reinterpret_cast<T*>(cur);
    const int words = (int)(bytes / 4);
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    const int valid = (int)(((size_t)count * sizeof(T)) / 4);
    for (int i = threadIdx.x; i < words; i += blockDim.x) d[i] = i < valid ? s[i] : 0u;
    cur += bytes;
    return dst;
}
// mode 0: lane-group kernels (packed ops in shared memory); 1: thread-per-path interpreter kernels (fops + segs in
// shared memory, the packed ops -- only read by the DoubleDouble interpreter -- stay in global memory);
// 2: specialised kernels (no fp64 / Taylor interpreter at all: only what the DoubleDouble interpreter reads)
__device__ inline void stage_program(DevProgram& P, unsigned char*& cur, int mode = 0) {
    if (mode == 1) {
        P.fops = stage_array(P.fops, P.n_fops, cur);
        P.segs = stage_array(P.segs, P.n_segs, cur);
    } else if (mode == 0) P.ops = stage_array(P.ops, P.n_ops, cur);
    P.level_end = stage_array(P.level_end, P.n_levels, cur);
    P.consts = stage_array(P.consts, P.C, cur);
    P.u_assign = stage_array(P.u_assign, P.nu, cur);
    if (mode != 2) P.U_assign = stage_array(P.U_assign, P.nU, cur);
}

// homotopy-level constant arrays: start / target parameters, fixed parameters of F and G
__device__ inline void stage_params(DevHomotopy& h, unsigned char*& cur) {
    if (h.kind == H_STRAIGHT_LINE) {
        h.G_params = stage_array(h.G_params, h.Ge.P > 0 ? h.Ge.P : 1, cur);
        h.F_params = stage_array(h.F_params, h.Fe.P > 0 ? h.Fe.P : 1, cur);
    } else {
        h.p = stage_array(h.p, h.P > 0 ? h.P : 1, cur);
        if (h.q) h.q = stage_array(h.q, h.P > 0 ? h.P : 1, cur);
    }
}

// Main loop of the thread-per-path engines.  Lanes park with an event (path finished, toric stage over, free)
// and the warp handles the parked lanes together once `refill_min` of them wait or nobody can step, so that
// the once-per-path work (init_newton!, first predictor update, condition number of the endpoint) runs with
// many lanes instead of stalling the warp once per lane.
template <class LaneT>
__device__ __forceinline__ void tpp_loop(LaneT& L, const KArgs& A, const KArgs& sA) {
    L.phase = PH_IDLE; L.ev = EV_START; L.stop_pending = false;
    const long long N = sA.B.N;
    const unsigned lane = threadIdx.x & 31u;
    bool drained = false;  // warp-uniform
    while (true) {
        const unsigned pend = __ballot_sync(0xffffffffu, L.ev != EV_NONE);
        const unsigned act = __ballot_sync(0xffffffffu, L.ev == EV_NONE && L.phase != PH_IDLE);
        // sync_cta: the warps of the CTA begin each round (one tracker step per lane) together, so that they walk the
        // same code at about the same time and share the instruction cache; a warp without work keeps arriving at the
        // barrier until the whole CTA is done
        if (A.sync_cta) { if (!__syncthreads_or((pend | act) != 0u)) break; }
        else if (pend == 0u && act == 0u) break;
        if (__popc(pend) >= (A.refill_min > 0 ? A.refill_min : 8) || act == 0u) {
            if (L.ev != EV_NONE) L.event_finish(sA.R);
            const unsigned want = __ballot_sync(0xffffffffu, L.ev == EV_START);
            long long k = -1;
            if (want != 0u && !drained) {
                if (A.cancel && *(const volatile int*)A.cancel) drained = true;  // stop_early_cb / interrupt: no new paths
            }
            if (want != 0u && !drained) {
                const int cnt = __popc(want);
                long long base = 0;
                if (lane == 0) base = (long long)atomicAdd(A.queue, (unsigned long long)cnt);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + cnt >= N) drained = true;
                if (L.ev == EV_START) k = base + __popc(want & ((1u << lane) - 1u));
            }
            if (L.ev != EV_NONE) L.event_begin(k, k >= 0 && k < N, sA.B, sA.R);
        }
        if (L.ev == EV_NONE && L.phase != PH_IDLE) L.iterate(sA.B, sA.R);
    }
}

// Lockstep variant (specialised kernels): all warps of the CTA go through a round together -- event handling in the
// same rounds, one tracker step with CTA-wide barriers inside (tracker_step_t<true>) -- so that they execute the same
// straight-line code at the same time and share what the SM's instruction cache holds.
template <class LaneT>
__device__ __forceinline__ void tpp_loop_sync(LaneT& L, const KArgs& A, const KArgs& sA) {
    L.phase = PH_IDLE; L.ev = EV_START; L.stop_pending = false;
    const long long N = sA.B.N;
    const unsigned lane = threadIdx.x & 31u;
    const int nwarps = (int)(blockDim.x >> 5);
    bool drained = false;  // warp-uniform
    // How many parked lanes make an event round worth it.  An event round costs E regular rounds whatever the number of
    // lanes in it; waiting for T lanes parks T / 2 of the B lanes on average.  With f events per round the overhead
    // T / (2 B) + E f / T is least at T = sqrt(2 B E f): 50 - 64 lanes for paths of 100 steps (cyclic-7, katsura: E = 1),
    // 130 - 200 for the 10-step paths of a parameter sweep (E = 4).  Thread 0 measures E (clock64 around the event part
    // and around the step of a round) and f (parked lanes per round since the last event round) and publishes the next
    // threshold through shared memory AFTER the step of the round -- every thread has passed a barrier of the step since
    // it read the old value, and reads the new one behind the barriers at the top of the next round, so the decision
    // stays CTA-uniform.  A.refill_min > 0 pins the threshold (lanes per warp); A.refill_k scales the model.
    __shared__ int s_thr;
    const int nthr0 = A.refill_min > 0 ? A.refill_min * nwarps : 8 * nwarps;
    if (threadIdx.x == 0) s_thr = nthr0;
    int since = 0, next_thr = nthr0;
    float t_reg = 0.f, t_ev = 0.f;
    while (true) {
        const int npend = __syncthreads_count(L.ev != EV_NONE);
        const int nact = __syncthreads_count(L.ev == EV_NONE && L.phase != PH_IDLE);
        if (npend == 0 && nact == 0) break;
        const int thr = s_thr;
        if (npend >= thr || nact == 0) {
            const long long c0 = clock64();
            if (L.ev != EV_NONE) L.event_finish(sA.R);
            const unsigned want = __ballot_sync(0xffffffffu, L.ev == EV_START);
            long long k = -1;
            if (want != 0u && !drained) {
                if (A.cancel && *(const volatile int*)A.cancel) drained = true;
            }
            if (want != 0u && !drained) {
                const int cnt = __popc(want);
                long long base = 0;
                if (lane == 0) base = (long long)atomicAdd(A.queue, (unsigned long long)cnt);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + cnt >= N) drained = true;
                if (L.ev == EV_START) k = base + __popc(want & ((1u << lane) - 1u));
            }
            if (L.ev != EV_NONE) L.event_begin(k, k >= 0 && k < N, sA.B, sA.R);
            if (threadIdx.x == 0 && A.refill_min <= 0) {
                const float dt = (float)(clock64() - c0);
                t_ev = t_ev > 0.f ? 0.75f * t_ev + 0.25f * dt : dt;
                if (nact != 0 && since > 0 && t_reg > 0.f) {
                    const float f = (float)npend / (float)since, E = t_ev / t_reg;
                    int t = (int)sqrtf(2.0f * A.refill_k * (float)blockDim.x * E * f);
                    const int lo = 4 * nwarps, hi = 24 * nwarps;
                    t = t < lo ? lo : (t > hi ? hi : t);
                    next_thr = (thr + t + 1) >> 1;
                }
            }
            since = 0;
        }
        const long long c2 = clock64();
        L.template iterate_t<true>(L.ev == EV_NONE && L.phase != PH_IDLE);
        if (threadIdx.x == 0 && A.refill_min <= 0) {
            const float dt = (float)(clock64() - c2);
            t_reg = t_reg > 0.f ? 0.9f * t_reg + 0.1f * dt : dt;
            s_thr = next_thr;
        }
        ++since;
    }
}
#endif  // __CUDACC__

#if !defined(__CUDACC_RTC__)
// Sequential driver of the host-compiled device code (tests/host_sim): one lane, path after path.
template <class LaneT>
inline void sim_loop(LaneT& L, const KArgs& A) {
    for (long long k = 0; k < A.B.N; ++k) {
        L.phase = PH_IDLE; L.ev = EV_START; L.stop_pending = false;
        L.event_begin(k, true, A.B, A.R);
        while (true) {
            if (L.ev != EV_NONE) {
                L.event_finish(A.R);
                if (L.ev == EV_START) break;  // path done
                L.event_begin(-1, false, A.B, A.R);
                continue;
            }
            L.iterate(A.B, A.R);
        }
    }
}
#endif

}  // namespace hc
