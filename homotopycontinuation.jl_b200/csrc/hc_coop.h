// Cooperative lane group: G lanes of one warp (G = 1, 2, 4, 8, 16, 32) track ONE solution path
// together.  The path's vectors and matrices live in the CTA's shared memory; every lane of the
// group executes the same control flow on replicated scalars (so there is no divergence inside a
// group, and with G = 32 none inside a warp), vector work is strided over the lanes, reductions are
// xor-butterflies whose result is bit-identical in all lanes (commutative, NaN-symmetric) -- which
// is what keeps the replicated control flow uniform.
//
// With HC_HOST_SIM (g++ build for CPU-side unit tests of the kernel logic) only G = 1 exists.
#pragma once
#include "hc_common.h"

namespace hc {

template <int G>
struct Grp {
    static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group size");
#if defined(__CUDA_ARCH__)
    unsigned mask;
    int lane;
    HC_D void init() {
        const unsigned wl = threadIdx.x & 31u;
        lane = (int)(wl & (unsigned)(G - 1));
        mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(unsigned)(G - 1)));
    }
    HC_D void sync() const { if (G > 1) __syncwarp(mask); }
    HC_D double xshfl(double v, int o) const { return __shfl_xor_sync(mask, v, o, G); }
    HC_D int xshfl(int v, int o) const { return __shfl_xor_sync(mask, v, o, G); }
    HC_D int bcast(int v, int src) const { return G > 1 ? __shfl_sync(mask, v, src, G) : v; }
    HC_D long long bcast(long long v, int src) const { return G > 1 ? __shfl_sync(mask, v, src, G) : v; }
    HC_D double bcast(double v, int src) const { return G > 1 ? __shfl_sync(mask, v, src, G) : v; }
#else
    static constexpr int lane = 0;
    void init() {}
    void sync() const {}
    double xshfl(double v, int) const { return v; }
    int xshfl(int v, int) const { return v; }
    int bcast(int v, int) const { return v; }
    long long bcast(long long v, int) const { return v; }
    double bcast(double v, int) const { return v; }
#endif
    // max over the group; a NaN in any lane gives NaN in all lanes
    HC_HD double rmax(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            double w = xshfl(v, o);
            v = (v != v || w != w) ? HC_NAN : (w > v ? w : v);
        }
        return v;
    }
    HC_HD double rmin(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            double w = xshfl(v, o);
            v = (v != v || w != w) ? HC_NAN : (w < v ? w : v);
        }
        return v;
    }
    HC_HD double rsum(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = v + xshfl(v, o);
        return v;
    }
    HC_HD bool rall(bool p) const {
        int v = p ? 1 : 0;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v &= xshfl(v, o);
        return v != 0;
    }
    HC_HD bool rany(bool p) const {
        int v = p ? 1 : 0;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v |= xshfl(v, o);
        return v != 0;
    }
    // arg max: larger value wins, ties go to the smaller index (values must not be NaN)
    HC_HD void argmax(double& v, int& idx) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            double w = xshfl(v, o);
            int j = xshfl(idx, o);
            if (w > v || (w == v && j < idx)) { v = w; idx = j; }
        }
    }
    // arg min with the same tie rule
    HC_HD void argmin(double& v, int& idx) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            double w = xshfl(v, o);
            int j = xshfl(idx, o);
            if (w < v || (w == v && j < idx)) { v = w; idx = j; }
        }
    }
};

}  // namespace hc
