// Device-side helpers shared by the sm_100a kernels: reference-order fp64 arithmetic (no FMA
// contraction), the angle formula, numpy's uniform-bin histogram rule, fixed-point periodic
// coordinates and small warp utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wol {

// The Fortran parameters (fortran/waterlib.f90:685-686): pi = 3.1415926535897931D0
__device__ constexpr double kPi = 3.1415926535897931;
__device__ constexpr double kTwoPi = 3.1415926535897931 * 2.0;
__device__ constexpr double kDegPerRad = 180.0 / 3.1415926535897931;

struct Box {
    double L[3];
    double iL[3];
};

// iBoxL = merge(1.d0/BoxL, 0.d0, BoxL >= 0.d0)   (fortran/waterlib.f90:41)
__device__ __forceinline__ void box_load(Box &b, const double *__restrict__ box3) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double L = box3[k];
        b.L[k] = L;
        b.iL[k] = (L >= 0.0) ? __ddiv_rn(1.0, L) : 0.0;
    }
}

// distvec = p - r ; distvec = distvec - BoxL * anint(distvec * iBoxL)  (fortran/waterlib.f90:43-44)
// Each operation is individually rounded, as in the reference's SSE2 build; `round` is
// round-half-away-from-zero like Fortran's anint.
__device__ __forceinline__ double min_image_1(double p, double r, double L, double iL) {
    double t = __dsub_rn(p, r);
    double s = __dmul_rn(t, iL);
    return __dsub_rn(t, __dmul_rn(L, round(s)));
}

// sum(v**2) in the order x, y, z
__device__ __forceinline__ double sumsq3(double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
}

// Clamped cosine of CosAngle3 (fortran/waterlib.f90:696-698): min(1, max(-1, dot / sqrt(n1*n2)))
__device__ __forceinline__ double clamped_cos(double dot, double n1, double n2) {
    double norm = __dsqrt_rn(__dmul_rn(n1, n2));
    double c = __ddiv_rn(dot, norm);
    return fmin(1.0, fmax(-1.0, c));
}

// Angle in degrees from the clamped cosine (fortran/waterlib.f90:699-702):
//   Phi = acos(c); A = mod(Phi + pi, 2 pi) - pi; if (A < -pi) A += 2 pi; A * DegPerRad
// Phi + pi < 2 pi for every Phi < pi, so the mod only acts when Phi == pi exactly (c == -1), where it
// yields 0 and the angle becomes -180.  The add/subtract of pi is kept
// because it rounds Phi onto a coarser grid and the reference's bits depend on it.
__device__ __forceinline__ double angle_deg_from_cos(double c) {
    double phi = acos(c);
    double a = __dadd_rn(phi, kPi);
    if (a >= kTwoPi) a = fmod(a, kTwoPi);
    a = __dsub_rn(a, kPi);
    return __dmul_rn(a, kDegPerRad);
}

// np.histogram(x, bins=nbins, range=[lo, hi]) bin of one value, -1 if outside (numpy
// lib/_histograms_impl.py, uniform-bin fast path; edges = linspace(lo, hi, nbins + 1)).
struct HistSpec {
    double lo, hi, denom, step;
    int nbins;
};

__device__ __forceinline__ HistSpec hist_spec(double lo, double hi, int nbins) {
    HistSpec h;
    h.lo = lo;
    h.hi = hi;
    h.nbins = nbins;
    h.denom = __dsub_rn(hi, lo);
    h.step = __ddiv_rn(h.denom, (double)nbins);
    return h;
}

__device__ __forceinline__ double hist_edge(const HistSpec &h, int k) {
    return (k == h.nbins) ? h.hi : __dadd_rn(__dmul_rn((double)k, h.step), h.lo);
}

__device__ __forceinline__ int hist_bin(const HistSpec &h, double x) {
    if (!(x >= h.lo) || !(x <= h.hi)) return -1;
    double f = __dmul_rn(__ddiv_rn(__dsub_rn(x, h.lo), h.denom), (double)h.nbins);
    int idx = (int)f;  // truncation, f >= 0
    if (idx == h.nbins) idx -= 1;
    if (x < hist_edge(h, idx)) {
        idx -= 1;
    } else if (idx != h.nbins - 1 && x >= hist_edge(h, idx + 1)) {
        idx += 1;
    }
    return idx;
}

// Periodic fixed-point coordinate: frac(x / L) scaled to 2^32.  Differences of two such values wrap
// in two's complement, which IS the minimum image; resolution L / 2^32 (7e-8 A at L = 310 A).
__device__ __forceinline__ uint32_t to_fixed(double x, double iL) {
    double t = x * iL;
    t -= floor(t);
    // t in [0,1]; the product is < 2^32 + 1, the cast wraps a value of exactly 2^32 to 0
    unsigned long long u = __double2ull_rd(t * 4294967296.0);
    return (uint32_t)u;
}

__device__ __forceinline__ int cell_coord(uint32_t xf, int nc) {
    return (int)__umulhi(xf, (uint32_t)nc);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

}  // namespace wol
