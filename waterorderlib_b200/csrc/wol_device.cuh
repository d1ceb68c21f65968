// Device-side helpers shared by the sm_100a kernels: reference-order arithmetic (every fp64 operation
// individually rounded, no FMA contraction -- the reference's prebuilt Fortran is SSE2 code), the
// minimum-image idiom, numpy's uniform-bin histogram rule and small warp utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wol {

constexpr unsigned kFullMask = 0xffffffffu;

// Arithmetic policy.  Ops<double> is the parity mode: each operation is one IEEE round-to-nearest
// operation, in the order the reference performs it.  Ops<float> is the fast mode: plain float
// arithmetic, contraction allowed.
template <typename T>
struct Ops;

template <>
struct Ops<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    // round-to-nearest-even integer via the 1.5 * 2^52 trick (|s| < 2^51)
    static __device__ __forceinline__ double rint_fast(double s) {
        return __dsub_rn(__dadd_rn(s, 6755399441055744.0), 6755399441055744.0);
    }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};

template <>
struct Ops<float> {
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float rint_fast(float s) { return rintf(s); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};

// Fortran anint: round half away from zero (fortran/waterlib.f90:44).  s - trunc(s) is exact.
template <typename T>
__device__ __forceinline__ T anint_exact(T s) {
    T n = (sizeof(T) == 8) ? (T)trunc((double)s) : (T)truncf((float)s);
    T f = s - n;
    if (f >= (T)0.5) n += (T)1;
    if (f <= (T)-0.5) n -= (T)1;
    return n;
}

// distvec = p - r ; distvec = distvec - BoxL * anint(distvec * iBoxL)   (fortran/waterlib.f90:43-44)
// EXACT = false uses round-half-even: it differs from anint only when |distvec| is L/2 to the last
// bit, which cannot happen for a pair that passes a cutoff below 0.49 L; the host picks EXACT = true
// whenever a cutoff reaches that far.
template <typename T, bool EXACT>
__device__ __forceinline__ T min_image_1(T p, T r, T L, T iL) {
    const T t = Ops<T>::sub(p, r);
    const T s = Ops<T>::mul(t, iL);
    const T n = EXACT ? anint_exact<T>(s) : Ops<T>::rint_fast(s);
    return Ops<T>::sub(t, Ops<T>::mul(L, n));
}

// sum(v**2) / dot_product in the order x, y, z
template <typename T>
__device__ __forceinline__ T sumsq3(T x, T y, T z) {
    return Ops<T>::add(Ops<T>::add(Ops<T>::mul(x, x), Ops<T>::mul(y, y)), Ops<T>::mul(z, z));
}

template <typename T>
__device__ __forceinline__ T dot3(T ax, T ay, T az, T bx, T by, T bz) {
    return Ops<T>::add(Ops<T>::add(Ops<T>::mul(ax, bx), Ops<T>::mul(ay, by)), Ops<T>::mul(az, bz));
}

// Clamped cosine of CosAngle3 (fortran/waterlib.f90:696-698): min(1, max(-1, dot / sqrt(n1*n2)))
template <typename T>
__device__ __forceinline__ T clamped_cos(T dot, T n1, T n2) {
    const T norm = Ops<T>::sqrt(Ops<T>::mul(n1, n2));
    const T c = Ops<T>::div(dot, norm);
    return (T)fmin((double)1.0, fmax((double)-1.0, (double)c));
}

// Angle in degrees from the clamped cosine (fortran/waterlib.f90:699-702), for the paths that must
// return angle VALUES (the histogram path never calls this, see wol_angle_table):
//   Phi = acos(c); A = mod(Phi + pi, 2 pi) - pi; if (A < -pi) A += 2 pi; A * DegPerRad
// Phi + pi < 2 pi for every Phi < pi, so the mod only acts when Phi == pi exactly (c == -1), where it
// yields 0 and the angle becomes -180.
__device__ __forceinline__ double angle_deg_from_cos(double c) {
    const double kPi = 3.1415926535897931;
    const double kTwoPi = 3.1415926535897931 * 2.0;
    const double kDegPerRad = 180.0 / 3.1415926535897931;
    const double phi = acos(c);
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

__host__ __device__ inline HistSpec hist_spec(double lo, double hi, int nbins) {
    HistSpec h;
    h.lo = lo;
    h.hi = hi;
    h.nbins = nbins;
    h.denom = hi - lo;
    h.step = h.denom / (double)nbins;
    return h;
}

__device__ __forceinline__ double hist_edge(const HistSpec &h, int k) {
    return (k == h.nbins) ? h.hi : __dadd_rn(__dmul_rn((double)k, h.step), h.lo);
}

__device__ __forceinline__ int hist_bin(const HistSpec &h, double x) {
    if (!(x >= h.lo) || !(x <= h.hi)) return -1;
    const double f = __dmul_rn(__ddiv_rn(__dsub_rn(x, h.lo), h.denom), (double)h.nbins);
    int idx = (int)f;  // truncation, f >= 0
    if (idx == h.nbins) idx -= 1;
    if (x < hist_edge(h, idx)) {
        idx -= 1;
    } else if (idx != h.nbins - 1 && x >= hist_edge(h, idx + 1)) {
        idx += 1;
    }
    return idx;
}

// Cell coordinate of a position along one axis: frac(x / L) * nc, clamped.  Used identically by the
// build and by the sweep (for centres that are not atoms), so both agree on every cell.
__device__ __forceinline__ int cell_coord(double x, double iL, int nc) {
    double f = __dmul_rn(x, iL);
    f = __dsub_rn(f, floor(f));
    int c = (int)__dmul_rn(f, (double)nc);
    return min(max(c, 0), nc - 1);
}

// Box-wrapped coordinate frac(x / L) * L as a float, for the prefilter of the sweep.  Same frac as
// cell_coord, so an atom's wrapped coordinate lies inside its cell (up to float rounding).
__device__ __forceinline__ float wrapped_coord(double x, double L, double iL) {
    double f = __dmul_rn(x, iL);
    f = __dsub_rn(f, floor(f));
    return (float)__dmul_rn(f, L);
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(kFullMask, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

}  // namespace wol
