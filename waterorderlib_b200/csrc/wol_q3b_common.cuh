// Pieces shared by the evaluation kernels (group-per-centre generic path and thread-per-centre fast path).
#pragma once
#include <limits.h>

#include "wol_device.cuh"
#include "wol_internal.h"
#include "wol_workspace.h"

namespace wol {

template <typename T>
struct alignas(16) Vec4 {
    T x, y, z, w;
};

struct Q3bParams {
    const void *recs;
    const uint32_t *cell_start;
    const double *box;
    const void *centres;  // nullptr: every atom is a centre (visited in cell order)
    const int32_t *n_valid;  // optional [n_frames]: only the first n_valid[f] centres of frame f are evaluated
    int centre_dtype;
    int n_frames, n_pos, n_centres;
    int nc0, nc1, nc2;
    double low3sq, high3sq, lowqsq, highqsq;
    double highq;
    double rc1;   // radius inside which a half-width-1 stencil is complete
    int wq_max;   // half-width at which the q search is complete whatever it finds
    int do_q, do_3b;
    int nbins, q_nbins;
    double hist_lo, hist_hi;
    const double *table;
    void *q;
    int32_t *nn_idx;
    int32_t *n3;
    unsigned long long *ang_hist;
    unsigned long long *q_hist;
    double *stats;
    int hist_per_frame;
    uint32_t *counters;
    uint32_t *fb_list;
    int tiles_per_frame;
    int chunk_tiles;          // thread-per-centre path: tiles per round-robin chunk
    long long total_tiles;
    // thread-per-centre fast path
    const float4 *wrapped;  // box-wrapped float coordinates in record order (nullptr: not built)
    const uint32_t *cellpack;  // FP32 mode: packed cell coordinates in record order
    float pre_thr2;         // prefilter acceptance threshold on the float squared distance
    void *ev_begin, *ev_end;  // optional cudaEvent_t around the dominant kernel
    // which queue a large-capacity / widened launch walks
    const uint32_t *list;     // entries
    int list_counter;         // index into counters[] of its length
    int list_w_start;         // half-width at which a q-only entry resumes its search
    uint32_t widen_lo, widen_hi;  // a widened-search launch works only when widen_lo <= counters[kCntWidened] < widen_hi
    uint32_t *list2;          // second-level queue (appended to by the thread-per-centre widened pass)
    float lowq_hi2;           // (lowq + margin)^2: float distances above this are certainly beyond lowCut
    float pre_thr2_w2;        // prefilter threshold of the half-width-2 pass
    float pre_cst_w2;         // slack added to the 4th-smallest float distance^2 of that pass
    float cell_eps;           // how far a wrapped float coordinate can sit outside its cell's nominal box
};

// ------------------------------------------------------------------------------------------------

template <typename T>
struct RecTraits;
template <>
struct RecTraits<double> {
    typedef RecD Rec;
    static __device__ __forceinline__ void load(const void *recs, size_t j, double &x, double &y, double &z, int &idx) {
        // the 32-byte record is one sector: one 256-bit load (LDG.256) instead of two 128-bit ones -- the load / store
        // unit's transaction rate, not bytes, is what the sweeps over the records run into
        const RecD *p = reinterpret_cast<const RecD *>(recs) + j;
        long long a, b, c, d;
        asm("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        x = __longlong_as_double(a);
        y = __longlong_as_double(b);
        z = __longlong_as_double(c);
        idx = (int)d;
    }
    static __device__ __forceinline__ int cell(const void *recs, size_t j) {
        return reinterpret_cast<const RecD *>(recs)[j].cell;
    }
};
template <>
struct RecTraits<float> {
    typedef RecF Rec;
    static __device__ __forceinline__ void load(const void *recs, size_t j, float &x, float &y, float &z, int &idx) {
        const int4 a = __ldg(reinterpret_cast<const int4 *>(reinterpret_cast<const RecF *>(recs) + j));
        x = __int_as_float(a.x);
        y = __int_as_float(a.y);
        z = __int_as_float(a.z);
        idx = a.w;
    }
};

template <typename T>
__device__ __forceinline__ bool key_less(T d0, int i0, T d1, int i1) {
    return d0 < d1 || (d0 == d1 && i0 < i1);
}

// Per-lane sorted top-4 by (distance, atom index); payload = where the candidate can be found again.
template <typename T>
struct Top4 {
    T d[4];
    int i[4];
    int p[4];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            d[k] = Ops<T>::inf();
            i[k] = INT_MAX;
            p[k] = -1;
        }
    }
    __device__ __forceinline__ void insert(T dd, int ii, int pp) {
        if (key_less(dd, ii, d[3], i[3])) {
            d[3] = dd;
            i[3] = ii;
            p[3] = pp;
#pragma unroll
            for (int k = 3; k > 0; --k) {
                if (key_less(d[k], i[k], d[k - 1], i[k - 1])) {
                    const T td = d[k]; d[k] = d[k - 1]; d[k - 1] = td;
                    const int ti = i[k]; i[k] = i[k - 1]; i[k - 1] = ti;
                    const int tp = p[k]; p[k] = p[k - 1]; p[k - 1] = tp;
                }
            }
        }
    }
    __device__ __forceinline__ void pop() {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            d[k] = d[k + 1];
            i[k] = i[k + 1];
            p[k] = p[k + 1];
        }
        d[3] = Ops<T>::inf();
        i[3] = INT_MAX;
        p[3] = -1;
    }
};

__device__ __forceinline__ double shfl_xor_t(double v, int o, int w) { return __shfl_xor_sync(kFullMask, v, o, w); }
__device__ __forceinline__ float shfl_xor_t(float v, int o, int w) { return __shfl_xor_sync(kFullMask, v, o, w); }
__device__ __forceinline__ double shfl_t(double v, int src, int w) { return __shfl_sync(kFullMask, v, src, w); }
__device__ __forceinline__ float shfl_t(float v, int src, int w) { return __shfl_sync(kFullMask, v, src, w); }

// Position of the angle belonging to clamped cosine c on the histogram axis: -1 below the range,
// 0..nbins-1 a bin, nbins above the range.  tab[k] (decreasing in k) is the largest c whose angle sits
// at or beyond bin k, so the position is the largest k with c <= tab[k]; the float acos only seeds
// the search.
__device__ __forceinline__ int angle_position(double c, const double *tab, int nbins, float lo, float inv_width) {
    if (c == -1.0) return (int)tab[nbins + 1];
    // seed: acos(|x|) ~ sqrt(1 - |x|) * cubic(|x|) (Abramowitz & Stegun 4.4.45, |error| < 7e-5 rad = 0.004 degrees, a
    // hundredth of a default bin); the exact position comes from the table below whatever the seed is
    const float x = (float)c, ax = fabsf(x);
    float r = fmaf(fmaf(fmaf(-0.0187293f, ax, 0.0742610f), ax, -0.2121144f), ax, 1.5707288f) * sqrtf(fmaxf(1.0f - ax, 0.f));
    r = x < 0.f ? 3.14159265f - r : r;
    int k = (int)((r * 57.29577951f - lo) * inv_width);
    k = min(max(k, 0), nbins - 1);
    const double t0 = tab[k], t1 = tab[k + 1];  // both thresholds of the seeded bin at once: the usual case ends here
    if (c <= t0 && !(c <= t1)) return k;
    if (c <= t1) {
        ++k;
        while (k < nbins && c <= tab[k + 1]) ++k;
    } else {
        --k;
        while (k >= 0 && !(c <= tab[k])) --k;
    }
    return k;
}

__device__ __forceinline__ int angle_position(double c, const double *tab, int nbins, double lo, double inv_width) {
    return angle_position(c, tab, nbins, (float)lo, (float)inv_width);
}

struct LaneStats {
    double q_sum, q_sumsq, tet_cos, tet_cossq;
    unsigned n_centres, tet_count, n_angles, n_neigh;
    __device__ __forceinline__ void reset() {
        q_sum = q_sumsq = tet_cos = tet_cossq = 0.0;
        n_centres = tet_count = n_angles = n_neigh = 0u;
    }
};

// (cold: once per frame a block touches -- kept out of line so the hot loop stays compact in the instruction cache)
static __device__ __noinline__ void flush_stats(const Q3bParams &P, int f, LaneStats &st) {
    double v[8];
    v[WOL_STAT_Q_SUM] = st.q_sum;
    v[WOL_STAT_Q_SUMSQ] = st.q_sumsq;
    v[WOL_STAT_N_CENTRES] = (double)st.n_centres;
    v[WOL_STAT_TET_COUNT] = (double)st.tet_count;
    v[WOL_STAT_TET_COS] = st.tet_cos;
    v[WOL_STAT_TET_COSSQ] = st.tet_cossq;
    v[WOL_STAT_N_ANGLES] = (double)st.n_angles;
    v[WOL_STAT_N_NEIGH] = (double)st.n_neigh;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double s = warp_sum(v[k]);
        if (P.stats && (threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(P.stats + (size_t)f * WOL_NSTATS + k, s);
    }
    st.reset();
}

template <typename T>
__device__ __forceinline__ void load_centre(const Q3bParams &P, int f, int m, T &rx, T &ry, T &rz) {
    const size_t o = ((size_t)f * P.n_centres + m) * 3;
    if (P.centre_dtype == WOL_F64) {
        const double *c = reinterpret_cast<const double *>(P.centres);
        rx = (T)c[o]; ry = (T)c[o + 1]; rz = (T)c[o + 2];
    } else {
        const float *c = reinterpret_cast<const float *>(P.centres);
        rx = (T)c[o]; ry = (T)c[o + 1]; rz = (T)c[o + 2];
    }
}


constexpr int kMaxSmemBins = 4096;

// q of one centre from its (up to) four selected neighbours, given as record indices in selection
// order; writes q / nn_idx / q histogram and accumulates the frame statistics.  fp64 reference
// arithmetic: reimage (waterlib.f90:43-45), the second reimage tetraCosAng applies (:880-883),
// CosAngle3's vectors and clamped cosine (:694-698), padding and sum of water_properties.py:379-388.
template <bool EXACT>
__device__ __forceinline__ void finish_q(const Q3bParams &P, int f, double rx, double ry, double rz, double Lx, double Ly,
                                         double Lz, double iLx, double iLy, double iLz, const Top4<double> &top,
                                         int n_found, size_t out_index, LaneStats &st, unsigned *s_qhist) {
    double vx[4], vy[4], vz[4], vn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        vx[k] = vy[k] = vz[k] = vn[k] = 0.0;
        if (k < n_found) {
            double px, py, pz;
            int idx;
            RecTraits<double>::load(P.recs, (size_t)top.p[k], px, py, pz, idx);
            const double dx = min_image_1<double, EXACT>(px, rx, Lx, iLx);
            const double dy = min_image_1<double, EXACT>(py, ry, Ly, iLy);
            const double dz = min_image_1<double, EXACT>(pz, rz, Lz, iLz);
            const double ex = __dsub_rn(__dadd_rn(rx, dx), rx);
            const double ey = __dsub_rn(__dadd_rn(ry, dy), ry);
            const double ez = __dsub_rn(__dadd_rn(rz, dz), rz);
            const double d2x = __dsub_rn(ex, __dmul_rn(Lx, anint_exact<double>(__dmul_rn(ex, iLx))));
            const double d2y = __dsub_rn(ey, __dmul_rn(Ly, anint_exact<double>(__dmul_rn(ey, iLy))));
            const double d2z = __dsub_rn(ez, __dmul_rn(Lz, anint_exact<double>(__dmul_rn(ez, iLz))));
            vx[k] = __dsub_rn(__dadd_rn(rx, d2x), rx);
            vy[k] = __dsub_rn(__dadd_rn(ry, d2y), ry);
            vz[k] = __dsub_rn(__dadd_rn(rz, d2z), rz);
            vn[k] = sumsq3<double>(vx[k], vy[k], vz[k]);
        }
    }
    // real angles in triu order, then the 180-degree padding (cos = -1), summed left to right like np.sum
    double acc = 0.0;
    int n_real = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b)
            if (b < n_found) {
                double c;
                if (vn[a] == 0.0 || vn[b] == 0.0) c = 1.0;
                else c = clamped_cos<double>(dot3<double>(vx[a], vy[a], vz[a], vx[b], vy[b], vz[b]), vn[a], vn[b]);
                const double u = c + (1.0 / 3.0);
                acc += u * u;
                ++n_real;
            }
    for (int k = n_real; k < 6; ++k) {
        const double u = -1.0 + (1.0 / 3.0);
        acc += u * u;
    }
    const double qv = (n_found == 0) ? 0.0 : 1.0 - (3.0 / 8.0) * acc;
    if (P.q) reinterpret_cast<double *>(P.q)[out_index] = qv;
    if (P.nn_idx) {
        int4 o;
        o.x = (n_found > 0) ? top.i[0] : -1;
        o.y = (n_found > 1) ? top.i[1] : -1;
        o.z = (n_found > 2) ? top.i[2] : -1;
        o.w = (n_found > 3) ? top.i[3] : -1;
        reinterpret_cast<int4 *>(P.nn_idx)[out_index] = o;
    }
    if (P.q_hist) {
        const HistSpec hs = hist_spec(0.0, 1.0, P.q_nbins);
        const int b = hist_bin(hs, qv);
        if (b >= 0) {
            if (s_qhist) atomicAdd(s_qhist + b, 1u);
            else atomicAdd(P.q_hist + (size_t)(P.hist_per_frame ? f : 0) * P.q_nbins + b, 1ull);
        }
    }
    st.q_sum += qv;
    st.q_sumsq += qv * qv;
    st.n_centres += 1u;
}

// flush of a block's shared-memory histograms into the global int64 bins
__device__ __forceinline__ void flush_bins(unsigned *s_bins, unsigned long long *g_bins, int nbins, bool clear) {
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
        const unsigned v = s_bins[i];
        if (v) atomicAdd(g_bins + i, (unsigned long long)v);
        if (clear) s_bins[i] = 0u;
    }
}
__device__ __forceinline__ void flush_hist(const Q3bParams &P, unsigned *s_hist, unsigned *s_qhist, int f, bool clear) {
    const size_t row = (size_t)(P.hist_per_frame ? f : 0);
    if (s_hist) flush_bins(s_hist, P.ang_hist + row * P.nbins, P.nbins, clear);
    if (s_qhist) flush_bins(s_qhist, P.q_hist + row * P.q_nbins, P.q_nbins, clear);
}

int q3b_tpc_launch(const Q3bParams &P, cudaStream_t stream, bool exact);
int q3b_tpc32_launch(const Q3bParams &P, cudaStream_t stream);
int q3b_tpc_widen_launch(const Q3bParams &P, cudaStream_t stream, bool exact, bool f32);
bool q3b_tpc_widen_supported(const Q3bParams &P);
bool q3b_tpc_supported(const Q3bParams &P);
bool q3b_brick_supported(const Q3bParams &P, bool exact);
int q3b_brick_launch(const Q3bParams &P, double box_max, cudaStream_t stream);
bool q3b_brick_ws_supported(const Q3bParams &P, bool exact);
bool q3b_brick32_supported(const Q3bParams &P);
int q3b_brick32_launch(const Q3bParams &P, cudaStream_t stream);
int q3b_brick_ws_launch(const Q3bParams &P, double box_max, cudaStream_t stream);

}  // namespace wol
