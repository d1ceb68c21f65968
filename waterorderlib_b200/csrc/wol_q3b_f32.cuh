// Float pieces shared by the FP32-mode kernels (thread-per-centre sweep and widened search).
#pragma once
#include "wol_q3b_common.cuh"

namespace wol {

struct Top4F {
    float d[4];
    int i[4];
    float x[4], y[4], z[4];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            d[k] = __int_as_float(0x7f800000);
            i[k] = INT_MAX;
            x[k] = y[k] = z[k] = 0.f;
        }
    }
    __device__ __forceinline__ void insert(float dd, int ii, float xx, float yy, float zz) {
        if (key_less(dd, ii, d[3], i[3])) {
            d[3] = dd; i[3] = ii; x[3] = xx; y[3] = yy; z[3] = zz;
#pragma unroll
            for (int k = 3; k > 0; --k) {
                if (key_less(d[k], i[k], d[k - 1], i[k - 1])) {
                    float t;
                    t = d[k]; d[k] = d[k - 1]; d[k - 1] = t;
                    t = x[k]; x[k] = x[k - 1]; x[k - 1] = t;
                    t = y[k]; y[k] = y[k - 1]; y[k - 1] = t;
                    t = z[k]; z[k] = z[k - 1]; z[k - 1] = t;
                    const int ti = i[k]; i[k] = i[k - 1]; i[k - 1] = ti;
                }
            }
        }
    }
};

// clamped cosine between two difference vectors with squared norms na, nb
__device__ __forceinline__ float cos32(float ax, float ay, float az, float na, float bx, float by, float bz, float nb) {
    const float dot = fmaf(az, bz, fmaf(ay, by, ax * bx));
    return fminf(1.f, fmaxf(-1.f, dot * rsqrtf(na * nb)));
}

// q of one centre from its (up to) four selected neighbours in FP32 mode: pair cosines in triu order, the
// 180-degree padding (cos = -1) for centres with fewer than four neighbours, q = 1 - 3/8 sum (cos + 1/3)^2
// (structureLibs/water_properties.py:379-388); writes q / nn_idx / q histogram, accumulates the statistics.
__device__ __forceinline__ void finish_q32(const Q3bParams &P, int f, const Top4F &top, int n_found, size_t out_index,
                                           LaneStats &st, unsigned *s_qhist, const HistSpec &qspec) {
    float acc = 0.f;
    int n_real = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b)
            if (b < n_found) {
                const float c = (top.d[a] == 0.f || top.d[b] == 0.f)
                                    ? 1.f
                                    : cos32(top.x[a], top.y[a], top.z[a], top.d[a], top.x[b], top.y[b], top.z[b], top.d[b]);
                const float u = c + (1.0f / 3.0f);
                acc += u * u;
                ++n_real;
            }
    acc += (float)(6 - n_real) * ((-1.0f + 1.0f / 3.0f) * (-1.0f + 1.0f / 3.0f));
    const float qv = (n_found == 0) ? 0.f : 1.0f - 0.375f * acc;
    if (P.q) reinterpret_cast<float *>(P.q)[out_index] = qv;
    if (P.nn_idx) {
        int4 o;
        o.x = (n_found > 0) ? top.i[0] : -1;
        o.y = (n_found > 1) ? top.i[1] : -1;
        o.z = (n_found > 2) ? top.i[2] : -1;
        o.w = (n_found > 3) ? top.i[3] : -1;
        reinterpret_cast<int4 *>(P.nn_idx)[out_index] = o;
    }
    if (P.q_hist) {
        const int b = hist_bin(qspec, (double)qv);
        if (b >= 0) {
            if (s_qhist) atomicAdd(s_qhist + b, 1u);
            else atomicAdd(P.q_hist + (size_t)(P.hist_per_frame ? f : 0) * P.q_nbins + b, 1ull);
        }
    }
    st.q_sum += (double)qv;
    st.q_sumsq += (double)qv * (double)qv;
    st.n_centres += 1u;
}

}  // namespace wol
