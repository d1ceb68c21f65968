// Same-sweep observables over the cell list, sm_100a, fp64 in the reference's
// operation order:
//   pair_hist_kernel   RadialDist / RadialDistSame / PairDistanceHistogram   fortran/waterlib.f90:193-231, :316-353, :358-389
//   psi_kernel         getOrderParamPsi                                         structureLibs/water_properties.py:393-433
#include <math.h>

#include "wol_q3b_common.cuh"

namespace wol {

struct PBox {
    double L[3], iL[3];
    double near;  // 0.49 x the smallest edge when all three are periodic, else 0 (pmin_image_3's short cut never taken)
};
__device__ __forceinline__ PBox load_pbox(const double *b) {
    PBox o;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.L[k] = b[k];
        o.iL[k] = (b[k] >= 0.0) ? __ddiv_rn(1.0, b[k]) : 0.0;
    }
    o.near = (b[0] > 0.0 && b[1] > 0.0 && b[2] > 0.0) ? 0.49 * fmin(b[0], fmin(b[1], b[2])) : 0.0;
    return o;
}

// distVec = p - r, minimum image (waterlib.f90:43-44) on the three axes with the common case short-cut: when every
// |p - r| is below 0.49 of the smallest edge, anint(distVec * iBoxL) is 0 on every axis and distVec - BoxL * 0 is
// distVec bit for bit, so the two products and the rounding are skipped (11 of the 12 fp64 instructions of an axis).
__device__ __forceinline__ void pmin_image_3(double px, double py, double pz, double rx, double ry, double rz, const PBox &b,
                                             double &dx, double &dy, double &dz) {
    dx = __dsub_rn(px, rx);
    dy = __dsub_rn(py, ry);
    dz = __dsub_rn(pz, rz);
    if (!(fabs(dx) < b.near && fabs(dy) < b.near && fabs(dz) < b.near)) {
        dx = __dsub_rn(dx, __dmul_rn(b.L[0], anint_exact<double>(__dmul_rn(dx, b.iL[0]))));
        dy = __dsub_rn(dy, __dmul_rn(b.L[1], anint_exact<double>(__dmul_rn(dy, b.iL[1]))));
        dz = __dsub_rn(dz, __dmul_rn(b.L[2], anint_exact<double>(__dmul_rn(dz, b.iL[2]))));
    }
}

struct PGrid {
    const uint32_t *cell_start;
    const void *recs;
    const float4 *wrapped;  // box-wrapped float coordinates in record order (float prefilter)
    int nc0, nc1, nc2;
};

// Float prefilter shared by the sweeps below: true when the record's float minimum-image distance^2 from the wrapped
// point w0 is within thr2 (a bound widened so that nothing inside the exact range can be rejected; everything that
// decides a result is recomputed exactly afterwards).
struct PFloat {
    float wx, wy, wz, Lx, Ly, Lz, iLx, iLy, iLz, thr2;
};
__device__ __forceinline__ PFloat make_pfloat(double x, double y, double z, const PBox &b, double reach) {
    PFloat p;
    p.wx = wrapped_coord(x, b.L[0], b.iL[0]);
    p.wy = wrapped_coord(y, b.L[1], b.iL[1]);
    p.wz = wrapped_coord(z, b.L[2], b.iL[2]);
    p.Lx = (float)b.L[0]; p.Ly = (float)b.L[1]; p.Lz = (float)b.L[2];
    p.iLx = 1.0f / p.Lx; p.iLy = 1.0f / p.Ly; p.iLz = 1.0f / p.Lz;
    const double m = reach + 16.0 * 5.9604644775390625e-8 * fmax(b.L[0], fmax(b.L[1], b.L[2]));
    p.thr2 = __double2float_ru(m * m * (1.0 + 1e-6));
    return p;
}
__device__ __forceinline__ bool pfloat_near(const PFloat &p, const float4 w) {
    float dx = w.x - p.wx, dy = w.y - p.wy, dz = w.z - p.wz;
    dx -= p.Lx * rintf(dx * p.iLx);
    dy -= p.Ly * rintf(dy * p.iLy);
    dz -= p.Lz * rintf(dz * p.iLz);
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= p.thr2;
}

template <typename T>
__device__ __forceinline__ void pload3(const void *p, int dtype, size_t i, T &x, T &y, T &z) {
    if (dtype == WOL_F64) {
        const double *d = reinterpret_cast<const double *>(p) + 3 * i;
        x = (T)d[0]; y = (T)d[1]; z = (T)d[2];
    } else {
        const float *d = reinterpret_cast<const float *>(p) + 3 * i;
        x = (T)d[0]; y = (T)d[1]; z = (T)d[2];
    }
}

// ---- pair-distance histograms --------------------------------------------------------------------------
// One thread per OUTER atom; it sweeps the 27-cell stencil (cell edge >= totbins * binwidth; every cell when an
// axis has <= 3) of the cell list built over the INNER set, bins nbin = ceiling(dist / binwidth) exactly as the
// Fortran does, and counts into a block-shared histogram flushed once.
//   mode 0 RadialDist            outer = Pos2, inner = Pos1, every pair
//   mode 1 RadialDistSame        outer = inner = Pos, pairs with inner index > outer index (the i < j loops)
//   mode 2 PairDistanceHistogram outer = Pos1, inner = Pos2, dist == 0 skipped

struct PairHistParams {
    PGrid grid;
    const double *box;
    const void *outer;
    int outer_dtype;
    int n_outer, n_inner;
    int mode;
    double binwidth;
    int totbins;
    double far_sq;  // distances^2 above this are beyond the last bin whatever the roundings: (totbins binwidth)^2 (1 + 1e-9)
    unsigned long long *counts;
};

constexpr int kPhThreads = 128;
constexpr int kPhQueue = 512;  // (owner lane, record) pairs a warp collects before it evaluates them
constexpr int kPhUnroll = 4;

// Three ideas, in the order they were measured (1M waters, 150 bins of 0.1 A: 12.1 ms at the start of this file's history):
//   * Cell order (mode 1): thread g takes the g-th atom of the cell-sorted list, so a warp sweeps the same cells and its
//     loads are broadcasts; a pair is counted once by its place in that list.
//   * A float pass decides which records can be in range at all (a sixth of the stencil); the periodic image of a whole
//     cell is known from its adjacency when every axis has more than 3 cells, so a candidate costs a 16-byte load, 3 FADD,
//     FMUL, 2 FFMA and a compare.
//   * What passes goes to a queue of the WARP (owner lane, record), appended with a ballot, and is evaluated by all 32 lanes
//     together when the queue fills: an atom near a face of its cell takes most of its pairs from the cell behind that face,
//     so per-lane lists fill at very different rates and a lane-owned exact pass runs at a third of its lanes.  Only the
//     fp64 value -- the reference's operations in the reference's order -- decides a bin.
//   * Bin edges in distance^2: the Fortran's bin, nb(s) = ceiling(RN(RN(sqrt(s)) / binwidth)), is a monotone step function
//     of s = distance^2, so with s_edge[k] = the LARGEST double s with nb(s) <= k it is settled by comparisons: bin k <=>
//     s_edge[k - 1] < s <= s_edge[k], no sqrt and no division per pair.  Every block finds the edges itself, a thread per
//     edge: from the estimate (k binwidth)^2, a few ulps away, it walks neighbouring doubles with the very operations it
//     replaces.  If a walk does not end within 64 steps the block keeps the per-pair arithmetic.
__global__ void __launch_bounds__(kPhThreads, 5) pair_hist_kernel(const PairHistParams P) {
    extern __shared__ unsigned s_cnt[];
    __shared__ uint2 s_queue[kPhThreads / 32][kPhQueue];
    __shared__ double s_ctr[3][kPhThreads];
    __shared__ int s_tab_bad;
    const bool smem = P.totbins <= kMaxSmemBins;
    double *const s_edge = reinterpret_cast<double *>(s_cnt + ((P.totbins + 1) & ~1));
    if (smem) {
        if (threadIdx.x == 0) s_tab_bad = 0;
        for (int i = threadIdx.x; i < P.totbins; i += blockDim.x) s_cnt[i] = 0u;
        __syncthreads();
        for (int k = threadIdx.x; k <= P.totbins; k += blockDim.x) {
            double t = 0.0;  // nb(0) = 0, nb(s) >= 1 for every s > 0
            if (k > 0) {
                const double e = __dmul_rn((double)k, P.binwidth);
                t = __dmul_rn(e, e);
                int steps = 0;
                while (ceil(__ddiv_rn(__dsqrt_rn(t), P.binwidth)) <= (double)k && steps < 64) {
                    t = __longlong_as_double(__double_as_longlong(t) + 1);
                    ++steps;
                }
                if (steps >= 64) s_tab_bad = 1;
                steps = 0;
                while (ceil(__ddiv_rn(__dsqrt_rn(t), P.binwidth)) > (double)k && steps < 64) {
                    t = __longlong_as_double(__double_as_longlong(t) - 1);
                    ++steps;
                }
                if (steps >= 64) s_tab_bad = 1;
            }
            s_edge[k] = t;
        }
        __syncthreads();
    }
    const bool use_edges = smem && s_tab_bad == 0;
    const float inv_bw_f = (float)(1.0 / P.binwidth);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lanes_below = (1u << lane) - 1u;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = g < P.n_outer;  // (every thread runs the loops: they hold warp-wide votes)
    const PBox b = load_pbox(P.box);
    double rx = 0.0, ry = 0.0, rz = 0.0;
    float wx = 0.f, wy = 0.f, wz = 0.f;
    const int nc0 = P.grid.nc0, nc1 = P.grid.nc1, nc2 = P.grid.nc2;
    int cx = 0, cy = 0, cz = 0;
    if (live) {
        if (P.mode == 1) {
            // outer = inner: the g-th atom of the cell-sorted list; dist(i, j) = dist(j, i) bit for bit (p - r and r - p are
            // exact negations, anint is odd), so counting a pair by list position (j > g) gives the i < j loops' counts
            const RecD *rp = reinterpret_cast<const RecD *>(P.grid.recs) + g;
            long long qa, qb, qc, qd;
            asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(qa), "=l"(qb), "=l"(qc), "=l"(qd) : "l"(rp));
            rx = __longlong_as_double(qa);
            ry = __longlong_as_double(qb);
            rz = __longlong_as_double(qc);
            const int cell = (int)(qd >> 32);
            cx = cell & 1023;
            cy = (cell >> 10) & 1023;
            cz = (cell >> 20) & 1023;
        } else {
            pload3<double>(P.outer, P.outer_dtype, (size_t)g, rx, ry, rz);
            cx = cell_coord(rx, b.iL[0], nc0);
            cy = cell_coord(ry, b.iL[1], nc1);
            cz = cell_coord(rz, b.iL[2], nc2);
        }
        wx = wrapped_coord(rx, b.L[0], b.iL[0]);
        wy = wrapped_coord(ry, b.L[1], b.iL[1]);
        wz = wrapped_coord(rz, b.L[2], b.iL[2]);
    }
    s_ctr[0][threadIdx.x] = rx;
    s_ctr[1][threadIdx.x] = ry;
    s_ctr[2][threadIdx.x] = rz;
    __syncwarp();
    const int j_min = (P.mode == 1) ? g + 1 : 0;
    const int j_last = P.n_inner - 1;  // last record of the cell-sorted arrays
    const int cntx = min(3, nc0), cnty = min(3, nc1), cntz = min(3, nc2);
    // Grids with <= 3 cells on some axis: adjacency does not determine the image, so the float pass lets everything through.
    const bool small = !(nc0 > 3 && nc1 > 3 && nc2 > 3);
    const float Lxf = (float)b.L[0], Lyf = (float)b.L[1], Lzf = (float)b.L[2];
    // float acceptance threshold for an exact far_sq: wrapped coordinates carry an absolute error below 2^-24 lmax each,
    // the shifted centre and the float arithmetic a few more roundings of that size
    const double reach_m = sqrt(P.far_sq) + 16.0 * 5.9604644775390625e-8 * fmax(b.L[0], fmax(b.L[1], b.L[2]));
    const float thr2 = small ? __int_as_float(0x7f800000) : __double2float_ru(reach_m * reach_m * (1.0 + 1e-6));
    uint2 *const queue = s_queue[warp];
    int qn = 0;  // the same in every lane of the warp
    auto drain = [&]() {
        __syncwarp();
        for (int i = lane; i < qn; i += 32) {
            const uint2 e = queue[i];
            const int o = warp * 32 + (int)e.x;
            double px, py, pz;
            int id;
            RecTraits<double>::load(P.grid.recs, (size_t)e.y, px, py, pz, id);
            // distVec = jPos - iPos, minimum image (:213-214)
            double dx, dy, dz;
            pmin_image_3(px, py, pz, s_ctr[0][o], s_ctr[1][o], s_ctr[2][o], b, dx, dy, dz);
            const double s = sumsq3<double>(dx, dy, dz);
            if (use_edges) {
                if (!(s <= s_edge[P.totbins])) continue;  // past the last bin (or NaN)
                int kb = min(max((int)(sqrtf((float)s) * inv_bw_f) + 1, 0), P.totbins);  // float seed, then the exact edges
                while (s > s_edge[kb]) ++kb;
                while (kb > 0 && s <= s_edge[kb - 1]) --kb;
                if (kb > 0) atomicAdd(s_cnt + kb - 1, 1u);  // bin 0 (dist == 0) is out of bounds in the Fortran
                continue;
            }
            if (!(s <= P.far_sq)) continue;  // certainly past the last bin, whatever the roundings of sqrt and division
            const double dist = __dsqrt_rn(s);
            const double nb = ceil(__ddiv_rn(dist, P.binwidth));
            if (!(nb >= 1.0) || !(nb <= (double)P.totbins)) continue;
            if (smem) atomicAdd(s_cnt + (int)nb - 1, 1u);
            else atomicAdd(P.counts + (int)nb - 1, 1ull);
        }
        __syncwarp();
        qn = 0;
    };
    for (int iz = 0; iz < cntz; ++iz) {
        int z = (nc2 <= 3) ? iz : cz - 1 + iz;
        float mz = wz;
        if (z < 0) { z += nc2; mz += Lzf; } else if (z >= nc2) { z -= nc2; mz -= Lzf; }
        for (int iy = 0; iy < cnty; ++iy) {
            int y = (nc1 <= 3) ? iy : cy - 1 + iy;
            float my = wy;
            if (y < 0) { y += nc1; my += Lyf; } else if (y >= nc1) { y -= nc1; my -= Lyf; }
            for (int ix = 0; ix < cntx; ++ix) {
                int x = (nc0 <= 3) ? ix : cx - 1 + ix;
                float mx = wx;
                if (x < 0) { x += nc0; mx += Lxf; } else if (x >= nc0) { x -= nc0; mx -= Lxf; }
                const size_t c = ((size_t)z * nc1 + y) * nc0 + x;
                int j0 = 0, len = 0;
                if (live) {
                    j0 = max((int)__ldg(P.grid.cell_start + c), j_min);  // (mode 1: do j = i + 1, NPos)
                    len = max((int)__ldg(P.grid.cell_start + c + 1) - j0, 0);
                    j0 = min(j0, j_last);  // (an empty run at the very end: keeps the loads below inside the array)
                }
                const int len_warp = __reduce_max_sync(kFullMask, len);
                for (int it = 0; it < len_warp; it += kPhUnroll) {
                    float4 w[kPhUnroll];
                    const float4 *wp = P.grid.wrapped + (j0 + it);
#pragma unroll
                    for (int u = 0; u < kPhUnroll; ++u)  // (lanes past their own run read on into the next cell, or the last atom, and drop it)
                        w[u] = __ldg(wp + min(u, j_last - (j0 + it)));
#pragma unroll
                    for (int u = 0; u < kPhUnroll; ++u) {
                        const float dx = w[u].x - mx, dy = w[u].y - my, dz = w[u].z - mz;
                        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        const bool take = it + u < len && r2 <= thr2;
                        const unsigned m = __ballot_sync(kFullMask, take);
                        if (take) queue[qn + __popc(m & lanes_below)] = make_uint2((unsigned)lane, (unsigned)(j0 + it + u));
                        qn += __popc(m);
                    }
                    if (qn > kPhQueue - 32 * kPhUnroll) drain();
                }
            }
        }
    }
    drain();
    if (smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < P.totbins; i += blockDim.x)
            if (s_cnt[i]) atomicAdd(P.counts + i, (unsigned long long)s_cnt[i]);
    }
}

// ---- psi (hexagonal order parameter) -------------------------------------------------------------------------
// One warp per centre: the lanes gather the neighbours inside (lowCut, highCut] (vectors as tetraCosAng sees them:
// reimaged twice, waterlib.f90:43-45 and :880-883) into a shared list, then share its K (K - 1) / 2 pairs.
// cos(6 theta) is the Chebyshev polynomial T6 of the clamped cosine, so no acos is needed; an exactly antiparallel
// pair (the -180 degrees of CosAngle3) gives T6(-1) = 1 = cos(-6 pi) and coincident positions (0 degrees) T6(1) = 1.
// The reference stores the complex mean of exp(6 i theta) into a real array (water_properties.py:428), which keeps
// only the real part: what it returns, and what is computed here, is | mean cos(6 theta) |.

constexpr int kPsiThreads = 128;
constexpr int kPsiCap = 320;  // neighbours per centre

struct PsiParams {
    PGrid grid;
    const double *box;
    const void *centres;
    int centre_dtype;
    int n_frames, n_pos, n_centres;
    double lowsq, highsq;
    double *psi;
    uint32_t *counters;
};

struct PsiSmem {
    Vec4<double> v[kPsiCap];
    int n;
};

__global__ void __launch_bounds__(kPsiThreads) psi_kernel(const PsiParams P) {
    extern __shared__ __align__(16) unsigned char psi_raw[];
    PsiSmem *S = reinterpret_cast<PsiSmem *>(psi_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PsiSmem &W = S[warp];
    const int nc0 = P.grid.nc0, nc1 = P.grid.nc1, nc2 = P.grid.nc2;
    const int cntx = min(3, nc0), cnty = min(3, nc1), cntz = min(3, nc2);
    const int ncell27 = cntx * cnty * cntz;
    const long long total = (long long)P.n_frames * P.n_centres;
    for (long long g = (long long)blockIdx.x * (kPsiThreads / 32) + warp; g < total; g += (long long)gridDim.x * (kPsiThreads / 32)) {
        const int f = (int)(g / P.n_centres);
        const PBox b = load_pbox(P.box + (size_t)f * 3);
        double rx, ry, rz;
        pload3<double>(P.centres, P.centre_dtype, (size_t)g, rx, ry, rz);
        const int cx = cell_coord(rx, b.iL[0], nc0), cy = cell_coord(ry, b.iL[1], nc1), cz = cell_coord(rz, b.iL[2], nc2);
        const int xs = (nc0 <= 3) ? 0 : (cx - 1 + nc0) % nc0, ys = (nc1 <= 3) ? 0 : (cy - 1 + nc1) % nc1,
                  zs = (nc2 <= 3) ? 0 : (cz - 1 + nc2) % nc2;
        const size_t cell_base = (size_t)f * nc0 * nc1 * nc2;
        const PFloat pf = make_pfloat(rx, ry, rz, b, sqrt(P.highsq));
        __syncwarp();
        if (lane == 0) W.n = 0;
        __syncwarp();
        for (int c27 = lane; c27 < ncell27; c27 += 32) {
            const int ix = c27 % cntx, iy = (c27 / cntx) % cnty, iz = c27 / (cntx * cnty);
            const size_t c = cell_base + ((size_t)((zs + iz) % nc2) * nc1 + (ys + iy) % nc1) * nc0 + (xs + ix) % nc0;
            const int j1 = (int)__ldg(P.grid.cell_start + c + 1);
            for (int j = (int)__ldg(P.grid.cell_start + c); j < j1; ++j) {
                if (!pfloat_near(pf, __ldg(P.grid.wrapped + j))) continue;  // certainly beyond highCut
                double px, py, pz;
                int id;
                RecTraits<double>::load(P.grid.recs, (size_t)j, px, py, pz, id);
                double dx, dy, dz;
                pmin_image_3(px, py, pz, rx, ry, rz, b, dx, dy, dz);
                const double s = sumsq3<double>(dx, dy, dz);
                if (!(s > P.lowsq && s <= P.highsq)) continue;
                // reimage: ref + d; tetraCosAng: ref + minimg((ref + d) - ref); CosAngle3: that - ref
                const double ex = __dsub_rn(__dadd_rn(rx, dx), rx), ey = __dsub_rn(__dadd_rn(ry, dy), ry),
                             ez = __dsub_rn(__dadd_rn(rz, dz), rz);
                const double d2x = __dsub_rn(ex, __dmul_rn(b.L[0], anint_exact<double>(__dmul_rn(ex, b.iL[0]))));
                const double d2y = __dsub_rn(ey, __dmul_rn(b.L[1], anint_exact<double>(__dmul_rn(ey, b.iL[1]))));
                const double d2z = __dsub_rn(ez, __dmul_rn(b.L[2], anint_exact<double>(__dmul_rn(ez, b.iL[2]))));
                Vec4<double> v;
                v.x = __dsub_rn(__dadd_rn(rx, d2x), rx);
                v.y = __dsub_rn(__dadd_rn(ry, d2y), ry);
                v.z = __dsub_rn(__dadd_rn(rz, d2z), rz);
                v.w = sumsq3<double>(v.x, v.y, v.z);
                const int at = atomicAdd(&W.n, 1);
                if (at < kPsiCap) W.v[at] = v;
            }
        }
        __syncwarp();
        const int n = W.n;
        if (n > kPsiCap) {
            if (lane == 0) atomicAdd(P.counters + kCntFatal, 1u);
            continue;
        }
        double acc = 0.0;
        const int npairs = n * (n - 1) / 2;
        for (int p = lane; p < npairs; p += 32) {
            int hi = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
            while (hi * (hi - 1) / 2 > p) --hi;
            while ((hi + 1) * hi / 2 <= p) ++hi;
            const int lo = p - hi * (hi - 1) / 2;
            const Vec4<double> va = W.v[lo], vb = W.v[hi];
            double c = 1.0;  // coincident positions: CosAngle3 returns 0 degrees (:690-693)
            if (va.w != 0.0 && vb.w != 0.0) c = clamped_cos<double>(dot3<double>(va.x, va.y, va.z, vb.x, vb.y, vb.z), va.w, vb.w);
            const double c2 = c * c;
            acc += ((32.0 * c2 - 48.0) * c2 + 18.0) * c2 - 1.0;  // T6(c) = cos(6 theta)
        }
        acc = warp_sum(acc);
        if (lane == 0) P.psi[g] = (n > 1) ? fabs(acc / (double)npairs) : 0.0;
    }
}

// ---- connected components of a symmetric 0/1 adjacency matrix (replaces sortlib's recursive depthFirstSort) ----------
// Minimum-label propagation with pointer jumping: every vertex repeatedly takes the smallest label among itself and
// its neighbours' labels, then shortcuts label chains; the host relaunches until nothing changes.

__global__ void components_init_kernel(int32_t *labels, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) labels[i] = i;
}

__global__ void __launch_bounds__(128) components_step_kernel(const int32_t *__restrict__ adj, int n, int32_t *labels, int32_t *changed) {
    // one warp per vertex: the lanes share its row
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= n) return;
    int best = labels[v];
    const int32_t *row = adj + (size_t)v * n;
    for (int j = lane; j < n; j += 32)
        if (row[j] == 1) best = min(best, labels[j]);
    best = __reduce_min_sync(kFullMask, best);
    if (lane == 0) {
        while (labels[best] < best) best = labels[best];  // pointer jumping
        if (best < labels[v]) {
            atomicMin(labels + v, best);
            *changed = 1;
        }
    }
}

static PGrid make_pgrid(void *workspace, const WorkspaceLayout &lay, const int32_t nc[3]) {
    char *ws = reinterpret_cast<char *>(workspace);
    PGrid g;
    g.cell_start = reinterpret_cast<const uint32_t *>(ws + lay.off_cell_start);
    g.recs = ws + lay.off_recs;
    g.wrapped = reinterpret_cast<const float4 *>(ws + lay.off_wrapped);
    g.nc0 = nc[0];
    g.nc1 = nc[1];
    g.nc2 = nc[2];
    return g;
}

}  // namespace wol

using namespace wol;


// ---- RadialDistPlane (fortran/waterlib.f90:237-314) -------------------------------------------------------------------
// The frame: v1 = P3 - P1, v2 = P2 - P1, v3 = v1 x v2, each minimum-imaged, v2 made orthogonal to v1 with the Fortran's
// expression v2 - (v1.v2) / (sum(v1 ** 2.0)) * v1, all normalised; Q(:, k) = v_k.  Every atom is minimum-imaged about the
// ORIGIN (Pos2 * BulkDens / BulkDens first: two roundings the reference performs), mapped with matmul(Q, .) -- libgfortran's
// matmul_r8 accumulates dest(x) += Q(x, n) * p(n) over n from a zeroed dest -- and counted in bin
// (ceiling(x' / w), ceiling(y' / w)) when |z'| <= 5.  A non-positive x' or y' indexes bin <= 0 in the Fortran, an
// out-of-bounds write; such atoms are skipped here and COUNTED in *n_bad so the caller can tell.
struct PlaneParams {
    const double *pos1;  // [3][3]
    const double *pos2;  // [n][3]
    const double *box;
    double binwidth, bulkdens;
    int totbins, n;
    unsigned long long *counts;  // [totbins][totbins], counts[kx][ky] (kx = x bin - 1)
    unsigned *n_bad;
};

__device__ __forceinline__ double plane_minimg(double v, double L) {
    const double iL = (L >= 0.0) ? __ddiv_rn(1.0, L) : 0.0;
    return __dsub_rn(v, __dmul_rn(L, anint_exact<double>(__dmul_rn(v, iL))));
}

__global__ void __launch_bounds__(128) radial_dist_plane_kernel(PlaneParams P) {
    __shared__ double Q[3][3];  // Q[row][col]: Q(row, col) = v_col(row)
    if (threadIdx.x == 0) {
        const double *p = P.pos1;
        double L[3] = {P.box[0], P.box[1], P.box[2]};
        double v1[3], v2[3], v3[3];
        for (int k = 0; k < 3; ++k) {
            v1[k] = __dsub_rn(p[2 * 3 + k], p[k]);
            v2[k] = __dsub_rn(p[1 * 3 + k], p[k]);
        }
        // crossProd3 (waterlib.f90:18-29): v3 = v1 x v2
        v3[0] = __dsub_rn(__dmul_rn(v1[1], v2[2]), __dmul_rn(v1[2], v2[1]));
        v3[1] = __dsub_rn(__dmul_rn(v1[2], v2[0]), __dmul_rn(v1[0], v2[2]));
        v3[2] = __dsub_rn(__dmul_rn(v1[0], v2[1]), __dmul_rn(v1[1], v2[0]));
        for (int k = 0; k < 3; ++k) {
            v1[k] = plane_minimg(v1[k], L[k]);
            v2[k] = plane_minimg(v2[k], L[k]);
            v3[k] = plane_minimg(v3[k], L[k]);
        }
        const double d12 = dot3<double>(v1[0], v1[1], v1[2], v2[0], v2[1], v2[2]);
        const double n1 = sumsq3<double>(v1[0], v1[1], v1[2]);  // v1 ** 2.0 is compiled as v1 * v1
        const double fac = __ddiv_rn(d12, n1);
        for (int k = 0; k < 3; ++k) v2[k] = __dsub_rn(v2[k], __dmul_rn(fac, v1[k]));
        const double l1 = __dsqrt_rn(sumsq3<double>(v1[0], v1[1], v1[2]));
        const double l2 = __dsqrt_rn(sumsq3<double>(v2[0], v2[1], v2[2]));
        const double l3 = __dsqrt_rn(sumsq3<double>(v3[0], v3[1], v3[2]));
        for (int k = 0; k < 3; ++k) {
            Q[k][0] = __ddiv_rn(v1[k], l1);
            Q[k][1] = __ddiv_rn(v2[k], l2);
            Q[k][2] = __ddiv_rn(v3[k], l3);
        }
    }
    __syncthreads();
    const double L[3] = {P.box[0], P.box[1], P.box[2]};
    // newPos1(1, :) = matmul(Q, 0) = 0: the slab is |z'| <= 5
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
        double q[3];
        for (int k = 0; k < 3; ++k) {
            const double scaled = __ddiv_rn(__dmul_rn(P.pos2[(size_t)i * 3 + k], P.bulkdens), P.bulkdens);
            q[k] = plane_minimg(scaled, L[k]);
        }
        double r[3];
        for (int x = 0; x < 3; ++x) {
            double acc = 0.0;
            for (int n = 0; n < 3; ++n) acc = __dadd_rn(acc, __dmul_rn(Q[x][n], q[n]));
            r[x] = acc;
        }
        if (r[2] <= 5.0 && r[2] >= -5.0) {
            const double bx = ceil(__ddiv_rn(r[0], P.binwidth)), by = ceil(__ddiv_rn(r[1], P.binwidth));
            if (bx <= (double)P.totbins && by <= (double)P.totbins) {
                if (bx >= 1.0 && by >= 1.0) atomicAdd(P.counts + ((size_t)((int)bx - 1) * P.totbins + ((int)by - 1)), 1ull);
                else atomicAdd(P.n_bad, 1u);
            }
        }
    }
}

// ---- np.histogramdd over two coordinates (numpy/lib/_histograms_impl.py: searchsorted(edges, x, 'right'), the last edge
// belongs to the last bin, everything outside is dropped) -- threeBodyCalc(output2D=True), orderParam_lib.py:1385-1393
__device__ __forceinline__ int edge_bin(const double *edges, int n_edges, double x) {
    int lo = 0, hi = n_edges;  // number of edges <= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edges[mid] <= x) lo = mid + 1; else hi = mid;
    }
    if (x == edges[n_edges - 1]) lo -= 1;
    return lo - 1;  // valid bins: 0 .. n_edges - 2
}

__global__ void __launch_bounds__(256) histogram2d_kernel(const double *__restrict__ x, const double *__restrict__ y, long long n,
                                                          const double *__restrict__ xedges, int nxe, const double *__restrict__ yedges,
                                                          int nye, unsigned long long *out) {
    extern __shared__ double s_edges[];
    double *sx = s_edges, *sy = s_edges + nxe;
    for (int i = threadIdx.x; i < nxe; i += blockDim.x) sx[i] = xedges[i];
    for (int i = threadIdx.x; i < nye; i += blockDim.x) sy[i] = yedges[i];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int bx = edge_bin(sx, nxe, x[i]), by = edge_bin(sy, nye, y[i]);
        if (bx >= 0 && bx < nxe - 1 && by >= 0 && by < nye - 1) atomicAdd(out + (size_t)bx * (nye - 1) + by, 1ull);
    }
}

extern "C" {

int wol_pair_hist(int32_t mode, const void *outer, int32_t outer_dtype, int32_t n_outer, const double *box, int32_t n_inner,
                  const int32_t nc[3], double edge_min, double binwidth, int32_t totbins, void *workspace, size_t workspace_bytes,
                  int64_t *counts, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (mode < 0 || mode > 2 || !outer || !box || !nc || !workspace || !counts) return set_error(WOL_ERR_INVALID, "wol_pair_hist: bad argument");
    if (totbins < 1 || !(binwidth > 0.0) || n_outer < 0 || n_inner < 0) return set_error(WOL_ERR_INVALID, "wol_pair_hist: bad bin spec or size");
    if (mode == 1 && n_outer != n_inner) return set_error(WOL_ERR_INVALID, "wol_pair_hist: mode 1 pairs a set with itself (n_outer == n_inner)");
    const double reach = binwidth * totbins;
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && reach * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "histogram range %.6g exceeds the planned cell edge %.6g", reach, edge_min);
    const WorkspaceLayout lay = workspace_layout(1, n_inner, n_inner, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    PairHistParams P;
    P.grid = make_pgrid(workspace, lay, nc);
    P.box = box;
    P.outer = outer;
    P.outer_dtype = outer_dtype;
    P.n_outer = n_outer;
    P.n_inner = n_inner;
    P.mode = mode;
    P.binwidth = binwidth;
    P.totbins = totbins;
    // dist^2 > r^2 (1 + 1e-9), r = totbins binwidth  =>  sqrt rounds to more than r (1 + 4e-10), the quotient by
    // binwidth to more than totbins, and ceiling() lands beyond the last bin
    P.far_sq = ((double)totbins * binwidth) * ((double)totbins * binwidth) * (1.0 + 1e-9);
    P.counts = reinterpret_cast<unsigned long long *>(counts);
    if (n_outer > 0 && n_inner > 0) {
        // shared bins + the table of bin edges in distance^2 (8-byte aligned behind the bins)
        const size_t smem = totbins <= kMaxSmemBins ? sizeof(unsigned) * ((totbins + 1) & ~1) + sizeof(double) * (totbins + 1) : 0;
        if (smem > 16 * 1024) {
            cudaError_t ea = cudaFuncSetAttribute(pair_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (ea != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(pair_hist)", ea);
        }
        pair_hist_kernel<<<(n_outer + 127) / 128, 128, smem, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_pair_hist", e);
    return WOL_OK;
}

int wol_radial_dist_plane(const double *pos1, const double *pos2, int32_t n_pos2, const double *box, double binwidth, int32_t totbins,
                          double bulkdens, int64_t *counts, int32_t *n_out_of_bounds, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!pos1 || !box || !counts || !n_out_of_bounds || (!pos2 && n_pos2 > 0))
        return set_error(WOL_ERR_INVALID, "wol_radial_dist_plane: null argument");
    if (totbins < 1 || !(binwidth > 0.0) || n_pos2 < 0 || bulkdens == 0.0)
        return set_error(WOL_ERR_INVALID, "wol_radial_dist_plane: need totbins >= 1, binwidth > 0, a non-zero BulkDens");
    PlaneParams P;
    P.pos1 = pos1;
    P.pos2 = pos2;
    P.box = box;
    P.binwidth = binwidth;
    P.bulkdens = bulkdens;
    P.totbins = totbins;
    P.n = n_pos2;
    P.counts = reinterpret_cast<unsigned long long *>(counts);
    P.n_bad = reinterpret_cast<unsigned *>(n_out_of_bounds);
    if (n_pos2 > 0) {
        const int blocks = (n_pos2 + 127) / 128 < sm_count() * 8 ? (n_pos2 + 127) / 128 : sm_count() * 8;
        radial_dist_plane_kernel<<<blocks, 128, 0, stream>>>(P);
        add_launches(1);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_radial_dist_plane", e);
    return WOL_OK;
}

int wol_histogram2d(const double *x, const double *y, int64_t n, const double *xedges, int32_t n_xedges, const double *yedges,
                    int32_t n_yedges, int64_t *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || n_xedges < 2 || n_yedges < 2 || !xedges || !yedges || !out || ((!x || !y) && n > 0))
        return set_error(WOL_ERR_INVALID, "wol_histogram2d: bad argument");
    if ((size_t)(n_xedges + n_yedges) * sizeof(double) > 40000) return set_error(WOL_ERR_RANGE, "wol_histogram2d: more than 5000 edges");
    if (n > 0) {
        long long blocks = (n + 255) / 256;
        if (blocks > sm_count() * 16) blocks = sm_count() * 16;
        histogram2d_kernel<<<(unsigned)blocks, 256, (size_t)(n_xedges + n_yedges) * sizeof(double), stream>>>(
            x, y, n, xedges, n_xedges, yedges, n_yedges, reinterpret_cast<unsigned long long *>(out));
        add_launches(1);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_histogram2d", e);
    return WOL_OK;
}

int wol_components(const int32_t *adj, int32_t n, int32_t *labels, int32_t *changed, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || !labels || !changed || (!adj && n > 0)) return set_error(WOL_ERR_INVALID, "wol_components: bad argument");
    if (n == 0) return WOL_OK;
    components_init_kernel<<<(n + 127) / 128, 128, 0, stream>>>(labels, n);
    add_launches(1);
    for (int it = 0; it < n + 1; ++it) {  // converges in O(diameter) sweeps; n + 1 is the hard bound
        int32_t flag = 0;
        cudaError_t e = cudaMemsetAsync(changed, 0, sizeof(int32_t), stream);
        if (e != cudaSuccess) return set_cuda_error("wol_components", e);
        components_step_kernel<<<(unsigned)(((long long)n * 32 + 127) / 128), 128, 0, stream>>>(adj, n, labels, changed);
        add_launches(1);
        e = cudaMemcpyAsync(&flag, changed, sizeof(int32_t), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return set_cuda_error("wol_components", e);
        if (!flag) break;
    }
    return WOL_OK;
}

int wol_psi(const void *centres, int32_t centre_dtype, const double *box, int32_t n_frames, int32_t n_pos, int32_t n_centres,
            const int32_t nc[3], double edge_min, double lowcut, double highcut, void *workspace, size_t workspace_bytes, double *psi,
            void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!centres || !box || !nc || !workspace || !psi) return set_error(WOL_ERR_INVALID, "wol_psi: null argument");
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && highcut * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "cutoff %.6g exceeds the planned cell edge %.6g", highcut, edge_min);
    const WorkspaceLayout lay = workspace_layout(n_frames, n_pos, n_centres, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    PsiParams P;
    P.grid = make_pgrid(workspace, lay, nc);
    P.box = box;
    P.centres = centres;
    P.centre_dtype = centre_dtype;
    P.n_frames = n_frames;
    P.n_pos = n_pos;
    P.n_centres = n_centres;
    P.lowsq = lowcut * lowcut;
    P.highsq = highcut * highcut;
    P.psi = psi;
    P.counters = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(workspace) + lay.off_counters);
    const long long total = (long long)n_frames * n_centres;
    if (total > 0) {
        const size_t smem = sizeof(PsiSmem) * (kPsiThreads / 32);
        cudaError_t e = cudaFuncSetAttribute(psi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(psi)", e);
        long long blocks = (total + 3) / 4;
        const long long cap = (long long)sm_count() * 4;
        if (blocks > cap) blocks = cap;
        psi_kernel<<<(unsigned)blocks, kPsiThreads, smem, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_psi", e);
    return WOL_OK;
}

}  // extern "C"
