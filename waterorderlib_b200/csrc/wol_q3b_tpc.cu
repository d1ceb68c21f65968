// K2, fast path: one THREAD per centre (sm_100a), used when every axis has at least 4 cells.
//
//   phase 1  float prefilter.  The thread walks the 9 cell rows of its 27-cell stencil (each row is one
//            contiguous x-run of the cell-sorted array, two when the run wraps) over box-wrapped float
//            coordinates; the periodic image is a per-row / per-run shift of the CENTRE, so a candidate
//            costs one 16-byte load, 3 FADD, 1 FMUL, 2 FFMA and a compare.  Survivors (squared distance
//            below cutoff + rounding margin) go to a per-thread shared-memory index list.
//   phase 2  exact re-evaluation of the survivors from the original fp64 coordinates in the reference's
//            operation order: cutoff tests, reimaged difference vectors and norms of the three-body
//            neighbours (kept in shared memory), register top-4 by (distance, atom index).
//   phase 3a three-body pairs, flattened over the WARP: a shuffle scan of the per-centre pair counts
//            and a 5-step search give each lane an equal share of the warp's pairs (K varies 2..9 per
//            centre, so per-thread loops would idle most lanes).  Clamped cosine -> bin by threshold
//            table -> shared-memory histogram.
//   phase 3b q from the four winners, per thread.
// Centres whose search is not provably complete inside the stencil, or whose lists overflow, are queued
// for the large-capacity pass (wol_q3b.cu), exactly like the generic group-per-centre path does.
//
// Why the prefilter cannot lose a neighbour: wrapped coordinates carry an absolute error below
// 2^-24 L per axis, the shifted centre another 2 * 2^-24 L, the subtraction one more rounding of a value
// below 2 L; the acceptance threshold is widened by 16 * 2^-24 * Lmax (see q3b_tpc_launch), several
// times the worst case.  For >= 4 cells per axis the image implied by cell adjacency IS the minimum
// image for every candidate closer than two cell edges.
#include <stdlib.h>

#include "wol_q3b_f32.cuh"

namespace wol {

constexpr int kTpcThreads = 256;  // two blocks per SM: the bin table and block histograms are shared by 8 warps
constexpr int kTpcListCap = 16;  // prefilter survivors per centre
constexpr int kTpcEntCap = 10;   // three-body neighbours per centre
constexpr int kTpcMaxPairs = kTpcEntCap * (kTpcEntCap - 1) / 2;

struct TpcSmem {
    Vec4<double> ent[kTpcEntCap][kTpcThreads];
    int lj[kTpcListCap][kTpcThreads];
    int woff[kTpcThreads / 32][33];
    unsigned char pair_ab[kTpcMaxPairs + 3];
};

template <bool EXACT>
__global__ void __launch_bounds__(kTpcThreads, 2) q3b_tpc_kernel(const __grid_constant__ Q3bParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TpcSmem &S = *reinterpret_cast<TpcSmem *>(smem_raw);
    unsigned char *after = smem_raw + sizeof(TpcSmem);
    const bool smem_hist = P.do_3b && P.ang_hist && P.nbins <= kMaxSmemBins;
    const bool smem_qhist = P.do_q && P.q_hist && P.q_nbins <= kMaxSmemBins;
    const bool smem_tab = P.nbins <= kMaxSmemBins;
    const int tab_len = P.do_3b ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
    double *s_tab = reinterpret_cast<double *>(after);
    unsigned *s_hist = reinterpret_cast<unsigned *>(after + (smem_tab ? sizeof(double) * tab_len : 0));
    unsigned *s_qhist = s_hist + (smem_hist ? P.nbins : 0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (smem_tab)
        for (int i = tid; i < tab_len; i += kTpcThreads) s_tab[i] = P.table[i];
    if (smem_hist)
        for (int i = tid; i < P.nbins; i += kTpcThreads) s_hist[i] = 0u;
    if (smem_qhist)
        for (int i = tid; i < P.q_nbins; i += kTpcThreads) s_qhist[i] = 0u;
    if (tid < kTpcMaxPairs) {
        // p = b (b - 1) / 2 + a, a < b
        int b = 1;
        while ((b + 1) * b / 2 <= tid) ++b;
        S.pair_ab[tid] = (unsigned char)((tid - b * (b - 1) / 2) | (b << 4));
    }
    __syncthreads();
    const double *tab = smem_tab ? s_tab : P.table;
    const double inv_width = (double)P.nbins / (P.hist_hi - P.hist_lo);
    const float hist_lo_f = (float)P.hist_lo, inv_width_f = (float)inv_width;  // the bin search is seeded in float
    const double tet_c_hi = P.do_3b ? tab[P.nbins + 3] : 0.0, tet_c_lo = P.do_3b ? tab[P.nbins + 4] : 0.0;  // cosines of the tetrahedral window
    const bool do3 = P.do_3b != 0, doq = P.do_q != 0;
    const double low3sq = P.low3sq, high3sq = P.high3sq, lowqsq = P.lowqsq, highqsq = P.highqsq;
    const bool last1 = P.wq_max <= 1;
    const double selsq1 = last1 ? highqsq : fmin(highqsq, fmin(P.highq, P.rc1) * fmin(P.highq, P.rc1));
    const float pre_thr2 = P.pre_thr2;
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;

    LaneStats st;
    st.reset();
    // Tiles are dealt to the blocks in round-robin chunks of P.chunk_tiles: at any moment the whole grid works
    // on one neighbourhood of one frame, so the stencil rows a block reads are (or soon will be) in L2 on
    // behalf of its neighbours and the DRAM traffic stays near one read of the frame.  Within a chunk the
    // tiles are consecutive, which keeps the rows shared by consecutive tiles in L1.
    const long long n_chunks = (P.total_tiles + P.chunk_tiles - 1) / P.chunk_tiles;
    int cur_f = -1;
    double Lx = 1, Ly = 1, Lz = 1, iLx = 1, iLy = 1, iLz = 1;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x)
    for (long long tile = chunk * P.chunk_tiles, t_end = min(P.total_tiles, tile + P.chunk_tiles); tile < t_end; ++tile) {
        const int f = (int)(tile / P.tiles_per_frame);
        const int m = (int)(tile - (long long)f * P.tiles_per_frame) * kTpcThreads + tid;
        if (f != cur_f) {
            if (cur_f >= 0) {
                flush_stats(P, cur_f, st);
                if ((smem_hist || smem_qhist) && P.hist_per_frame) {
                    __syncthreads();
                    flush_hist(P, smem_hist ? s_hist : nullptr, smem_qhist ? s_qhist : nullptr, cur_f, true);
                    __syncthreads();
                }
            }
            cur_f = f;
            Lx = P.box[(size_t)f * 3 + 0];
            Ly = P.box[(size_t)f * 3 + 1];
            Lz = P.box[(size_t)f * 3 + 2];
            iLx = __ddiv_rn(1.0, Lx);
            iLy = __ddiv_rn(1.0, Ly);
            iLz = __ddiv_rn(1.0, Lz);
        }
        const bool valid = m < P.n_centres && (P.n_valid == nullptr || m < __ldg(P.n_valid + f));
        double rx = 0, ry = 0, rz = 0;
        float wx = 0, wy = 0, wz = 0;
        int cx = 0, cy = 0, cz = 0, self_j = -1;
        size_t out_index = 0;
        uint32_t fb_id = 0;
        if (valid) {
            if (P.centres == nullptr) {
                const size_t j = (size_t)f * P.n_pos + m;
                const int4 *rp = reinterpret_cast<const int4 *>(reinterpret_cast<const RecD *>(P.recs) + j);
                const int4 a = __ldg(rp), b = __ldg(rp + 1);
                rx = __hiloint2double(a.y, a.x);
                ry = __hiloint2double(a.w, a.z);
                rz = __hiloint2double(b.y, b.x);
                cx = b.w & 1023;
                cy = (b.w >> 10) & 1023;
                cz = (b.w >> 20) & 1023;
                const float4 w = __ldg(P.wrapped + j);
                wx = w.x; wy = w.y; wz = w.z;
                self_j = (int)j;
                out_index = (size_t)f * P.n_pos + b.z;
                fb_id = (uint32_t)j;
            } else {
                load_centre<double>(P, f, m, rx, ry, rz);
                cx = cell_coord(rx, iLx, nc0);
                cy = cell_coord(ry, iLy, nc1);
                cz = cell_coord(rz, iLz, nc2);
                wx = wrapped_coord(rx, Lx, iLx);
                wy = wrapped_coord(ry, Ly, iLy);
                wz = wrapped_coord(rz, Lz, iLz);
                out_index = (size_t)f * P.n_centres + m;
                fb_id = (uint32_t)out_index;
            }
        }

        // ---------------- phase 1: float prefilter over the 9 rows of the stencil -----------------
        int nl = 0;
        if (valid) {
            const float Lxf = (float)Lx, Lyf = (float)Ly, Lzf = (float)Lz;
            const size_t cell_base = (size_t)f * nc0 * nc1 * nc2;
            // x-run: cells [cx-1, cx+1]; the part that falls off the row comes from the other end
            const int xa0 = max(cx - 1, 0), xa1 = min(cx + 1, nc0 - 1) + 1;
            int xb0 = 0, xb1 = 0;
            float sxb = 0.f;
            if (cx == 0) {
                xb0 = nc0 - 1; xb1 = nc0; sxb = Lxf;      // neighbours near x = L are images at x - L
            } else if (cx == nc0 - 1) {
                xb0 = 0; xb1 = 1; sxb = -Lxf;             // neighbours near x = 0 are images at x + L
            }
            const float4 *wr = P.wrapped;
            const int j_last = P.n_frames * P.n_pos - 1;
            // One z-plane of the stencil at a time (keeps the loop body small enough for the instruction cache);
            // inside a plane: the 6 row bounds first (independent loads in flight together), then the 3 rows.
#pragma unroll 1
            for (int pz = -1; pz <= 1; ++pz) {
                int z = cz + pz;
                float czs = wz;
                if (z < 0) { z += nc2; czs += Lzf; } else if (z >= nc2) { z -= nc2; czs -= Lzf; }
                int rj0[3], rj1[3];
                float rcy[3];
#pragma unroll
                for (int row = 0; row < 3; ++row) {
                    int y = cy + row - 1;
                    float cys = wy;
                    if (y < 0) { y += nc1; cys += Lyf; } else if (y >= nc1) { y -= nc1; cys -= Lyf; }
                    const uint32_t *cs = P.cell_start + cell_base + ((size_t)z * nc1 + y) * nc0;
                    rj0[row] = (int)__ldg(cs + xa0);
                    rj1[row] = (int)__ldg(cs + xa1);
                    rcy[row] = cys;
                }
                // Software pipeline: the first pair of the NEXT row and the next pair of THIS row are in flight while
                // a pair is evaluated (rows hold ~4 candidates).  Prefetch indices are clamped into the array; a
                // prefetched element beyond its row is never used.
                float4 p0 = __ldg(wr + min(rj0[0], j_last)), p1 = __ldg(wr + min(rj0[0] + 1, j_last));
#pragma unroll
                for (int row = 0; row < 3; ++row) {
                    const float cys = rcy[row];
                    const int j1 = rj1[row];
                    int j = rj0[row];
                    float4 w0 = p0, w1 = p1;
                    if (row < 2) {
                        p0 = __ldg(wr + min(rj0[row < 2 ? row + 1 : 2], j_last));
                        p1 = __ldg(wr + min(rj0[row < 2 ? row + 1 : 2] + 1, j_last));
                    }
                    while (j < j1) {
                        const int jn = j + 2;
                        float4 n0 = w0, n1 = w1;
                        if (jn < j1) {
                            n0 = __ldg(wr + jn);
                            n1 = __ldg(wr + min(jn + 1, j_last));
                        }
                        const float ax = w0.x - wx, ay = w0.y - cys, az = w0.z - czs;
                        const float bx = w1.x - wx, by = w1.y - cys, bz = w1.z - czs;
                        const float ra = fmaf(az, az, fmaf(ay, ay, ax * ax));
                        const float rb = fmaf(bz, bz, fmaf(by, by, bx * bx));
                        if (ra <= pre_thr2 && j != self_j) {
                            if (nl < kTpcListCap) S.lj[nl][tid] = j;
                            ++nl;
                        }
                        if (j + 1 < j1 && rb <= pre_thr2 && j + 1 != self_j) {
                            if (nl < kTpcListCap) S.lj[nl][tid] = j + 1;
                            ++nl;
                        }
                        w0 = n0; w1 = n1; j = jn;
                    }
                }
            }
            if (xb1 != 0) {  // the wrapped end of the x-run (first / last cell column only)
                const float cxs = wx + sxb;
#pragma unroll 1
                for (int row = 0; row < 9; ++row) {
                    const int dz = row / 3 - 1, dy = row % 3 - 1;
                    int y = cy + dy, z = cz + dz;
                    float cys = wy, czs = wz;
                    if (y < 0) { y += nc1; cys += Lyf; } else if (y >= nc1) { y -= nc1; cys -= Lyf; }
                    if (z < 0) { z += nc2; czs += Lzf; } else if (z >= nc2) { z -= nc2; czs -= Lzf; }
                    const uint32_t *cs = P.cell_start + cell_base + ((size_t)z * nc1 + y) * nc0;
                    const int j1 = (int)__ldg(cs + xb1);
                    for (int j = (int)__ldg(cs + xb0); j < j1; ++j) {
                        const float4 w = __ldg(wr + j);
                        const float dx = w.x - cxs, dyv = w.y - cys, dzv = w.z - czs;
                        const float r2 = fmaf(dzv, dzv, fmaf(dyv, dyv, dx * dx));
                        if (r2 <= pre_thr2 && j != self_j) {
                            if (nl < kTpcListCap) S.lj[nl][tid] = j;
                            ++nl;
                        }
                    }
                }
            }
        }
        bool overflow = nl > kTpcListCap;

        // ---------------- phase 2: exact fp64 re-evaluation of the survivors ----------------------
        Top4<double> top;
        top.reset();
        int K3 = 0, nq = 0;
        if (valid && !overflow) {
            double nx = 0, ny = 0, nz = 0;
            int nidx = 0, nj = 0;
            if (nl > 0) {
                nj = S.lj[0][tid];
                RecTraits<double>::load(P.recs, (size_t)nj, nx, ny, nz, nidx);
            }
            for (int k = 0; k < nl; ++k) {
                const int j = nj, idx = nidx;
                const double px = nx, py = ny, pz = nz;
                if (k + 1 < nl) {  // next survivor's record is in flight while this one is evaluated
                    nj = S.lj[k + 1][tid];
                    RecTraits<double>::load(P.recs, (size_t)nj, nx, ny, nz, nidx);
                }
                const double dx = min_image_1<double, EXACT>(px, rx, Lx, iLx);
                const double dy = min_image_1<double, EXACT>(py, ry, Ly, iLy);
                const double dz = min_image_1<double, EXACT>(pz, rz, Lz, iLz);
                const double s = sumsq3<double>(dx, dy, dz);
                const bool in3 = do3 && (s > low3sq) && (s <= high3sq);
                const bool inq = doq && (s > lowqsq) && (s <= selsq1);
                if (in3 || inq) {
                    // ReimagedPos = RefPos + distvec (waterlib.f90:45); Vec = Pos - RefPos (:694-695)
                    Vec4<double> v;
                    v.x = __dsub_rn(__dadd_rn(rx, dx), rx);
                    v.y = __dsub_rn(__dadd_rn(ry, dy), ry);
                    v.z = __dsub_rn(__dadd_rn(rz, dz), rz);
                    v.w = sumsq3<double>(v.x, v.y, v.z);
                    if (in3) {
                        if (K3 < kTpcEntCap) S.ent[K3][tid] = v;
                        ++K3;
                    }
                    if (inq) {
                        ++nq;
                        top.insert(__dsqrt_rn(v.w), idx, j);
                    }
                }
            }
            if (K3 > kTpcEntCap) overflow = true;
        }
        bool q_go = valid && doq && !overflow;
        bool b3_go = valid && do3 && !overflow;
        if (valid && overflow) {
            const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
            P.fb_list[at] = fb_id | (do3 ? kFbNeed3b : 0u) | (doq ? kFbNeedQ : 0u);
            atomicAdd(P.counters + kCntOverflow, 1u);
        }
        if (q_go && nq < 4 && !last1) {
            // fewer than four inside the radius the stencil guarantees: the widened search decides
            const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
            P.fb_list[at] = fb_id | kFbNeedQ;
            atomicAdd(P.counters + kCntWidened, 1u);
            q_go = false;
        }

        // ---------------- phase 3a: three-body pairs, flattened over the warp ---------------------
        if (do3) {
            const int npair = b3_go ? K3 * (K3 - 1) / 2 : 0;
            int inc = npair;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(kFullMask, inc, o);
                if (lane >= o) inc += n;
            }
            const int total = __shfl_sync(kFullMask, inc, 31);
            __syncwarp();
            S.woff[warp][lane] = inc - npair;
            if (lane == 31) S.woff[warp][32] = total;
            __syncwarp();
            const int *woff = S.woff[warp];
            for (int w = lane; w < total; w += 32) {
                int t = 0, base = 0;  // woff[0] == 0; the search keeps woff[t] so that no load follows it
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int v = woff[t + step];
                    if (v <= w) {
                        t += step;
                        base = v;
                    }
                }
                const int p = w - base;
                const int ab = S.pair_ab[p];
                const int col = warp * 32 + t;
                const Vec4<double> va = S.ent[ab & 15][col], vb = S.ent[ab >> 4][col];
                int pos;
                if (va.w == 0.0 || vb.w == 0.0) {  // coincident positions: CosAngle3 returns 0 (:690-693)
                    pos = (int)tab[P.nbins + 2];
                } else {
                    const double dot = dot3<double>(va.x, va.y, va.z, vb.x, vb.y, vb.z);
                    const double c = clamped_cos<double>(dot, va.w, vb.w);
                    pos = angle_position(c, tab, P.nbins, hist_lo_f, inv_width_f);
                    if (c != -1.0 && c <= tet_c_hi && c >= tet_c_lo) {
                        st.tet_count += 1u;
                        st.tet_cos += c;
                        st.tet_cossq += c * c;
                    }
                }
                st.n_angles += 1u;
                if (pos >= 0 && pos < P.nbins) {
                    if (smem_hist) atomicAdd(s_hist + pos, 1u);
                    else if (P.ang_hist) atomicAdd(P.ang_hist + (size_t)(P.hist_per_frame ? f : 0) * P.nbins + pos, 1ull);
                }
            }
            __syncwarp();  // every lane is done with this tile's lists before the next tile refills them
            if (b3_go) {
                if (P.n3) P.n3[out_index] = K3;
                st.n_neigh += (unsigned)K3;
            }
        }

        // ---------------- phase 3b: q from the four winners ---------------------------------------
        if (q_go)
            finish_q<EXACT>(P, f, rx, ry, rz, Lx, Ly, Lz, iLx, iLy, iLz, top, min(nq, 4), out_index, st, smem_qhist ? s_qhist : nullptr);
    }
    if (cur_f >= 0) flush_stats(P, cur_f, st);
    if (smem_hist || smem_qhist) {
        __syncthreads();
        if (cur_f >= 0) flush_hist(P, smem_hist ? s_hist : nullptr, smem_qhist ? s_qhist : nullptr, cur_f, false);
    }
}

// ------------------------------------------------------------------------------------------------
// Widened q search, one THREAD per queued centre (a centre with fewer than four neighbours inside the radius
// the 27-cell stencil guarantees).  The four nearest are almost always just outside that radius, so the search
// is driven by a BOUND instead of by the stencil size:
//   bound   the thread scans its 27-cell stencil over the float coordinates and keeps the four smallest
//           distances^2 (only candidates certainly beyond lowCut count).  Whatever else there is can only
//           matter if it is closer than the 4th of them (+ rounding slack, capped at the radius a
//           half-width-2 stencil guarantees).
//   collect the 25 rows of the half-width-2 stencil, each clipped to the cells the bound can reach: a row is
//           skipped when the (y, z) distance from the centre to the row's box already exceeds the bound, and
//           its x-run shrinks to the cells within sqrt(bound^2 - that distance^2).  In practice a few
//           cells behind the faces the sphere pokes through.  Survivors go to a per-thread shared list.
//   exact   survivors re-evaluated in fp64 reference arithmetic, register top-4 by (distance, atom index), q.
// A centre still short of four neighbours inside the guaranteed radius (or with more survivors than the list
// holds) goes to the second-level queue (group-per-centre pass, half-width 3 and up).
constexpr int kWidenThreads = 128;
constexpr int kWidenListCap = 16;

// distance from a point at offset u inside its own cell (edge e) to the cell d cells away, along one axis
__device__ __forceinline__ float cell_gap(int d, float u, float e) {
    return d > 0 ? (float)d * e - u : (d < 0 ? u + (float)(-d - 1) * e : 0.f);
}

// One contiguous run of records [j0, j1) seen from a (shifted) centre.  MODE 0: keep the four smallest
// distances^2 beyond `lo2`.  MODE 1: append everything within `thr` to the thread's list.
template <int MODE>
__device__ __forceinline__ void widen_scan(const float4 *__restrict__ wr, int j0, int j1, float sx, float sy, float sz, float lo2,
                                           float thr, int self_j, float &a0, float &a1, float &a2, float &a3, int *list, int &nl) {
    for (int j = j0; j < j1; ++j) {
        const float4 w = __ldg(wr + j);
        const float dx = w.x - sx, dy = w.y - sy, dz = w.z - sz;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (MODE == 0) {
            if (r2 > lo2) {
                float v = r2, m;
                m = fminf(a0, v); v = fmaxf(a0, v); a0 = m;
                m = fminf(a1, v); v = fmaxf(a1, v); a1 = m;
                m = fminf(a2, v); v = fmaxf(a2, v); a2 = m;
                a3 = fminf(a3, v);
            }
        } else if (r2 <= thr && j != self_j) {
            if (nl < kWidenListCap) list[nl * kWidenThreads] = j;
            ++nl;
        }
    }
}

// The x-run of cells [cx + lo, cx + hi] of row `cs` (lo, hi within [-2, 2]), split where it wraps around the box.
template <int MODE>
__device__ __forceinline__ void widen_row(const float4 *__restrict__ wr, const uint32_t *__restrict__ cs, int nc0, int cx, int lo,
                                          int hi, float wx, float sy, float sz, float Lxf, float lo2, float thr, int self_j,
                                          float &a0, float &a1, float &a2, float &a3, int *list, int &nl) {
    const int x0 = cx + lo, x1 = cx + hi;  // inclusive, may leave [0, nc0)
    const int m0 = max(x0, 0), m1 = min(x1, nc0 - 1);
    widen_scan<MODE>(wr, (int)__ldg(cs + m0), (int)__ldg(cs + m1 + 1), wx, sy, sz, lo2, thr, self_j, a0, a1, a2, a3, list, nl);
    if (x0 < 0)  // cells x0 + nc0 .. nc0 - 1 hold the images at x - L
        widen_scan<MODE>(wr, (int)__ldg(cs + x0 + nc0), (int)__ldg(cs + nc0), wx + Lxf, sy, sz, lo2, thr, self_j, a0, a1, a2, a3, list, nl);
    if (x1 >= nc0)  // cells 0 .. x1 - nc0 hold the images at x + L
        widen_scan<MODE>(wr, (int)__ldg(cs), (int)__ldg(cs + x1 - nc0 + 1), wx - Lxf, sy, sz, lo2, thr, self_j, a0, a1, a2, a3, list, nl);
}

template <bool EXACT, bool F32>
__global__ void __launch_bounds__(kWidenThreads) q3b_tpc_widen_kernel(const __grid_constant__ Q3bParams P) {
    __shared__ int s_list[kWidenListCap * kWidenThreads];
    extern __shared__ unsigned s_qhist_dyn[];
    const uint32_t n_items = P.counters[P.list_counter];
    {
        const uint32_t n_widened = P.counters[kCntWidened];
        if (n_widened == 0u || n_widened < P.widen_lo || n_widened >= P.widen_hi) return;
    }
    // one shared row only: the queue mixes frames, so per-frame histograms go straight to global memory
    unsigned *s_qhist = (P.q_hist && !P.hist_per_frame && P.q_nbins <= kMaxSmemBins) ? s_qhist_dyn : nullptr;
    if (s_qhist) {
        for (int i = threadIdx.x; i < P.q_nbins; i += kWidenThreads) s_qhist[i] = 0u;
        __syncthreads();
    }
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
    const double lowqsq = P.lowqsq, highqsq = P.highqsq;
    const bool last2 = P.wq_max <= 2;
    const double rsel = fmin(P.highq, 2.0 * P.rc1);
    const double selsq2 = last2 ? highqsq : fmin(highqsq, rsel * rsel);
    const float lowq_hi2 = P.lowq_hi2, thr_cap = P.pre_thr2_w2, cst = P.pre_cst_w2, eps = P.cell_eps;
    const float kInf = __int_as_float(0x7f800000);
    const float4 *__restrict__ wr = P.wrapped;
    int *list = s_list + threadIdx.x;
    // warp-uniform trip count: the statistics of a warp's centres are combined with shuffles
    for (uint32_t it0 = blockIdx.x * kWidenThreads + (threadIdx.x & ~31u); it0 < n_items; it0 += gridDim.x * kWidenThreads) {
        const uint32_t it = it0 + (threadIdx.x & 31u);
        const uint32_t e = it < n_items ? P.list[it] : 0u;
        // list overflows (three-body redo) belong to the large-capacity pass
        const bool valid = it < n_items && (e & kFbNeed3b) == 0u && (e & kFbNeedQ) != 0u;
        bool done = false;
        int f = 0;
        LaneStats st;
        st.reset();
        if (valid) {
        const uint32_t id = e & kFbIdMask;
        double rx = 0, ry = 0, rz = 0;
        float wx, wy, wz;
        int cx, cy, cz, self_j = -1;
        size_t out_index;
        if (P.centres == nullptr && F32) {  // FP32 records: everything the search needs is in the wrapped array
            f = (int)(id / (uint32_t)P.n_pos);
            const float4 w = __ldg(wr + id);
            const uint32_t cp = __ldg(P.cellpack + id);
            wx = w.x; wy = w.y; wz = w.z;
            cx = cp & 1023; cy = (cp >> 10) & 1023; cz = (cp >> 20) & 1023;
            self_j = (int)id;
            out_index = (size_t)f * P.n_pos + __float_as_int(w.w);
        } else if (P.centres == nullptr) {
            f = (int)(id / (uint32_t)P.n_pos);
            const int4 *rp = reinterpret_cast<const int4 *>(reinterpret_cast<const RecD *>(P.recs) + id);
            const int4 a = __ldg(rp), b = __ldg(rp + 1);
            rx = __hiloint2double(a.y, a.x);
            ry = __hiloint2double(a.w, a.z);
            rz = __hiloint2double(b.y, b.x);
            cx = b.w & 1023;
            cy = (b.w >> 10) & 1023;
            cz = (b.w >> 20) & 1023;
            const float4 w = __ldg(wr + id);
            wx = w.x; wy = w.y; wz = w.z;
            self_j = (int)id;
            out_index = (size_t)f * P.n_pos + b.z;
        } else {
            f = (int)(id / (uint32_t)P.n_centres);
            load_centre<double>(P, f, (int)(id - (uint32_t)f * P.n_centres), rx, ry, rz);
            out_index = id;
        }
        const double Lx = P.box[(size_t)f * 3 + 0], Ly = P.box[(size_t)f * 3 + 1], Lz = P.box[(size_t)f * 3 + 2];
        const double iLx = __ddiv_rn(1.0, Lx), iLy = __ddiv_rn(1.0, Ly), iLz = __ddiv_rn(1.0, Lz);
        if (P.centres != nullptr) {
            cx = cell_coord(rx, iLx, nc0);
            cy = cell_coord(ry, iLy, nc1);
            cz = cell_coord(rz, iLz, nc2);
            wx = wrapped_coord(rx, Lx, iLx);
            wy = wrapped_coord(ry, Ly, iLy);
            wz = wrapped_coord(rz, Lz, iLz);
        }
        const float Lxf = (float)Lx, Lyf = (float)Ly, Lzf = (float)Lz;
        const uint32_t *cs_frame = P.cell_start + (size_t)f * nc0 * nc1 * nc2;

        // ---- bound: four smallest float distances^2 inside the 27-cell stencil ---------------------
        float a0 = kInf, a1 = kInf, a2 = kInf, a3 = kInf;
        int nl = 0;
#pragma unroll 1
        for (int row = 0; row < 9; ++row) {
            int y = cy + row % 3 - 1, z = cz + row / 3 - 1;
            float sy = wy, sz = wz;
            if (y < 0) { y += nc1; sy += Lyf; } else if (y >= nc1) { y -= nc1; sy -= Lyf; }
            if (z < 0) { z += nc2; sz += Lzf; } else if (z >= nc2) { z -= nc2; sz -= Lzf; }
            widen_row<0>(wr, cs_frame + ((size_t)z * nc1 + y) * nc0, nc0, cx, -1, 1, wx, sy, sz, Lxf, lowq_hi2, 0.f, self_j, a0, a1,
                         a2, a3, list, nl);
        }
        const float thr = fminf(a3 + cst, thr_cap);

        // ---- collect: the rows and cells of the half-width-2 stencil the bound can reach -----------
        {
            const float ex = Lxf / (float)nc0, ey = Lyf / (float)nc1, ez = Lzf / (float)nc2, iex = 1.0f / ex;
            const float ux = wx - (float)cx * ex, uy = wy - (float)cy * ey, uz = wz - (float)cz * ez;
#pragma unroll 1
            for (int row = 0; row < 25; ++row) {
                const int oy = row % 5 - 2, oz = row / 5 - 2;
                const float gy = fmaxf(cell_gap(oy, uy, ey) - eps, 0.f), gz = fmaxf(cell_gap(oz, uz, ez) - eps, 0.f);
                const float left = thr - fmaf(gz, gz, gy * gy);
                if (left < 0.f) continue;
                const float reach = sqrtf(left) + eps;
                const int lo = max(-2, (int)floorf((ux - reach) * iex)), hi = min(2, (int)floorf((ux + reach) * iex));
                int y = cy + oy, z = cz + oz;
                float sy = wy, sz = wz;
                if (y < 0) { y += nc1; sy += Lyf; } else if (y >= nc1) { y -= nc1; sy -= Lyf; }
                if (z < 0) { z += nc2; sz += Lzf; } else if (z >= nc2) { z -= nc2; sz -= Lzf; }
                widen_row<1>(wr, cs_frame + ((size_t)z * nc1 + y) * nc0, nc0, cx, lo, hi, wx, sy, sz, Lxf, 0.f, thr, self_j, a0, a1,
                             a2, a3, list, nl);
            }
        }

        if (F32) {
            // ---- FP32 mode: the float minimum-image distance IS the distance --------------------------------
            const float lowqsq_f = (float)lowqsq, selsq2_f = (float)selsq2;
            const float iLxf = 1.0f / Lxf, iLyf = 1.0f / Lyf, iLzf = 1.0f / Lzf;
            Top4F top;
            top.reset();
            int nq = 0;
            if (nl <= kWidenListCap) {
                for (int k = 0; k < nl; ++k) {
                    const float4 w = __ldg(wr + list[k * kWidenThreads]);
                    float dx = w.x - wx, dy = w.y - wy, dz = w.z - wz;
                    dx -= Lxf * rintf(dx * iLxf);
                    dy -= Lyf * rintf(dy * iLyf);
                    dz -= Lzf * rintf(dz * iLzf);
                    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (r2 > lowqsq_f && r2 <= selsq2_f) {
                        ++nq;
                        top.insert(r2, __float_as_int(w.w), dx, dy, dz);
                    }
                }
            }
            if (nl > kWidenListCap || (nq < 4 && !last2)) {
                const uint32_t at = atomicAdd(P.counters + kCntLevel2, 1u);
                P.list2[at] = id | kFbNeedQ;
            } else {
                finish_q32(P, f, top, min(nq, 4), out_index, st, s_qhist, hist_spec(0.0, 1.0, P.q_nbins));
                done = true;
            }
        } else {
        // ---- exact evaluation of the survivors ------------------------------------------------------
        Top4<double> top;
        top.reset();
        int nq = 0;
        if (nl <= kWidenListCap) {
            for (int k = 0; k < nl; ++k) {
                const int j = list[k * kWidenThreads];
                double px, py, pz;
                int idx;
                RecTraits<double>::load(P.recs, (size_t)j, px, py, pz, idx);
                const double dx = min_image_1<double, EXACT>(px, rx, Lx, iLx);
                const double dy = min_image_1<double, EXACT>(py, ry, Ly, iLy);
                const double dz = min_image_1<double, EXACT>(pz, rz, Lz, iLz);
                const double s = sumsq3<double>(dx, dy, dz);
                if ((s > lowqsq) && (s <= selsq2)) {
                    ++nq;
                    const double vx = __dsub_rn(__dadd_rn(rx, dx), rx);
                    const double vy = __dsub_rn(__dadd_rn(ry, dy), ry);
                    const double vz = __dsub_rn(__dadd_rn(rz, dz), rz);
                    top.insert(__dsqrt_rn(sumsq3<double>(vx, vy, vz)), idx, j);
                }
            }
        }
        if (nl > kWidenListCap || (nq < 4 && !last2)) {
            const uint32_t at = atomicAdd(P.counters + kCntLevel2, 1u);
            P.list2[at] = id | kFbNeedQ;
        } else {
            finish_q<EXACT>(P, f, rx, ry, rz, Lx, Ly, Lz, iLx, iLy, iLz, top, min(nq, 4), out_index, st, s_qhist);
            done = true;
        }
        }  // fp64 evaluation
        }  // valid
        // frame statistics: one set of atomics per warp when its centres share a frame (the queue is nearly
        // frame-ordered), per thread otherwise
        const unsigned fin = __ballot_sync(kFullMask, done);
        if (fin != 0u && P.stats) {
            const int f0 = __shfl_sync(kFullMask, f, __ffs(fin) - 1);
            if (__all_sync(kFullMask, !done || f == f0)) {
                const double s1 = warp_sum(st.q_sum), s2 = warp_sum(st.q_sumsq);
                if ((threadIdx.x & 31) == 0) {
                    atomicAdd(P.stats + (size_t)f0 * WOL_NSTATS + WOL_STAT_Q_SUM, s1);
                    atomicAdd(P.stats + (size_t)f0 * WOL_NSTATS + WOL_STAT_Q_SUMSQ, s2);
                    atomicAdd(P.stats + (size_t)f0 * WOL_NSTATS + WOL_STAT_N_CENTRES, (double)__popc(fin));
                }
            } else if (done) {
                atomicAdd(P.stats + (size_t)f * WOL_NSTATS + WOL_STAT_Q_SUM, st.q_sum);
                atomicAdd(P.stats + (size_t)f * WOL_NSTATS + WOL_STAT_Q_SUMSQ, st.q_sumsq);
                atomicAdd(P.stats + (size_t)f * WOL_NSTATS + WOL_STAT_N_CENTRES, 1.0);
            }
        }
    }
    if (s_qhist) {
        __syncthreads();
        flush_bins(s_qhist, P.q_hist, P.q_nbins, false);
    }
}

// ------------------------------------------------------------------------------------------------
// The same search, one WARP per queued centre (fp64 records).  A batch queues a few hundred centres per million;
// with a thread per centre that is a handful of warps each walking ~150 cells one dependent load after the other --
// a ~100 us floor per call whatever the count, which a host-fed pipeline pays once per batch.  Here the lanes share
// the cells: 27 lanes take one cell each for the bound, 25 lanes one row each for the collection, one lane per
// survivor for the exact arithmetic; the four nearest come out of four warp-wide (distance, index) minima.  Same
// thresholds, same exact evaluation and the same finish_q as the thread-per-centre version above.
constexpr int kWidenWarps = 8;        // warps per block
constexpr int kWidenWarpCap = 32;     // survivors per centre: one lane each

__device__ __forceinline__ void warp_key_min(double &d, int &i, int &p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(kFullMask, d, o);
        const int oi = __shfl_xor_sync(kFullMask, i, o), op = __shfl_xor_sync(kFullMask, p, o);
        if (key_less(od, oi, d, i)) { d = od; i = oi; p = op; }
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(kWidenWarps * 32) q3b_widen_warp_kernel(const __grid_constant__ Q3bParams P) {
    __shared__ int s_list[kWidenWarps][kWidenWarpCap];
    __shared__ int s_count[kWidenWarps];
    extern __shared__ unsigned s_qhist_dyn[];
    const uint32_t n_items = P.counters[P.list_counter];
    {
        const uint32_t n_widened = P.counters[kCntWidened];
        if (n_widened == 0u || n_widened < P.widen_lo || n_widened >= P.widen_hi) return;
    }
    unsigned *s_qhist = (P.q_hist && !P.hist_per_frame && P.q_nbins <= kMaxSmemBins) ? s_qhist_dyn : nullptr;
    if (s_qhist) {
        for (int i = threadIdx.x; i < P.q_nbins; i += kWidenWarps * 32) s_qhist[i] = 0u;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
    const double lowqsq = P.lowqsq, highqsq = P.highqsq;
    const bool last2 = P.wq_max <= 2;
    const double rsel = fmin(P.highq, 2.0 * P.rc1);
    const double selsq2 = last2 ? highqsq : fmin(highqsq, rsel * rsel);
    const float lowq_hi2 = P.lowq_hi2, thr_cap = P.pre_thr2_w2, cst = P.pre_cst_w2, eps = P.cell_eps;
    const float kInf = __int_as_float(0x7f800000);
    const float4 *__restrict__ wr = P.wrapped;
    int *list = s_list[warp];
    LaneStats st;  // lane 0 accumulates; flushed when the frame changes
    st.reset();
    int st_f = -1;
    for (uint32_t it = blockIdx.x * kWidenWarps + warp; it < n_items; it += gridDim.x * kWidenWarps) {
        const uint32_t e = P.list[it];
        if ((e & kFbNeed3b) != 0u || (e & kFbNeedQ) == 0u) continue;  // list overflows belong to the large-capacity pass
        const uint32_t id = e & kFbIdMask;
        double rx, ry, rz;
        float wx, wy, wz;
        int cx, cy, cz, self_j = -1, f;
        size_t out_index;
        if (P.centres == nullptr) {
            f = (int)(id / (uint32_t)P.n_pos);
            const int4 *rp = reinterpret_cast<const int4 *>(reinterpret_cast<const RecD *>(P.recs) + id);
            const int4 a = __ldg(rp), b = __ldg(rp + 1);
            rx = __hiloint2double(a.y, a.x);
            ry = __hiloint2double(a.w, a.z);
            rz = __hiloint2double(b.y, b.x);
            cx = b.w & 1023;
            cy = (b.w >> 10) & 1023;
            cz = (b.w >> 20) & 1023;
            const float4 w = __ldg(wr + id);
            wx = w.x; wy = w.y; wz = w.z;
            self_j = (int)id;
            out_index = (size_t)f * P.n_pos + b.z;
        } else {
            f = (int)(id / (uint32_t)P.n_centres);
            load_centre<double>(P, f, (int)(id - (uint32_t)f * P.n_centres), rx, ry, rz);
            out_index = id;
        }
        const double Lx = P.box[(size_t)f * 3 + 0], Ly = P.box[(size_t)f * 3 + 1], Lz = P.box[(size_t)f * 3 + 2];
        const double iLx = __ddiv_rn(1.0, Lx), iLy = __ddiv_rn(1.0, Ly), iLz = __ddiv_rn(1.0, Lz);
        if (P.centres != nullptr) {
            cx = cell_coord(rx, iLx, nc0);
            cy = cell_coord(ry, iLy, nc1);
            cz = cell_coord(rz, iLz, nc2);
            wx = wrapped_coord(rx, Lx, iLx);
            wy = wrapped_coord(ry, Ly, iLy);
            wz = wrapped_coord(rz, Lz, iLz);
        }
        const float Lxf = (float)Lx, Lyf = (float)Ly, Lzf = (float)Lz;
        const uint32_t *cs_frame = P.cell_start + (size_t)f * nc0 * nc1 * nc2;
        if (lane == 0) s_count[warp] = 0;
        __syncwarp();

        // ---- bound: 4th smallest float distance^2 inside the 27-cell stencil, one cell per lane -------------------
        float a0 = kInf, a1 = kInf, a2 = kInf, a3 = kInf;
        int nl_unused = 0;
        if (lane < 27) {
            const int ox = lane % 3 - 1, oy = (lane / 3) % 3 - 1, oz = lane / 9 - 1;
            int y = cy + oy, z = cz + oz;
            float sy = wy, sz = wz;
            if (y < 0) { y += nc1; sy += Lyf; } else if (y >= nc1) { y -= nc1; sy -= Lyf; }
            if (z < 0) { z += nc2; sz += Lzf; } else if (z >= nc2) { z -= nc2; sz -= Lzf; }
            widen_row<0>(wr, cs_frame + ((size_t)z * nc1 + y) * nc0, nc0, cx, ox, ox, wx, sy, sz, Lxf, lowq_hi2, 0.f, self_j, a0, a1, a2,
                         a3, list, nl_unused);
        }
        float b4 = kInf;  // 4th smallest over the warp: four rounds of "take the minimum of the lanes' heads"
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            float m = a0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(kFullMask, m, o));
            b4 = m;
            const unsigned who = __ballot_sync(kFullMask, a0 == m && m < kInf);
            if (who != 0u && lane == __ffs(who) - 1) { a0 = a1; a1 = a2; a2 = a3; a3 = kInf; }
        }
        const float thr = fminf(b4 + cst, thr_cap);

        // ---- collect: the rows of the half-width-2 stencil the bound can reach, one row per lane --------------------
        if (lane < 25) {
            const float ex = Lxf / (float)nc0, ey = Lyf / (float)nc1, ez = Lzf / (float)nc2, iex = 1.0f / ex;
            const float ux = wx - (float)cx * ex, uy = wy - (float)cy * ey, uz = wz - (float)cz * ez;
            const int oy = lane % 5 - 2, oz = lane / 5 - 2;
            const float gy = fmaxf(cell_gap(oy, uy, ey) - eps, 0.f), gz = fmaxf(cell_gap(oz, uz, ez) - eps, 0.f);
            const float left = thr - fmaf(gz, gz, gy * gy);
            if (left >= 0.f) {
                const float reach = sqrtf(left) + eps;
                const int lo = max(-2, (int)floorf((ux - reach) * iex)), hi = min(2, (int)floorf((ux + reach) * iex));
                int y = cy + oy, z = cz + oz;
                float sy = wy, sz = wz;
                if (y < 0) { y += nc1; sy += Lyf; } else if (y >= nc1) { y -= nc1; sy -= Lyf; }
                if (z < 0) { z += nc2; sz += Lzf; } else if (z >= nc2) { z -= nc2; sz -= Lzf; }
                const uint32_t *cs = cs_frame + ((size_t)z * nc1 + y) * nc0;
                // the x-run [cx + lo, cx + hi], split where it wraps around the box (as widen_row does)
                for (int piece = 0; piece < 3; ++piece) {
                    const int x0 = cx + lo, x1 = cx + hi;
                    int j0, j1;
                    float sx;
                    if (piece == 0) {
                        const int m0 = max(x0, 0), m1 = min(x1, nc0 - 1);
                        if (m0 > m1) continue;
                        j0 = (int)__ldg(cs + m0); j1 = (int)__ldg(cs + m1 + 1); sx = wx;
                    } else if (piece == 1) {
                        if (x0 >= 0) continue;
                        j0 = (int)__ldg(cs + x0 + nc0); j1 = (int)__ldg(cs + nc0); sx = wx + Lxf;
                    } else {
                        if (x1 < nc0) continue;
                        j0 = (int)__ldg(cs); j1 = (int)__ldg(cs + x1 - nc0 + 1); sx = wx - Lxf;
                    }
                    for (int jb = j0; jb < j1; jb += 4) {  // four loads in flight: the run is a chain of cache misses otherwise
                        float4 w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) w[u] = __ldg(wr + min(jb + u, j1 - 1));
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = jb + u;
                            const float dx = w[u].x - sx, dy = w[u].y - sy, dz = w[u].z - sz;
                            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            if (j < j1 && r2 <= thr && j != self_j) {
                                const int at = atomicAdd(&s_count[warp], 1);
                                if (at < kWidenWarpCap) list[at] = j;
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        const int nl = s_count[warp];

        // ---- exact evaluation, one survivor per lane; the four nearest by (distance, atom index) --------------------
        double key = Ops<double>::inf();
        int kidx = INT_MAX, kj = -1;
        bool inq = false;
        if (nl <= kWidenWarpCap && lane < nl) {
            const int j = list[lane];
            double px, py, pz;
            int idx;
            RecTraits<double>::load(P.recs, (size_t)j, px, py, pz, idx);
            const double dx = min_image_1<double, EXACT>(px, rx, Lx, iLx);
            const double dy = min_image_1<double, EXACT>(py, ry, Ly, iLy);
            const double dz = min_image_1<double, EXACT>(pz, rz, Lz, iLz);
            const double s = sumsq3<double>(dx, dy, dz);
            if ((s > lowqsq) && (s <= selsq2)) {
                inq = true;
                const double vx = __dsub_rn(__dadd_rn(rx, dx), rx);
                const double vy = __dsub_rn(__dadd_rn(ry, dy), ry);
                const double vz = __dsub_rn(__dadd_rn(rz, dz), rz);
                key = __dsqrt_rn(sumsq3<double>(vx, vy, vz));
                kidx = idx;
                kj = j;
            }
        }
        const int nq = __popc(__ballot_sync(kFullMask, inq));
        Top4<double> top;
        top.reset();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double d = key;
            int i = kidx, pj = kj;
            warp_key_min(d, i, pj);
            top.d[k] = d; top.i[k] = i; top.p[k] = pj;
            if (pj >= 0 && pj == kj) { key = Ops<double>::inf(); kidx = INT_MAX; kj = -1; }  // the winner leaves the pool
        }
        if (lane == 0) {
            if (nl > kWidenWarpCap || (nq < 4 && !last2)) {
                const uint32_t at = atomicAdd(P.counters + kCntLevel2, 1u);
                P.list2[at] = id | kFbNeedQ;
            } else {
                if (f != st_f) {
                    if (st_f >= 0 && P.stats) {
                        atomicAdd(P.stats + (size_t)st_f * WOL_NSTATS + WOL_STAT_Q_SUM, st.q_sum);
                        atomicAdd(P.stats + (size_t)st_f * WOL_NSTATS + WOL_STAT_Q_SUMSQ, st.q_sumsq);
                        atomicAdd(P.stats + (size_t)st_f * WOL_NSTATS + WOL_STAT_N_CENTRES, (double)st.n_centres);
                    }
                    st.reset();
                    st_f = f;
                }
                finish_q<EXACT>(P, f, rx, ry, rz, Lx, Ly, Lz, iLx, iLy, iLz, top, min(nq, 4), out_index, st, s_qhist);
            }
        }
        __syncwarp();
    }
    if (lane == 0 && st_f >= 0 && P.stats) {
        atomicAdd(P.stats + (size_t)st_f * WOL_NSTATS + WOL_STAT_Q_SUM, st.q_sum);
        atomicAdd(P.stats + (size_t)st_f * WOL_NSTATS + WOL_STAT_Q_SUMSQ, st.q_sumsq);
        atomicAdd(P.stats + (size_t)st_f * WOL_NSTATS + WOL_STAT_N_CENTRES, (double)st.n_centres);
    }
    if (s_qhist) {
        __syncthreads();
        flush_bins(s_qhist, P.q_hist, P.q_nbins, false);
    }
}

// adjacency image == minimum image needs 3 cell edges < L / 2
bool q3b_tpc_widen_supported(const Q3bParams &P) {
    return P.wrapped != nullptr && P.nc0 >= 7 && P.nc1 >= 7 && P.nc2 >= 7;
}

// How many widened centres make the thread-per-centre kernel the faster one (fp64 records).  One warp per centre has short
// dependent chains -- 39 us for 6 000 centres, where a thread per centre has a ~100 us floor -- but costs 4.2 us per 1000
// centres against 0.5: a liquid-like 8 x 1M-water batch queues 160 000 (678 us against 181 us).
constexpr uint32_t kWidenPerThreadFrom = 24576;

int q3b_tpc_widen_launch(const Q3bParams &P0, cudaStream_t stream, bool exact, bool f32) {
    Q3bParams P = P0;
    P.widen_lo = 0u;
    P.widen_hi = 0xffffffffu;
    const int grid = sm_count() * 8;
    const size_t smem = (P.q_hist && !P.hist_per_frame && P.q_nbins <= kMaxSmemBins) ? sizeof(unsigned) * P.q_nbins : 0;
    if (f32) {
        q3b_tpc_widen_kernel<false, true><<<grid, kWidenThreads, smem, stream>>>(P);
        add_launches(1);
        return WOL_OK;
    }
    // fp64 records: BOTH kernels are launched and the device-side count picks the one that works (the host does not know
    // the count without a synchronisation; the other launch returns at once).  WOL_WIDEN_THREAD=1 / =0 force one of them.
    const char *env = getenv("WOL_WIDEN_THREAD");
    const bool only_thread = env && env[0] == '1', only_warp = env && env[0] == '0';
    if (!only_thread) {
        P.widen_lo = 0u;
        P.widen_hi = only_warp ? 0xffffffffu : kWidenPerThreadFrom;
        if (exact) q3b_widen_warp_kernel<true><<<sm_count() * 16, kWidenWarps * 32, smem, stream>>>(P);
        else q3b_widen_warp_kernel<false><<<sm_count() * 16, kWidenWarps * 32, smem, stream>>>(P);
        add_launches(1);
    }
    if (!only_warp) {
        P.widen_lo = only_thread ? 0u : kWidenPerThreadFrom;
        P.widen_hi = 0xffffffffu;
        if (exact) q3b_tpc_widen_kernel<true, false><<<grid, kWidenThreads, smem, stream>>>(P);
        else q3b_tpc_widen_kernel<false, false><<<grid, kWidenThreads, smem, stream>>>(P);
        add_launches(1);
    }
    return WOL_OK;
}

bool q3b_tpc_supported(const Q3bParams &P) {
    return P.wrapped != nullptr && P.nc0 >= 4 && P.nc1 >= 4 && P.nc2 >= 4;
}

template <bool EXACT>
static int launch_tpc(const Q3bParams &P0, cudaStream_t stream) {
    Q3bParams P = P0;
    P.tiles_per_frame = (P.n_centres + kTpcThreads - 1) / kTpcThreads;
    P.total_tiles = (long long)P.tiles_per_frame * P.n_frames;
    const int tab_len = P.do_3b ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
    size_t smem = sizeof(TpcSmem);
    if (P.nbins <= kMaxSmemBins) smem += sizeof(double) * tab_len;
    if (P.do_3b && P.ang_hist && P.nbins <= kMaxSmemBins) smem += sizeof(unsigned) * P.nbins;
    if (P.do_q && P.q_hist && P.q_nbins <= kMaxSmemBins) smem += sizeof(unsigned) * P.q_nbins;
    cudaError_t e = cudaFuncSetAttribute(q3b_tpc_kernel<EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(tpc)", e);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, q3b_tpc_kernel<EXACT>, kTpcThreads, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > P.total_tiles) grid = P.total_tiles;
    P.chunk_tiles = 1;  // round-robin tiles (measured: -12 % kernel time against contiguous per-block ranges)
    if (grid > 0) {
        q3b_tpc_kernel<EXACT><<<(unsigned)grid, kTpcThreads, smem, stream>>>(P);
        add_launches(1);
    }
    return WOL_OK;
}

int q3b_tpc_launch(const Q3bParams &P, cudaStream_t stream, bool exact) {
    return exact ? launch_tpc<true>(P, stream) : launch_tpc<false>(P, stream);
}

}  // namespace wol
