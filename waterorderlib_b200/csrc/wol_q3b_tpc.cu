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
#include "wol_q3b_common.cuh"

namespace wol {

constexpr int kTpcThreads = 128;
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
__global__ void __launch_bounds__(kTpcThreads, 4) q3b_tpc_kernel(const __grid_constant__ Q3bParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TpcSmem &S = *reinterpret_cast<TpcSmem *>(smem_raw);
    unsigned char *after = smem_raw + sizeof(TpcSmem);
    const bool smem_hist = P.do_3b && P.ang_hist && P.nbins <= kMaxSmemBins;
    const bool smem_tab = P.nbins <= kMaxSmemBins;
    const int tab_len = P.do_3b ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
    double *s_tab = reinterpret_cast<double *>(after);
    unsigned *s_hist = reinterpret_cast<unsigned *>(after + (smem_tab ? sizeof(double) * tab_len : 0));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (smem_tab)
        for (int i = tid; i < tab_len; i += kTpcThreads) s_tab[i] = P.table[i];
    if (smem_hist)
        for (int i = tid; i < P.nbins; i += kTpcThreads) s_hist[i] = 0u;
    if (tid < kTpcMaxPairs) {
        // p = b (b - 1) / 2 + a, a < b
        int b = 1;
        while ((b + 1) * b / 2 <= tid) ++b;
        S.pair_ab[tid] = (unsigned char)((tid - b * (b - 1) / 2) | (b << 4));
    }
    __syncthreads();
    const double *tab = smem_tab ? s_tab : P.table;
    const double inv_width = (double)P.nbins / (P.hist_hi - P.hist_lo);
    const bool do3 = P.do_3b != 0, doq = P.do_q != 0;
    const double low3sq = P.low3sq, high3sq = P.high3sq, lowqsq = P.lowqsq, highqsq = P.highqsq;
    const bool last1 = P.wq_max <= 1;
    const double selsq1 = last1 ? highqsq : fmin(highqsq, fmin(P.highq, P.rc1) * fmin(P.highq, P.rc1));
    const float pre_thr2 = P.pre_thr2;
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;

    LaneStats st;
    st.reset();
    const long long chunk = (P.total_tiles + gridDim.x - 1) / gridDim.x;
    const long long t_begin = chunk * blockIdx.x;
    const long long t_end = min(P.total_tiles, t_begin + chunk);
    int cur_f = -1;
    double Lx = 1, Ly = 1, Lz = 1, iLx = 1, iLy = 1, iLz = 1;
    for (long long tile = t_begin; tile < t_end; ++tile) {
        const int f = (int)(tile / P.tiles_per_frame);
        const int m = (int)(tile - (long long)f * P.tiles_per_frame) * kTpcThreads + tid;
        if (f != cur_f) {
            if (cur_f >= 0) {
                flush_stats(P, cur_f, st);
                if (smem_hist && P.hist_per_frame) {
                    __syncthreads();
                    flush_hist(P, s_hist, cur_f, true);
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
        const bool valid = m < P.n_centres;
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
            // all 18 row bounds first (independent loads in flight together), then the rows
            int rj0[9], rj1[9];
            float rcy[9], rcz[9];
#pragma unroll
            for (int row = 0; row < 9; ++row) {
                const int dz = row / 3 - 1, dy = row % 3 - 1;
                int y = cy + dy, z = cz + dz;
                float cys = wy, czs = wz;
                if (y < 0) { y += nc1; cys += Lyf; } else if (y >= nc1) { y -= nc1; cys -= Lyf; }
                if (z < 0) { z += nc2; czs += Lzf; } else if (z >= nc2) { z -= nc2; czs -= Lzf; }
                const uint32_t *cs = P.cell_start + cell_base + ((size_t)z * nc1 + y) * nc0;
                rj0[row] = (int)__ldg(cs + xa0);
                rj1[row] = (int)__ldg(cs + xa1);
                rcy[row] = cys;
                rcz[row] = czs;
            }
            const float4 *wr = P.wrapped;
#pragma unroll
            for (int row = 0; row < 9; ++row) {
                const float cys = rcy[row], czs = rcz[row];
                const int j1 = rj1[row];
                for (int j = rj0[row]; j < j1; j += 2) {
                    const bool two = j + 1 < j1;
                    const float4 w0 = __ldg(wr + j);
                    const float4 w1 = __ldg(wr + (two ? j + 1 : j));
                    const float ax = w0.x - wx, ay = w0.y - cys, az = w0.z - czs;
                    const float bx = w1.x - wx, by = w1.y - cys, bz = w1.z - czs;
                    const float ra = fmaf(az, az, fmaf(ay, ay, ax * ax));
                    const float rb = fmaf(bz, bz, fmaf(by, by, bx * bx));
                    if (ra <= pre_thr2 && j != self_j) {
                        if (nl < kTpcListCap) S.lj[nl][tid] = j;
                        ++nl;
                    }
                    if (two && rb <= pre_thr2 && j + 1 != self_j) {
                        if (nl < kTpcListCap) S.lj[nl][tid] = j + 1;
                        ++nl;
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
                int t = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (woff[t + step] <= w) t += step;
                const int p = w - woff[t];
                const int ab = S.pair_ab[p];
                const int col = warp * 32 + t;
                const Vec4<double> va = S.ent[ab & 15][col], vb = S.ent[ab >> 4][col];
                int pos;
                if (va.w == 0.0 || vb.w == 0.0) {  // coincident positions: CosAngle3 returns 0 (:690-693)
                    pos = (int)tab[P.nbins + 2];
                } else {
                    const double dot = dot3<double>(va.x, va.y, va.z, vb.x, vb.y, vb.z);
                    const double c = clamped_cos<double>(dot, va.w, vb.w);
                    pos = angle_position(c, tab, P.nbins, P.hist_lo, inv_width);
                    if (c != -1.0 && c <= tab[P.nbins + 3] && c >= tab[P.nbins + 4]) {
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
        if (q_go) finish_q<EXACT>(P, f, rx, ry, rz, Lx, Ly, Lz, iLx, iLy, iLz, top, min(nq, 4), out_index, st);
    }
    if (cur_f >= 0) flush_stats(P, cur_f, st);
    if (smem_hist) {
        __syncthreads();
        if (cur_f >= 0) flush_hist(P, s_hist, cur_f, false);
    }
}

// ------------------------------------------------------------------------------------------------
// Widened q search, one WARP per queued centre: half-width-2 stencil = 25 rows of a 5-cell x-run, one row
// per lane, so the 25 dependent load chains of a centre run side by side.
//   pass 1  every lane scans its row over the float coordinates and keeps its four smallest distances^2
//           (only candidates certainly beyond lowCut count); four REDUX-min rounds give the warp's 4th
//           smallest;
//   pass 2  every lane rescans its row and appends whatever lies within that bound (+ rounding slack,
//           capped at the radius the stencil guarantees) to a shared list -- a handful of candidates;
//   exact   lane k evaluates survivor k in fp64 reference arithmetic; ranks by (distance, index) through
//           shuffles pick the four winners; lanes 0-3 build their vectors, lanes 0-5 the pair terms.
// A centre still short of four neighbours inside the guaranteed radius goes to the second-level queue
// (group-per-centre pass, half-width 3 and up).
constexpr int kWidenThreads = 128;
constexpr int kWidenWarps = kWidenThreads / 32;

__device__ __forceinline__ void widen_scan_row(const float4 *__restrict__ wr, int j0, int j1, float cxs, float cys, float czs,
                                               int pass, float lowq_hi2, float thr, int self_j, float &a0, float &a1,
                                               float &a2, float &a3, int *s_cnt, int *s_list) {
    for (int j = j0; j < j1; ++j) {
        const float4 w = __ldg(wr + j);
        const float dx = w.x - cxs, dyv = w.y - cys, dzv = w.z - czs;
        const float r2 = fmaf(dzv, dzv, fmaf(dyv, dyv, dx * dx));
        if (pass == 0) {
            if (r2 > lowq_hi2) {  // insert into the lane's sorted quadruple
                float v = r2, m;
                m = fminf(a0, v); v = fmaxf(a0, v); a0 = m;
                m = fminf(a1, v); v = fmaxf(a1, v); a1 = m;
                m = fminf(a2, v); v = fmaxf(a2, v); a2 = m;
                a3 = fminf(a3, v);
            }
        } else if (r2 <= thr && j != self_j) {
            const int at = atomicAdd(s_cnt, 1);
            if (at < 32) s_list[at] = j;
        }
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(kWidenThreads) q3b_tpc_widen_kernel(const __grid_constant__ Q3bParams P) {
    __shared__ int s_list[kWidenWarps][32];
    __shared__ int s_cnt[kWidenWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_items = P.counters[P.list_counter];
    if (P.counters[kCntWidened] == 0u) return;
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
    const double lowqsq = P.lowqsq, highqsq = P.highqsq;
    const bool last2 = P.wq_max <= 2;
    const double rsel = fmin(P.highq, 2.0 * P.rc1);
    const double selsq2 = last2 ? highqsq : fmin(highqsq, rsel * rsel);
    const float lowq_hi2 = P.lowq_hi2, thr_cap = P.pre_thr2_w2, cst = P.pre_cst_w2;
    const float kInf = __int_as_float(0x7f800000);
    const uint32_t warps_total = gridDim.x * kWidenWarps;
    for (uint32_t it = blockIdx.x * kWidenWarps + warp; it < n_items; it += warps_total) {
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
            const float4 w = __ldg(P.wrapped + id);
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
        // this lane's row (lanes 25..31 have none) and the two pieces of its x-run [cx-2, cx+2]
        int ja0 = 0, ja1 = 0, jb0 = 0, jb1 = 0;
        float cys = wy, czs = wz, cxb = wx;
        if (lane < 25) {
            const int dz = lane / 5 - 2, dy = lane % 5 - 2;
            int y = cy + dy, z = cz + dz;
            if (y < 0) { y += nc1; cys += Lyf; } else if (y >= nc1) { y -= nc1; cys -= Lyf; }
            if (z < 0) { z += nc2; czs += Lzf; } else if (z >= nc2) { z -= nc2; czs -= Lzf; }
            const uint32_t *cs = P.cell_start + (size_t)f * nc0 * nc1 * nc2 + ((size_t)z * nc1 + y) * nc0;
            ja0 = (int)__ldg(cs + max(cx - 2, 0));
            ja1 = (int)__ldg(cs + min(cx + 2, nc0 - 1) + 1);
            if (cx < 2) {
                jb0 = (int)__ldg(cs + nc0 - (2 - cx));
                jb1 = (int)__ldg(cs + nc0);
                cxb = wx + Lxf;
            } else if (cx > nc0 - 3) {
                jb0 = (int)__ldg(cs);
                jb1 = (int)__ldg(cs + cx + 3 - nc0);
                cxb = wx - Lxf;
            }
        }
        if (lane == 0) s_cnt[warp] = 0;
        float a0 = kInf, a1 = kInf, a2 = kInf, a3 = kInf;
        widen_scan_row(P.wrapped, ja0, ja1, wx, cys, czs, 0, lowq_hi2, 0.f, self_j, a0, a1, a2, a3, nullptr, nullptr);
        widen_scan_row(P.wrapped, jb0, jb1, cxb, cys, czs, 0, lowq_hi2, 0.f, self_j, a0, a1, a2, a3, nullptr, nullptr);
        // 4th smallest over the warp: extract the minimum four times (non-negative floats order like uints)
        float fourth = kInf;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const unsigned m = __reduce_min_sync(kFullMask, __float_as_uint(a0));
            fourth = __uint_as_float(m);
            const unsigned holders = __ballot_sync(kFullMask, __float_as_uint(a0) == m);
            if (lane == __ffs(holders) - 1) {
                a0 = a1; a1 = a2; a2 = a3; a3 = kInf;
            }
        }
        const float thr = fminf(fourth + cst, thr_cap);
        __syncwarp();
        widen_scan_row(P.wrapped, ja0, ja1, wx, cys, czs, 1, lowq_hi2, thr, self_j, a0, a1, a2, a3, &s_cnt[warp], s_list[warp]);
        widen_scan_row(P.wrapped, jb0, jb1, cxb, cys, czs, 1, lowq_hi2, thr, self_j, a0, a1, a2, a3, &s_cnt[warp], s_list[warp]);
        __syncwarp();
        const int n = s_cnt[warp];
        // exact evaluation: lane k takes survivor k
        bool elig = false;
        double dist = 0.0;
        int idx = INT_MAX, j = -1;
        if (n <= 32 && lane < n) {
            j = s_list[warp][lane];
            double px, py, pz;
            RecTraits<double>::load(P.recs, (size_t)j, px, py, pz, idx);
            const double dx = min_image_1<double, EXACT>(px, rx, Lx, iLx);
            const double dy = min_image_1<double, EXACT>(py, ry, Ly, iLy);
            const double dz = min_image_1<double, EXACT>(pz, rz, Lz, iLz);
            const double s = sumsq3<double>(dx, dy, dz);
            if ((s > lowqsq) && (s <= selsq2)) {
                elig = true;
                const double ex = __dsub_rn(__dadd_rn(rx, dx), rx);
                const double ey = __dsub_rn(__dadd_rn(ry, dy), ry);
                const double ez = __dsub_rn(__dadd_rn(rz, dz), rz);
                dist = __dsqrt_rn(sumsq3<double>(ex, ey, ez));
            }
        }
        const unsigned emask = __ballot_sync(kFullMask, elig);
        const int nq = __popc(emask);
        if (n > 32 || (nq < 4 && !last2)) {
            if (lane == 0) {
                const uint32_t at = atomicAdd(P.counters + kCntLevel2, 1u);
                P.list2[at] = id | kFbNeedQ;
            }
            continue;
        }
        // rank among the eligible by (distance, atom index)
        int rank = 0;
        for (unsigned mm = emask; mm; mm &= mm - 1) {
            const int src = __ffs(mm) - 1;
            const double od = __shfl_sync(kFullMask, dist, src);
            const int oi = __shfl_sync(kFullMask, idx, src);
            if (key_less(od, oi, dist, idx)) ++rank;
        }
        const int n_found = min(nq, 4);
        Top4<double> top;
        top.reset();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const unsigned who = __ballot_sync(kFullMask, elig && rank == r);
            const int src = who ? __ffs(who) - 1 : 0;
            const int wj = __shfl_sync(kFullMask, j, src), wi = __shfl_sync(kFullMask, idx, src);
            if (r < n_found) {
                top.p[r] = wj;
                top.i[r] = wi;
            }
        }
        if (lane == 0) {
            LaneStats st;
            st.reset();
            finish_q<EXACT>(P, f, rx, ry, rz, Lx, Ly, Lz, iLx, iLy, iLz, top, n_found, out_index, st);
            if (P.stats) {
                atomicAdd(P.stats + (size_t)f * WOL_NSTATS + WOL_STAT_Q_SUM, st.q_sum);
                atomicAdd(P.stats + (size_t)f * WOL_NSTATS + WOL_STAT_Q_SUMSQ, st.q_sumsq);
                atomicAdd(P.stats + (size_t)f * WOL_NSTATS + WOL_STAT_N_CENTRES, 1.0);
            }
        }
    }
}

// adjacency image == minimum image needs 3 cell edges < L / 2
bool q3b_tpc_widen_supported(const Q3bParams &P) {
    return P.wrapped != nullptr && P.nc0 >= 7 && P.nc1 >= 7 && P.nc2 >= 7;
}

int q3b_tpc_widen_launch(const Q3bParams &P, cudaStream_t stream, bool exact) {
    const int grid = sm_count() * 12;
    if (exact) q3b_tpc_widen_kernel<true><<<grid, kWidenThreads, 0, stream>>>(P);
    else q3b_tpc_widen_kernel<false><<<grid, kWidenThreads, 0, stream>>>(P);
    add_launches(1);
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
    cudaError_t e = cudaFuncSetAttribute(q3b_tpc_kernel<EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(tpc)", e);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, q3b_tpc_kernel<EXACT>, kTpcThreads, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > P.total_tiles) grid = P.total_tiles;
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
