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
#pragma unroll 1
            for (int row = 0; row < 9; ++row) {
                const int dz = row / 3 - 1, dy = row - (row / 3) * 3 - 1;
                int y = cy + dy, z = cz + dz;
                float cys = wy, czs = wz;
                if (y < 0) { y += nc1; cys += Lyf; } else if (y >= nc1) { y -= nc1; cys -= Lyf; }
                if (z < 0) { z += nc2; czs += Lzf; } else if (z >= nc2) { z -= nc2; czs -= Lzf; }
                const uint32_t *cs = P.cell_start + cell_base + ((size_t)z * nc1 + y) * nc0;
#pragma unroll 1
                for (int piece = 0; piece < 2; ++piece) {
                    int j0, j1;
                    float cxs;
                    if (piece == 0) {
                        j0 = (int)__ldg(cs + xa0);
                        j1 = (int)__ldg(cs + xa1);
                        cxs = wx;
                    } else {
                        if (xb1 == 0) break;
                        j0 = (int)__ldg(cs + xb0);
                        j1 = (int)__ldg(cs + xb1);
                        cxs = wx + sxb;
                    }
                    for (int j = j0; j < j1; ++j) {
                        const float4 w = __ldg(P.wrapped + j);
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
            for (int k = 0; k < nl; ++k) {
                const int j = S.lj[k][tid];
                double px, py, pz;
                int idx;
                RecTraits<double>::load(P.recs, (size_t)j, px, py, pz, idx);
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
        if (q_go) {
            const int n_found = min(nq, 4);
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
                    // tetraCosAng reimages the already reimaged position again (waterlib.f90:880-883)
                    const double d2x = __dsub_rn(ex, __dmul_rn(Lx, anint_exact<double>(__dmul_rn(ex, iLx))));
                    const double d2y = __dsub_rn(ey, __dmul_rn(Ly, anint_exact<double>(__dmul_rn(ey, iLy))));
                    const double d2z = __dsub_rn(ez, __dmul_rn(Lz, anint_exact<double>(__dmul_rn(ez, iLz))));
                    vx[k] = __dsub_rn(__dadd_rn(rx, d2x), rx);
                    vy[k] = __dsub_rn(__dadd_rn(ry, d2y), ry);
                    vz[k] = __dsub_rn(__dadd_rn(rz, d2z), rz);
                    vn[k] = sumsq3<double>(vx[k], vy[k], vz[k]);
                }
            }
            // real angles in triu order, then the 180-degree padding (cos = -1) of
            // water_properties.py:379-384, summed left to right like np.sum
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
                if (b >= 0) atomicAdd(P.q_hist + (size_t)(P.hist_per_frame ? f : 0) * P.q_nbins + b, 1ull);
            }
            st.q_sum += qv;
            st.q_sumsq += qv * qv;
            st.n_centres += 1u;
        }
    }
    if (cur_f >= 0) flush_stats(P, cur_f, st);
    if (smem_hist) {
        __syncthreads();
        if (cur_f >= 0) flush_hist(P, s_hist, cur_f, false);
    }
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
