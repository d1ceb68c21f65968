// K2, FP32 arithmetic mode: one THREAD per centre, everything in float on the box-wrapped coordinates
// (sm_100a).  Same decomposition as the fp64 kernel (wol_q3b_tpc.cu) without its exact re-evaluation: the
// float distance of the sweep IS the distance of this mode, so a candidate costs one 16-byte load, 3 FADD,
// 1 FMUL, 2 FFMA and the cutoff compares, neighbours go straight to the shared-memory list and the register
// top-4, and the pair phase uses rsqrt + acosf.  North-star tolerance of the mode: 1e-4 on q and cosines;
// neighbour selection and bin membership may differ from the fp64 reference where two values agree to ~1e-7.
//
// Reference semantics kept: cutoff test low^2 < r^2 <= high^2 (fortran/waterlib.f90:737,855), selection by
// (distance, atom index), CosAngle3's 0-degree return for coincident positions and its -180 for an exactly
// antiparallel pair (dropped from the histogram, :699-702), padding and q formula (water_properties.py:379-388).
#include "wol_q3b_f32.cuh"

namespace wol {

constexpr int kT32Threads = 256;
constexpr int kT32EntCap = 12;  // candidates inside the sweep radius per centre (three-body neighbours are a subset)
constexpr int kT32MaxPairs = kT32EntCap * (kT32EntCap - 1) / 2;

struct T32Smem {
    float4 ent[kT32EntCap][kT32Threads];  // (dx, dy, dz, r^2)
    int eidx[kT32EntCap][kT32Threads];    // atom index of the entry
    int woff[kT32Threads / 32][33];
    unsigned char pair_ab[kT32MaxPairs + 2];
};

__global__ void __launch_bounds__(kT32Threads, 3) q3b_tpc32_kernel(const __grid_constant__ Q3bParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T32Smem &S = *reinterpret_cast<T32Smem *>(smem_raw);
    unsigned *s_hist = reinterpret_cast<unsigned *>(smem_raw + sizeof(T32Smem));
    const bool smem_hist = P.do_3b && P.ang_hist && P.nbins <= kMaxSmemBins;
    const bool smem_qhist = P.do_q && P.q_hist && P.q_nbins <= kMaxSmemBins;
    unsigned *s_qhist = s_hist + (smem_hist ? P.nbins : 0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (smem_hist)
        for (int i = tid; i < P.nbins; i += kT32Threads) s_hist[i] = 0u;
    if (smem_qhist)
        for (int i = tid; i < P.q_nbins; i += kT32Threads) s_qhist[i] = 0u;
    if (tid < kT32MaxPairs) {  // p = b (b - 1) / 2 + a, a < b
        int b = 1;
        while ((b + 1) * b / 2 <= tid) ++b;
        S.pair_ab[tid] = (unsigned char)((tid - b * (b - 1) / 2) | (b << 4));
    }
    __syncthreads();
    const bool do3 = P.do_3b != 0, doq = P.do_q != 0;
    const float low3sq = (float)P.low3sq, high3sq = (float)P.high3sq, lowqsq = (float)P.lowqsq;
    const bool last1 = P.wq_max <= 1;
    const float selsq1 = (float)(last1 ? P.highqsq : fmin(P.highqsq, fmin(P.highq, P.rc1) * fmin(P.highq, P.rc1)));
    const float reach2 = fmaxf(do3 ? high3sq : 0.f, doq ? selsq1 : 0.f);
    const float hist_lo = (float)P.hist_lo, hist_hi = (float)P.hist_hi;
    const float inv_width = (float)((double)P.nbins / (P.hist_hi - P.hist_lo));
    const HistSpec qspec = hist_spec(0.0, 1.0, P.q_nbins);
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
    const float4 *__restrict__ wr = P.wrapped;

    LaneStats st;
    st.reset();
    const long long n_chunks = (P.total_tiles + P.chunk_tiles - 1) / P.chunk_tiles;
    int cur_f = -1;
    float Lxf = 1, Lyf = 1, Lzf = 1;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x)
    for (long long tile = chunk * P.chunk_tiles, t_end = min(P.total_tiles, tile + P.chunk_tiles); tile < t_end; ++tile) {
        const int f = (int)(tile / P.tiles_per_frame);
        const int m = (int)(tile - (long long)f * P.tiles_per_frame) * kT32Threads + tid;
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
            Lxf = (float)P.box[(size_t)f * 3 + 0];
            Lyf = (float)P.box[(size_t)f * 3 + 1];
            Lzf = (float)P.box[(size_t)f * 3 + 2];
        }
        const bool valid = m < P.n_centres && (P.n_valid == nullptr || m < __ldg(P.n_valid + f));
        float wx = 0, wy = 0, wz = 0;
        int cx = 0, cy = 0, cz = 0, self_j = -1;
        size_t out_index = 0;
        uint32_t fb_id = 0;
        if (valid) {
            if (P.centres == nullptr) {
                const size_t j = (size_t)f * P.n_pos + m;
                const float4 w = __ldg(wr + j);
                const uint32_t cp = __ldg(P.cellpack + j);
                wx = w.x; wy = w.y; wz = w.z;
                cx = cp & 1023; cy = (cp >> 10) & 1023; cz = (cp >> 20) & 1023;
                self_j = (int)j;
                out_index = (size_t)f * P.n_pos + __float_as_int(w.w);
                fb_id = (uint32_t)j;
            } else {
                float rx, ry, rz;
                load_centre<float>(P, f, m, rx, ry, rz);
                const double Lx = P.box[(size_t)f * 3 + 0], Ly = P.box[(size_t)f * 3 + 1], Lz = P.box[(size_t)f * 3 + 2];
                const double iLx = __ddiv_rn(1.0, Lx), iLy = __ddiv_rn(1.0, Ly), iLz = __ddiv_rn(1.0, Lz);
                cx = cell_coord((double)rx, iLx, nc0);
                cy = cell_coord((double)ry, iLy, nc1);
                cz = cell_coord((double)rz, iLz, nc2);
                wx = wrapped_coord((double)rx, Lx, iLx);
                wy = wrapped_coord((double)ry, Ly, iLy);
                wz = wrapped_coord((double)rz, Lz, iLz);
                out_index = (size_t)f * P.n_centres + m;
                fb_id = (uint32_t)out_index;
            }
        }

        // ---------------- sweep: 9 rows of the stencil, final distances --------------------------
        // Two steps, like the fp64 kernel: the sweep only APPENDS what lies inside the reach (the append is a few
        // instructions, so the divergent sweep loop stays cheap); classification, the three-body compaction and
        // the top-4 insertion run afterwards in a dense per-lane loop over the ~6 entries.
        int nl = 0;
        auto visit = [&](int j, float4 w, float sx, float sy, float sz) {
            const float dx = w.x - sx, dy = w.y - sy, dz = w.z - sz;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (r2 <= reach2 && j != self_j) {
                if (nl < kT32EntCap) {
                    S.ent[nl][tid] = make_float4(dx, dy, dz, r2);
                    S.eidx[nl][tid] = __float_as_int(w.w);
                }
                ++nl;
            }
        };
        if (valid) {
            const size_t cell_base = (size_t)f * nc0 * nc1 * nc2;
            const int xa0 = max(cx - 1, 0), xa1 = min(cx + 1, nc0 - 1) + 1;
            int xb0 = 0, xb1 = 0;
            float sxb = 0.f;
            if (cx == 0) {
                xb0 = nc0 - 1; xb1 = nc0; sxb = Lxf;
            } else if (cx == nc0 - 1) {
                xb0 = 0; xb1 = 1; sxb = -Lxf;
            }
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
#pragma unroll
                for (int row = 0; row < 3; ++row) {
                    const int j1 = rj1[row];
                    for (int j = rj0[row]; j < j1; j += 2) {
                        const bool two = j + 1 < j1;
                        const float4 w0 = __ldg(wr + j);
                        const float4 w1 = __ldg(wr + (two ? j + 1 : j));
                        visit(j, w0, wx, rcy[row], czs);
                        if (two) visit(j + 1, w1, wx, rcy[row], czs);
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
                    for (int j = (int)__ldg(cs + xb0); j < j1; ++j) visit(j, __ldg(wr + j), cxs, cys, czs);
                }
            }
        }
        const bool overflow = nl > kT32EntCap;
        Top4F top;
        top.reset();
        int K3 = 0, nq = 0;
        if (valid && !overflow) {
            for (int k = 0; k < nl; ++k) {
                const float4 v = S.ent[k][tid];
                if (doq && v.w > lowqsq && v.w <= selsq1) {
                    ++nq;
                    top.insert(v.w, S.eidx[k][tid], v.x, v.y, v.z);
                }
                if (do3 && v.w > low3sq && v.w <= high3sq) {
                    if (K3 != k) S.ent[K3][tid] = v;  // compact the three-body neighbours to the front (K3 <= k)
                    ++K3;
                }
            }
        }
        bool q_go = valid && doq && !overflow;
        const bool b3_go = valid && do3 && !overflow;
        if (valid && overflow) {
            const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
            P.fb_list[at] = fb_id | (do3 ? kFbNeed3b : 0u) | (doq ? kFbNeedQ : 0u);
            atomicAdd(P.counters + kCntOverflow, 1u);
        }
        if (q_go && nq < 4 && !last1) {
            const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
            P.fb_list[at] = fb_id | kFbNeedQ;
            atomicAdd(P.counters + kCntWidened, 1u);
            q_go = false;
        }

        // ---------------- three-body pairs, flattened over the warp -------------------------------
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
                const int ab = S.pair_ab[w - woff[t]];
                const int col = warp * 32 + t;
                const float4 va = S.ent[ab & 15][col], vb = S.ent[ab >> 4][col];
                float th;
                bool binned = true;
                if (va.w == 0.f || vb.w == 0.f) {
                    th = 0.f;  // coincident positions: CosAngle3 returns 0
                } else {
                    const float c = cos32(va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w);
                    th = acosf(c) * 57.29577951308232f;
                    if (c == -1.f) binned = false;  // the reference returns -180 here: outside every range starting at 0
                    else if (th >= 100.f && th <= 120.f) {
                        st.tet_count += 1u;
                        st.tet_cos += (double)c;
                        st.tet_cossq += (double)c * (double)c;
                    }
                }
                st.n_angles += 1u;
                if (binned && th >= hist_lo && th <= hist_hi) {
                    const int pos = min((int)((th - hist_lo) * inv_width), P.nbins - 1);
                    if (smem_hist) atomicAdd(s_hist + pos, 1u);
                    else if (P.ang_hist) atomicAdd(P.ang_hist + (size_t)(P.hist_per_frame ? f : 0) * P.nbins + pos, 1ull);
                }
            }
            __syncwarp();
            if (b3_go) {
                if (P.n3) P.n3[out_index] = K3;
                st.n_neigh += (unsigned)K3;
            }
        }

        // ---------------- q from the four winners ----------------------------------------------------
        if (q_go) finish_q32(P, f, top, min(nq, 4), out_index, st, smem_qhist ? s_qhist : nullptr, qspec);
    }
    if (cur_f >= 0) flush_stats(P, cur_f, st);
    if (smem_hist || smem_qhist) {
        __syncthreads();
        if (cur_f >= 0) flush_hist(P, smem_hist ? s_hist : nullptr, smem_qhist ? s_qhist : nullptr, cur_f, false);
    }
}

int q3b_tpc32_launch(const Q3bParams &P0, cudaStream_t stream) {
    Q3bParams P = P0;
    P.tiles_per_frame = (P.n_centres + kT32Threads - 1) / kT32Threads;
    P.total_tiles = (long long)P.tiles_per_frame * P.n_frames;
    P.chunk_tiles = 1;
    size_t smem = sizeof(T32Smem);
    if (P.do_3b && P.ang_hist && P.nbins <= kMaxSmemBins) smem += sizeof(unsigned) * P.nbins;
    if (P.do_q && P.q_hist && P.q_nbins <= kMaxSmemBins) smem += sizeof(unsigned) * P.q_nbins;
    cudaError_t e = cudaFuncSetAttribute(q3b_tpc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(tpc32)", e);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, q3b_tpc32_kernel, kT32Threads, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > P.total_tiles) grid = P.total_tiles;
    if (grid > 0) {
        q3b_tpc32_kernel<<<(unsigned)grid, kT32Threads, smem, stream>>>(P);
        add_launches(1);
    }
    return WOL_OK;
}

}  // namespace wol
