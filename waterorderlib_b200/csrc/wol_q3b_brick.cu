// K2, brick path (sm_100a): the dominant kernel for large frames, fp64 mode, every atom a centre.
//
// One persistent CTA per SM: 15 consumer warps + 1 producer warp.  The unit of work is a BRICK of cells
// (about 15 x 4 x 4) plus its one-cell halo.
//
//   producer   takes the next brick from a device counter, reads the cell starts at the ends of the
//              (y, z) rows of brick + halo (every row is one contiguous x-run of the cell-sorted arrays,
//              plus one cell from the other end of the box where the run wraps), and stages the rows' float
//              prefilter coordinates (`wrapped`, 16 B per atom) into one of three shared-memory stages with
//              1-D bulk async copies (cp.async.bulk -> mbarrier complete_tx), up to two bricks ahead of
//              the consumers.  Rows (or row ends) that come from the other side of the box are shifted by
//              the box edge in place once they have landed, so the sweep has no image logic at all; .w of a
//              staged atom is its position in the cell-sorted arrays (written by the cell build).  A brick
//              whose halo does not fit a stage is split along x on the spot.
//   consumers  warps take chunks of 32 consecutive centres of a brick from a shared counter; a warp that
//              finds none left moves on to the next brick, so no warp ever waits for another.
//     phase 1  float prefilter over the 9 stencil rows, all operands from shared memory (LDS.128 per
//              candidate, 3 FADD + 1 FMUL + 2 FFMA + compare); a survivor costs one predicated store of
//              (distance^2 | stage slot) -- nothing else is allowed inside this loop, because with 32 lanes some
//              lane has a survivor in nearly every iteration.
//     phase 1b dense pass over the ~9 survivors: min/max network for the four smallest float distances; a
//              survivor beyond the three-body cutoff is kept only while it can still be one of the four
//              nearest (4th smallest so far + rounding slack): ~5.5 remain.
//     phase 2  exact re-evaluation of those in the reference's fp64 operation order from the fp64 records
//              (gathered from L2, next one in flight): cutoff tests and neighbour counts are bit-exact by
//              construction; the UNIT vector of every kept neighbour goes to a per-thread shared column.
//     phase 3a three-body pairs flattened over the warp; cosine = dot of two unit vectors (3 DFMA).
//     phase 3b q from the four winners' unit vectors.
//
// "Certified" decisions.  The reference computes the cosine as dot / sqrt(n1 * n2) from vectors
// (r + d) - r, every operation rounded (waterlib.f90:694-698); the unit-vector cosine differs from it by
// at most eps_c = O(2^-50 (|r| + cutoff) / |d|) (derivation in q3b_brick_launch, a few 1e-12 for a 310 A
// box).  A histogram bin, the tetrahedral-window test, the order of the four nearest and the q bin are
// taken from the fast value only when it is farther than that bound from every decision boundary;
// otherwise the pair is re-evaluated in the reference's exact arithmetic (bk_exact_pair), or the centre's
// q is handed to the exact widened-search kernel.  Outputs are therefore bit-identical to the exact path;
// the slow paths fire for ~1e-8 of the angles (counted in counters[kCntSlowPair]).
#include "wol_q3b_brick.cuh"

namespace wol {

constexpr int kBkWarps = 15;                    // consumer warps (16 warps with the producer: 128 registers each;
                                                // a 17th warp would be charged as 20 -- warps are allocated in fours)
constexpr int kBkConsumers = kBkWarps * 32;
constexpr int kBkThreads = kBkConsumers + 32;   // + one producer warp
constexpr int kBkAtomCap = 1536;                // atoms of brick + halo per stage (slot fits 11 bits)
static_assert(kBkAtomCap <= 2048, "slot must fit 11 bits");

struct BkSmem {
    static constexpr int kAtomCap = kBkAtomCap;
    static constexpr int kProducerRegs = 124;      // registers the producer warp may use
    float4 loc[kBkStages][kBkAtomCap];
    double ent[kBkEntCap][3][kBkConsumers];
    unsigned char ent_k[kBkEntCap][kBkConsumers];  // which kept survivor the entry is (its list entry says where the record is)
    unsigned lj[kBkListCap + 1][kBkConsumers];   // survivor lists (+ one row that absorbs overflowing stores)
    unsigned short cs[kBkStages][kBkRowCap * kBkCsW];
    int cgj[kBkConsumers];                       // where the centre's fp64 record is
    int woff[kBkWarps][33];
    BkItem item[kBkStages];
    BkRow prow[kBkRowCap];
    double pbox[6];                              // producer scratch: box edges of the frame being staged and their reciprocals
    unsigned long long bar_full[kBkStages], bar_raw[kBkStages], bar_empty[kBkStages];
    unsigned char pair_ab[kBkMaxPairs + 4];
    __device__ __forceinline__ BkRow *prow_of(int) { return prow; }
    __device__ __forceinline__ double *pbox_of(int) { return pbox; }
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kBkConsumers) : "memory"); }

__device__ __forceinline__ void bk_flush_bins(unsigned *s_bins, unsigned long long *g_bins, int nbins, bool clear, int tid) {
    for (int i = tid; i < nbins; i += kBkConsumers) {
        const unsigned v = s_bins[i];
        if (v) atomicAdd(g_bins + i, (unsigned long long)v);
        if (clear) s_bins[i] = 0u;
    }
}

__global__ void __launch_bounds__(kBkThreads, 1) q3b_brick_kernel(const __grid_constant__ Q3bParams P, const __grid_constant__ BrickPlan B) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BkSmem &S = *reinterpret_cast<BkSmem *>(smem_raw);
    unsigned char *after = smem_raw + sizeof(BkSmem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do3 = P.do_3b != 0, doq = P.do_q != 0;
    const bool use_hist = do3 && P.ang_hist, use_qhist = doq && P.q_hist;
    const int tab_len = do3 ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
    double *s_tab = reinterpret_cast<double *>(after);
    unsigned *s_hist = reinterpret_cast<unsigned *>(after + sizeof(double) * tab_len);
    unsigned *s_qhist = s_hist + (use_hist ? P.nbins : 0);
    for (int i = tid; i < tab_len; i += kBkThreads) s_tab[i] = P.table[i];
    if (use_hist)
        for (int i = tid; i < P.nbins; i += kBkThreads) s_hist[i] = 0u;
    if (use_qhist)
        for (int i = tid; i < P.q_nbins; i += kBkThreads) s_qhist[i] = 0u;
    if (tid < kBkMaxPairs) {
        int b = 1;  // p = b (b - 1) / 2 + a, a < b
        while ((b + 1) * b / 2 <= tid) ++b;
        S.pair_ab[tid] = (unsigned char)((tid - b * (b - 1) / 2) | (b << 4));
    }
    if (tid == 0) {
        for (int s = 0; s < kBkStages; ++s) {
            mbar_init(&S.bar_full[s], 32);
            mbar_init(&S.bar_raw[s], 1);
            mbar_init(&S.bar_empty[s], kBkWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kBkWarps) {
        bk_producer<BkSmem, 1>(P, B, S, lane, 0);
        return;
    }

    const double *tab = s_tab;
    const int nbins = P.nbins;
    const double inv_width = (double)nbins / (P.hist_hi - P.hist_lo);
    const float hist_lo_f = (float)P.hist_lo, inv_width_f = (float)inv_width;
    const double tet_c_hi = do3 ? tab[nbins + 3] : 0.0, tet_c_lo = do3 ? tab[nbins + 4] : 0.0;
    // bins that hold the ends of the tetrahedral window (exact positions of the two cosines)
    const int tet_pos_hi = do3 ? angle_position(tet_c_hi, tab, nbins, hist_lo_f, inv_width_f) : -2;
    const int tet_pos_lo = do3 ? angle_position(tet_c_lo, tab, nbins, hist_lo_f, inv_width_f) : -2;
    const double low3sq = P.low3sq, high3sq = P.high3sq, lowqsq = P.lowqsq, highqsq = P.highqsq;
    const bool last1 = P.wq_max <= 1;
    const double selsq1 = last1 ? highqsq : fmin(highqsq, fmin(P.highq, P.rc1) * fmin(P.highq, P.rc1));
    const float pre_thr2 = P.pre_thr2, pre_thr3 = B.pre_thr3, pre_cst1 = B.pre_cst1, lowq_hi2 = P.lowq_hi2;
    const float kInf = __int_as_float(0x7f800000);
    const HistSpec qhs = hist_spec(0.0, 1.0, P.q_nbins);
    unsigned *const my_list = &S.lj[0][tid];

    LaneStats st;
    st.reset();
    int cur_f = -1;
    for (unsigned it = 0;; ++it) {
        const int s = (int)(it % kBkStages);
        mbar_wait(&S.bar_full[s], (it / kBkStages) & 1u);
        BkItem &I = S.item[s];
        if (I.done) break;
        const int f = I.frame;
        if (f != cur_f) {
            if (cur_f >= 0) {
                bk_flush_stats(P, cur_f, st);
                if ((use_hist || use_qhist) && P.hist_per_frame) {
                    consumer_bar();
                    if (use_hist) bk_flush_bins(s_hist, P.ang_hist + (size_t)cur_f * nbins, nbins, true, tid);
                    if (use_qhist) bk_flush_bins(s_qhist, P.q_hist + (size_t)cur_f * P.q_nbins, P.q_nbins, true, tid);
                    consumer_bar();
                }
            }
            cur_f = f;
        }
        const float4 *loc = S.loc[s];
        const unsigned short *cst = S.cs[s];
        const int nbx = I.nbx, rstride = I.nby + 2, n_centres = I.n_centres, n_chunks = I.n_chunks, n_crows = I.n_crows;
        for (;;) {
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&I.next, 1);
            chunk = __shfl_sync(kFullMask, chunk, 0);
            if (chunk >= n_chunks) break;
            const int ci = chunk * 32 + lane;
            const bool valid = ci < n_centres;
            int slot = 0, hx1 = 1, hrow = rstride + 1;
            if (valid) {
                int r = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int t = r + step;
                    if (t < n_crows && I.crow_off[t] <= ci) r = t;
                }
                slot = I.crow_slot[r] + (ci - I.crow_off[r]);
                hrow = I.crow_hrow[r];
                const unsigned short *row = cst + hrow * kBkCsW;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int t = hx1 + step;
                    if (t <= nbx && (int)row[t] <= slot) hx1 = t;
                }
            }
            const float4 me = loc[slot];
            const int gj = __float_as_int(me.w);  // the centre's place in the cell-sorted arrays = its id in the queues
            S.cgj[tid] = gj;
            // the centre's fp64 record is needed from phase 2 on: in flight during the sweep
            double rx = 0, ry = 0, rz = 0;
            int my_idx = 0;
            if (valid) bk_load_rec(P.recs, gj, rx, ry, rz, my_idx);

            // ---------------- phase 1: float prefilter over the 9 rows of the stencil ---------------------
            // A survivor is ONE predicated store (distance^2 with the slot in its low mantissa bits); the centre itself
            // passes (distance 0) and is dropped in phase 1b.
            int nl = 0;
            if (valid) {
                const unsigned short *row = cst + (hrow - rstride - 1) * kBkCsW + hx1 - 1;
#pragma unroll 1
                for (int r9 = 0; r9 < 9; ++r9) {
                    int j = row[0];
                    const int jend = row[3];
                    row += (r9 == 2 || r9 == 5) ? (rstride - 2) * kBkCsW : kBkCsW;
                    float4 w = loc[j];
                    // (left to itself the compiler unrolls by four with a remainder loop; the lanes of a warp have different
                    // trip counts, so every step of unrolling is lanes idling: by two measured best, -3.5 %)
#pragma unroll 2
                    while (j < jend) {
                        const float4 wn = loc[j + 1];  // a stage holds one spare entry
                        const float dx = w.x - me.x, dy = w.y - me.y, dz = w.z - me.z;
                        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (r2 <= pre_thr2) {
                            my_list[min(nl, kBkListCap) * kBkConsumers] = (__float_as_uint(r2) & ~kBkSlotMask) | (unsigned)j;
                            ++nl;
                        }
                        w = wn;
                        ++j;
                    }
                }
            }
            bool overflow = nl > kBkListCap;

            // ---------------- phase 1b: which survivors can matter -------------------------------------------
            // Four smallest distances^2 (truncated to 12 mantissa bits: the slack pre_cst1 covers that) among the
            // survivors certainly beyond lowCut.  A survivor is kept if it can be a three-body neighbour, or MARKED (bit 31)
            // if it can be one of the four nearest.  Two passes, so that what is kept does not depend on the order of the
            // atoms inside a cell (the cell build hands out slots with an atomic).  Kept entries become positions in the
            // cell-sorted arrays: the stage is not needed again.
            int nk = 0;
            if (valid && !overflow) {
                float a0 = kInf, a1 = kInf, a2 = kInf, a3 = kInf;
#pragma unroll 2
                for (int k = 0; k < nl; ++k) {
                    const unsigned e = my_list[k * kBkConsumers];
                    const float r2 = __uint_as_float(e & ~kBkSlotMask);
                    if ((int)(e & kBkSlotMask) != slot && r2 > lowq_hi2) {
                        float v = r2, m;
                        m = fminf(a0, v); v = fmaxf(a0, v); a0 = m;
                        m = fminf(a1, v); v = fmaxf(a1, v); a1 = m;
                        m = fminf(a2, v); v = fmaxf(a2, v); a2 = m;
                        a3 = fminf(a3, v);
                    }
                }
                const float thr_q = doq ? a3 + pre_cst1 : -1.f, thr_keep = fmaxf(pre_thr3, thr_q);
#pragma unroll 2
                for (int k = 0; k < nl; ++k) {
                    const unsigned e = my_list[k * kBkConsumers];
                    const int j = (int)(e & kBkSlotMask);
                    const float r2 = __uint_as_float(e & ~kBkSlotMask);
                    if (j != slot && r2 <= thr_keep) {
                        my_list[nk * kBkConsumers] = (unsigned)__float_as_int(loc[j].w) | (r2 <= thr_q ? 0x80000000u : 0u);
                        ++nk;
                    }
                }
            }

            // ---------------- phase 2: exact fp64 re-evaluation, unit vectors ---------------------------------
            Top4S top;
            top.reset();
            double rej_min = Ops<double>::inf();  // smallest distance^2 among the marked candidates that did not make the four
            int K3 = 0, Kb = 0, nq = 0;
            float bmax = 0.f;
            if (valid) bmax = __double2float_ru(fmax(fmax(fabs(rx), fabs(ry)), fabs(rz)));
            if (valid && !overflow) {
                const double Lx = I.L[0], Ly = I.L[1], Lz = I.L[2], iLx = I.iL[0], iLy = I.iL[1], iLz = I.iL[2];
                double nx = 0, ny = 0, nz = 0;
                int nidx = 0;
                unsigned ne = 0;
                if (nk > 0) {
                    ne = my_list[0];
                    bk_load_rec(P.recs, (int)(ne & 0x7fffffffu), nx, ny, nz, nidx);
                }
                for (int k = 0; k < nk; ++k) {
                    const bool marked = (ne >> 31) != 0u;
                    const double px = nx, py = ny, pz = nz;
                    if (k + 1 < nk) {  // next survivor's record is in flight while this one is evaluated
                        ne = my_list[(k + 1) * kBkConsumers];
                        bk_load_rec(P.recs, (int)(ne & 0x7fffffffu), nx, ny, nz, nidx);
                    }
                    const double dx = min_image_1<double, false>(px, rx, Lx, iLx);
                    const double dy = min_image_1<double, false>(py, ry, Ly, iLy);
                    const double dz = min_image_1<double, false>(pz, rz, Lz, iLz);
                    const double sq = sumsq3<double>(dx, dy, dz);
                    const bool in3 = do3 && (sq > low3sq) && (sq <= high3sq);
                    const bool inq = doq && (sq > lowqsq) && (sq <= selsq1);
                    nq += inq ? 1 : 0;
                    // an unmarked candidate is farther than four others by more than the float arithmetic can hide
                    const bool want_q = inq && marked;
                    if (in3 || want_q) {
                        const int e = in3 ? K3++ : kBkEntCap - 1 - Kb++;
                        if (K3 + Kb > kBkEntCap || sq < B.floor2) {
                            overflow = true;
                        } else {
                            const double rs = rsqrt(sq);
                            // column rotated by 3 e inside the warp's 32: the lanes of phase 3a that work on one centre's
                            // pairs read different entries of it, and a plain [e][k][t] layout puts those in one bank
                            const int te = (tid & ~31) | ((tid + 3 * e) & 31);
                            S.ent[e][0][te] = dx * rs;
                            S.ent[e][1][te] = dy * rs;
                            S.ent[e][2][te] = dz * rs;
                            S.ent_k[e][tid] = (unsigned char)k;
                            if (want_q) {
                                if (sq < top.d[3]) {
                                    rej_min = fmin(rej_min, top.d[3]);
                                    top.insert(sq, e);
                                } else {
                                    rej_min = fmin(rej_min, sq);
                                }
                            }
                        }
                    }
                }
            }
            // one bound per warp: the lanes evaluate one another's pairs in phase 3a
            const float bw = __uint_as_float(__reduce_max_sync(kFullMask, __float_as_uint(bmax)));
            const double eps_c = fma(B.eps_a, (double)bw, B.eps_b);
            bool q_go = valid && doq && !overflow;
            const bool b3_go = valid && do3 && !overflow;
            const size_t out_index = (size_t)f * P.n_pos + my_idx;
            if (valid && overflow) {
                const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
                P.fb_list[at] = (uint32_t)gj | (do3 ? kFbNeed3b : 0u) | (doq ? kFbNeedQ : 0u);
                atomicAdd(P.counters + kCntOverflow, 1u);
            }
            if (q_go) {
                bool requeue = nq < 4 && !last1;  // fewer than four inside the radius the stencil guarantees
                if (!requeue) {
                    // the order of the four nearest must survive the distance between this arithmetic and the reference's
                    const int nf = min(nq, 4);
                    const double band = 4.0 * eps_c;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (k + 1 < nf && !(top.d[k + 1] - top.d[k] > band * top.d[k + 1])) requeue = true;
                    if (rej_min < Ops<double>::inf() && !(rej_min - top.d[3] > band * rej_min)) requeue = true;
                }
                if (requeue) {
                    bk_push_q(P, (uint32_t)gj);
                    q_go = false;
                }
            }

            // ---------------- phase 3a: three-body pairs, flattened over the warp ---------------------------
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
                    int t = 0, base = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const int v = woff[t + step];
                        if (v <= w) {
                            t += step;
                            base = v;
                        }
                    }
                    const int ab = S.pair_ab[w - base];
                    const int col = warp * 32 + t, ea = ab & 15, eb = ab >> 4;
                    const int ca = warp * 32 + ((t + 3 * ea) & 31), cb = warp * 32 + ((t + 3 * eb) & 31);
                    double c = fma(S.ent[ea][0][ca], S.ent[eb][0][cb],
                                   fma(S.ent[ea][1][ca], S.ent[eb][1][cb], S.ent[ea][2][ca] * S.ent[eb][2][cb]));
                    // Bin of the fast value with its certificate: the reference's (clamped) cosine lies within eps_c of c, so
                    // the bin is settled when [c - eps_c, c + eps_c] sits inside one bin's cosine interval and away from -1
                    // (whose angle the reference turns into -180 degrees).  No clamp: a |c| beyond 1 fails the certificate.
                    const double chi = c + eps_c, clo = c - eps_c;
                    int pos = bk_seed_position(c, nbins, hist_lo_f, inv_width_f);
                    bool sure = clo > -1.0 && chi <= tab[pos] && clo > tab[pos + 1];
                    if (pos == tet_pos_hi || pos == tet_pos_lo)  // the tetrahedral window's ends fall inside these two bins
                        if ((chi >= tet_c_hi && clo <= tet_c_hi) || (chi >= tet_c_lo && clo <= tet_c_lo)) sure = false;
                    if (!sure) {
                        // seeded one bin off (2 % of the pairs), outside the histogram range, or really too close to call
                        c = fmin(1.0, fmax(-1.0, c));
                        pos = angle_position(c, tab, nbins, hist_lo_f, inv_width_f);
                        sure = clo > -1.0;
                        if (pos >= 0 && !(chi <= tab[pos])) sure = false;
                        if (pos < nbins && !(clo > tab[pos + 1])) sure = false;
                        if ((chi >= tet_c_hi && clo <= tet_c_hi) || (chi >= tet_c_lo && clo <= tet_c_lo)) sure = false;
                        if (!sure) {
                            c = bk_exact_pair(reinterpret_cast<const RecD *>(P.recs), S.cgj[col],
                                              (int)(S.lj[S.ent_k[ea][col]][col] & 0x7fffffffu), (int)(S.lj[S.ent_k[eb][col]][col] & 0x7fffffffu), I.L, I.iL);
                            pos = bk_exact_position(c, tab, nbins, hist_lo_f, inv_width_f);
                            atomicAdd(P.counters + kCntSlowPair, 1u);
                        }
                    }
                    if (c != -1.0 && c <= tet_c_hi && c >= tet_c_lo) {
                        st.tet_count += 1u;
                        st.tet_cos += c;
                        st.tet_cossq += c * c;
                    }
                    st.n_angles += 1u;
                    if (pos >= 0 && pos < nbins) {
                        if (use_hist) atomicAdd(s_hist + pos, 1u);
                    }
                }
                __syncwarp();
                if (b3_go) {
                    if (P.n3) P.n3[out_index] = K3;
                    st.n_neigh += (unsigned)K3;
                }
            }

            // ---------------- phase 3b: q from the four winners' unit vectors ------------------------------
            if (q_go) {
                const int nf = min(nq, 4);
                double ux[4], uy[4], uz[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int e = top.p[k];
                    const int te = (tid & ~31) | ((tid + 3 * e) & 31);
                    ux[k] = S.ent[e][0][te];
                    uy[k] = S.ent[e][1][te];
                    uz[k] = S.ent[e][2][te];
                }
                double acc = 0.0;
                int n_real = 0;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = a + 1; b < 4; ++b)
                        if (b < nf) {
                            double c = fma(ux[a], ux[b], fma(uy[a], uy[b], uz[a] * uz[b]));
                            c = fmin(1.0, fmax(-1.0, c));
                            const double u = c + (1.0 / 3.0);
                            acc = fma(u, u, acc);
                            ++n_real;
                        }
                for (int k = n_real; k < 6; ++k) {
                    const double u = -1.0 + (1.0 / 3.0);
                    acc += u * u;
                }
                const double qv = (nf == 0) ? 0.0 : 1.0 - (3.0 / 8.0) * acc;
                int bin = -1;
                bool sure = true;
                if (use_qhist) {
                    // q differs from the reference's by at most 6 eps_c + rounding; its bin must not depend on that
                    const double eps_q = 8.0 * eps_c;
                    bin = hist_bin(qhs, qv);
                    if (nf > 0) {
                        if (bin < 0) sure = (qv < -eps_q) || (qv > 1.0 + eps_q);
                        else sure = (qv - eps_q >= hist_edge(qhs, bin)) && (qv + eps_q < hist_edge(qhs, bin + 1));
                    }
                }
                if (!sure) {
                    bk_push_q(P, (uint32_t)gj);
                } else {
                    if (P.q) reinterpret_cast<double *>(P.q)[out_index] = qv;
                    if (P.nn_idx) {
                        const RecD *recs = reinterpret_cast<const RecD *>(P.recs);
                        int4 o;
                        o.x = (nf > 0) ? __ldg(&recs[my_list[S.ent_k[top.p[0]][tid] * kBkConsumers] & 0x7fffffffu].idx) : -1;
                        o.y = (nf > 1) ? __ldg(&recs[my_list[S.ent_k[top.p[1]][tid] * kBkConsumers] & 0x7fffffffu].idx) : -1;
                        o.z = (nf > 2) ? __ldg(&recs[my_list[S.ent_k[top.p[2]][tid] * kBkConsumers] & 0x7fffffffu].idx) : -1;
                        o.w = (nf > 3) ? __ldg(&recs[my_list[S.ent_k[top.p[3]][tid] * kBkConsumers] & 0x7fffffffu].idx) : -1;
                        reinterpret_cast<int4 *>(P.nn_idx)[out_index] = o;
                    }
                    if (bin >= 0) atomicAdd(s_qhist + bin, 1u);
                    st.q_sum += qv;
                    st.q_sumsq += qv * qv;
                    st.n_centres += 1u;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.bar_empty[s]);
    }
    if (cur_f >= 0) bk_flush_stats(P, cur_f, st);
    if (use_hist || use_qhist) {
        consumer_bar();
        if (cur_f >= 0) {
            const size_t rowi = (size_t)(P.hist_per_frame ? cur_f : 0);
            if (use_hist) bk_flush_bins(s_hist, P.ang_hist + rowi * nbins, nbins, false, tid);
            if (use_qhist) bk_flush_bins(s_qhist, P.q_hist + rowi * P.q_nbins, P.q_nbins, false, tid);
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------

static size_t brick_smem_bytes(const Q3bParams &P) {
    const int tab_len = P.do_3b ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
    size_t smem = sizeof(BkSmem) + sizeof(double) * tab_len;
    if (P.do_3b && P.ang_hist) smem += sizeof(unsigned) * P.nbins;
    if (P.do_q && P.q_hist) smem += sizeof(unsigned) * P.q_nbins;
    return smem;
}

bool q3b_brick_supported(const Q3bParams &P, bool exact) {
    if (P.centres != nullptr || P.n_valid != nullptr || P.wrapped == nullptr || exact) return false;
    if (P.nc0 < 4 || P.nc1 < 4 || P.nc2 < 4) return false;
    if (brick_smem_bytes(P) > 227u * 1024u) return false;
    const char *env = getenv("WOL_BRICK");  // test switch: 1 = also for small batches, 0 = never, 3 = also, and the
                                            // warp-specialised kernel where it applies (wol_q3b_brick_ws.cu)
    if (env && env[0] == '0') return false;
    if (env && (env[0] == '1' || env[0] == '3')) return true;
    int nb[3];
    brick_dims(P, nb, kBkConsumers, kBkAtomCap);
    return (long long)nb[0] * nb[1] * nb[2] * P.n_frames >= 2LL * sm_count();
}

int q3b_brick_launch(const Q3bParams &P, double box_max, cudaStream_t stream) {
    BrickPlan B;
    int nb[3];
    brick_dims(P, nb, kBkConsumers, kBkAtomCap);
    B.nb0 = nb[0];
    B.nb1 = nb[1];
    B.nb2 = nb[2];
    B.bricks_per_frame = nb[0] * nb[1] * nb[2];
    const long long total = (long long)B.bricks_per_frame * P.n_frames;
    if (total >= (1LL << 31)) return set_error(WOL_ERR_RANGE, "too many bricks");
    B.total = (unsigned)total;
    B.m_bpf = bk_div_magic((unsigned)B.bricks_per_frame);
    B.m_nb0 = bk_div_magic((unsigned)nb[0]);
    B.m_nb1 = bk_div_magic((unsigned)nb[1]);
    B.m_nb2 = bk_div_magic((unsigned)nb[2]);
    brick_plan_bounds(P, box_max, B);
    const size_t smem = brick_smem_bytes(P);
    cudaError_t e = cudaFuncSetAttribute(q3b_brick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(brick)", e);
    long long grid = sm_count();
    if (grid > total) grid = total;
    if (grid > 0) {
        q3b_brick_kernel<<<(unsigned)grid, kBkThreads, smem, stream>>>(P, B);
        add_launches(1);
        e = cudaGetLastError();
        if (e != cudaSuccess) return set_cuda_error("brick kernel launch", e);
        if (getenv("WOL_DEBUG_SYNC")) {  // debugging aid: surface device-side faults at the launch that caused them
            e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) return set_cuda_error("brick kernel", e);
        }
    }
    return WOL_OK;
}

}  // namespace wol