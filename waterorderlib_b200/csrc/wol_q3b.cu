// K2: fused neighbour-cell sweep -> neighbour list -> {three-body angle histogram, 4-nearest selection
// -> tetrahedral q}, for a batch of frames (sm_100a).
//
// Work decomposition: a GROUP of G lanes (G = 8 in the fast pass, a whole warp in the large-capacity
// pass) owns one centre.
//   phase 1  the lanes walk the flattened candidate list of the centre's cell stencil (contiguous
//            x-runs of the cell-sorted records), evaluate the reference's minimum-image arithmetic
//            exactly and append accepted candidates to the group's shared-memory list with a
//            ballot/popc prefix (three-body neighbours from the front, q-only candidates from the back);
//   phase 2  per list entry: the reimaged difference vector the reference's angle code sees, its
//            squared norm, and a per-lane register top-4 by (distance, atom index);
//   phase 3a the K(K-1)/2 neighbour pairs are spread over the lanes: clamped cosine in reference
//            operation order, bin by threshold-table compare (no device acos), shared-memory histogram;
//   phase 3b the group merges its lanes' top-4 with shuffles, and six lanes evaluate the pair cosines
//            of the four winners -> q.
// Centres whose 4th neighbour is not certainly inside the 27-cell stencil, or whose list overflows the
// fast capacity, are queued and redone by the same code instantiated with G = 32, a 1024-entry list
// and a stencil that widens until the search is provably complete.
//
// Reference arithmetic being reproduced (relative to /root/reference):
//   cutoff test            fortran/waterlib.f90:733-741, :851-859
//   reimage                fortran/waterlib.f90:43-45
//   tetraCosAng/CosAngle3  fortran/waterlib.f90:878-893, :690-702
//   4-NN selection         structureLibs/water_properties.py:372-374
//   q and padding          structureLibs/water_properties.py:379-388
//   angle histogram        structureLibs/water_properties.py:328
#include <math.h>
#include <stdlib.h>

#include "wol_q3b_common.cuh"

namespace wol {

// ------------------------------------------------------------------------------------------------
// One centre, one group.  Every lane of the WARP executes this function in lock step (the loops are
// bounded with warp votes); lanes of groups without work pass valid = false.

template <typename T, int G, int CAP, int MAXSEG, bool EXACT, bool LISTMODE>
struct GroupWorker {
    // per-group shared memory
    struct Smem {
        Vec4<T> ent[CAP];
        int eidx[CAP];
        int segj0[MAXSEG];
        int segpref[MAXSEG + 1];
        Vec4<T> win[4];
    };

    const Q3bParams &P;
    Smem &S;
    const int gl;          // lane inside the group
    const unsigned gshift;  // first warp lane of the group
    const unsigned gmask;

    __device__ GroupWorker(const Q3bParams &p, Smem &s)
        : P(p), S(s), gl((threadIdx.x & 31) & (G - 1)), gshift((threadIdx.x & 31) & ~(G - 1)),
          gmask(G == 32 ? 0xffffffffu : ((1u << G) - 1u)) {}

    __device__ __forceinline__ unsigned group_ballot(bool pred) const {
        return (__ballot_sync(kFullMask, pred) >> gshift) & gmask;
    }
    __device__ __forceinline__ int group_sum(int v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o, G);
        return v;
    }

    // Builds the group's segment table for half-width w; returns the candidate total.
    __device__ int build_segments(bool valid, int f, int cx, int cy, int cz, int w) {
        const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
        const int span = 2 * w + 1;
        const bool fullx = span >= nc0, fully = span >= nc1, fullz = span >= nc2;
        const int cntx = fullx ? nc0 : span, cnty = fully ? nc1 : span, cntz = fullz ? nc2 : span;
        const int xs = fullx ? 0 : (cx - w + nc0) % nc0;
        const int ys = fully ? 0 : (cy - w + nc1) % nc1;
        const int zs = fullz ? 0 : (cz - w + nc2) % nc2;
        const int npieces = (xs + cntx > nc0) ? 2 : 1;
        const int nseg = valid ? cntz * cnty * npieces : 0;
        const size_t cell_base = (size_t)f * nc0 * nc1 * nc2;
        int carry = 0;
        const int nseg_max = __reduce_max_sync(kFullMask, nseg);
        for (int base = 0; base < nseg_max; base += G) {
            const int seg = base + gl;
            int len = 0;
            if (seg < nseg) {
                const int piece = seg % npieces;
                const int rest = seg / npieces;
                const int iy = rest % cnty, iz = rest / cnty;
                int y = ys + iy;
                if (y >= nc1) y -= nc1;
                int z = zs + iz;
                if (z >= nc2) z -= nc2;
                int x0, x1;
                if (piece == 0) {
                    x0 = xs;
                    x1 = min(xs + cntx, nc0);
                } else {
                    x0 = 0;
                    x1 = xs + cntx - nc0;
                }
                const size_t row = cell_base + ((size_t)z * nc1 + y) * nc0;
                const int j0 = (int)__ldg(P.cell_start + row + x0);
                const int j1 = (int)__ldg(P.cell_start + row + x1);
                S.segj0[seg] = j0;
                len = j1 - j0;
            }
            // inclusive scan of len inside the group
            int inc = len;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                const int n = __shfl_up_sync(kFullMask, inc, o, G);
                if (gl >= o) inc += n;
            }
            if (seg < nseg) S.segpref[seg + 1] = carry + inc;
            carry += __shfl_sync(kFullMask, inc, G - 1, G);
        }
        if (gl == 0) S.segpref[0] = 0;
        __syncwarp();
        return valid ? carry : 0;
    }

    // Everything for one centre.  flags: bit0 = do three-body, bit1 = do q.  w_start: first stencil
    // half-width of the q search (the three-body sweep always uses half-width 1).
    __device__ void run(bool valid, int f, T rx, T ry, T rz, int cx, int cy, int cz, size_t out_index, bool do3,
                        bool doq, int w_start, uint32_t fb_id, unsigned *s_hist, const double *tab, LaneStats &st) {
        const T Lx = (T)P.box[(size_t)f * 3 + 0], Ly = (T)P.box[(size_t)f * 3 + 1], Lz = (T)P.box[(size_t)f * 3 + 2];
        const T iLx = Ops<T>::div((T)1, Lx), iLy = Ops<T>::div((T)1, Ly), iLz = Ops<T>::div((T)1, Lz);
        const T low3sq = (T)P.low3sq, high3sq = (T)P.high3sq, lowqsq = (T)P.lowqsq, highqsq = (T)P.highqsq;
        const double inv_width = (double)P.nbins / (P.hist_hi - P.hist_lo);
        __syncwarp();  // the previous centre's shared-memory reads are over

        Top4<T> top;
        top.reset();
        int n_sel = 0;          // q-eligible candidates inside the selection radius
        bool q_done = !doq;
        bool q_from_list = false;
        int K3 = 0, Kb = 0;
        bool overflow = false;

        // ---------------- sweep at half-width 1: three-body list (+ q candidates) -----------------
        const bool sweep1 = valid && (do3 || (doq && w_start <= 1));
        const bool q_in_sweep1 = doq && w_start <= 1;
        // selection radius at half-width 1
        const bool full1 = (3 >= P.nc0) && (3 >= P.nc1) && (3 >= P.nc2);
        const bool last1 = full1 || P.wq_max <= 1;
        T selsq1;
        {
            const double r = last1 ? P.highq : fmin(P.highq, P.rc1);
            selsq1 = last1 ? highqsq : (T)fmin((double)highqsq, r * r);
        }
        {
            const int total = build_segments(sweep1, f, cx, cy, cz, 1);
            int seg = 0;
            for (int t0 = 0; __any_sync(kFullMask, t0 < total); t0 += G) {
                const int t = t0 + gl;
                const bool act = t < total;
                bool in3 = false, inq = false;
                T dx = 0, dy = 0, dz = 0, s = 0;
                int idx = -1;
                if (act) {
                    while (t >= S.segpref[seg + 1]) ++seg;
                    const size_t j = (size_t)(S.segj0[seg] + (t - S.segpref[seg]));
                    T px, py, pz;
                    RecTraits<T>::load(P.recs, j, px, py, pz, idx);
                    dx = min_image_1<T, EXACT>(px, rx, Lx, iLx);
                    dy = min_image_1<T, EXACT>(py, ry, Ly, iLy);
                    dz = min_image_1<T, EXACT>(pz, rz, Lz, iLz);
                    s = sumsq3<T>(dx, dy, dz);
                    in3 = do3 && (s > low3sq) && (s <= high3sq);
                    inq = q_in_sweep1 && (s > lowqsq) && (s <= selsq1);
                }
                const unsigned b3 = group_ballot(in3);
                const unsigned bq = group_ballot(inq && !in3);
                const unsigned lt = (1u << gl) - 1u;
                const int n3new = __popc(b3), nbnew = __popc(bq);
                if (K3 + Kb + n3new + nbnew > CAP) overflow = true;
                if (!overflow) {
                    int slot = -1;
                    if (in3) slot = K3 + __popc(b3 & lt);
                    else if (inq) slot = CAP - 1 - (Kb + __popc(bq & lt));
                    if (slot >= 0) {
                        Vec4<T> v;
                        v.x = dx; v.y = dy; v.z = dz; v.w = s;
                        S.ent[slot] = v;
                        S.eidx[slot] = idx;
                    }
                    K3 += n3new;
                    Kb += nbnew;
                }
            }
            __syncwarp();
        }

        if (overflow) {
            if (!LISTMODE) {
                // hand the whole centre to the large-capacity pass
                if (gl == 0) {
                    const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
                    P.fb_list[at] = fb_id | (do3 ? kFbNeed3b : 0u) | (doq ? kFbNeedQ : 0u);
                    atomicAdd(P.counters + kCntOverflow, 1u);
                }
            } else if (gl == 0) {
                atomicAdd(P.counters + kCntFatal, 1u);
            }
            valid = false;
            do3 = false;
            doq = false;
            q_done = true;
            K3 = Kb = 0;
        }

        // ---------------- phase 2: reimaged vectors, norms, top-4 over the list --------------------
        {
            const int nent = valid ? K3 + Kb : 0;
            for (int m0 = 0; __any_sync(kFullMask, m0 < nent); m0 += G) {
                const int m = m0 + gl;
                if (m < nent) {
                    const int slot = (m < K3) ? m : CAP - 1 - (m - K3);
                    Vec4<T> v = S.ent[slot];
                    const T s = v.w;
                    // ReimagedPos = RefPos + distvec (waterlib.f90:45); Vec = Pos - RefPos (:694-695)
                    const T ex = Ops<T>::sub(Ops<T>::add(rx, v.x), rx);
                    const T ey = Ops<T>::sub(Ops<T>::add(ry, v.y), ry);
                    const T ez = Ops<T>::sub(Ops<T>::add(rz, v.z), rz);
                    const T ne = sumsq3<T>(ex, ey, ez);
                    v.x = ex; v.y = ey; v.z = ez; v.w = ne;
                    S.ent[slot] = v;
                    if (q_in_sweep1 && (s > lowqsq) && (s <= selsq1)) {
                        ++n_sel;
                        top.insert(Ops<T>::sqrt(ne), S.eidx[slot], slot);
                    }
                }
            }
            __syncwarp();
            if (q_in_sweep1) {
                n_sel = group_sum(n_sel);
                q_from_list = true;
                if (n_sel >= 4 || last1) q_done = true;
            }
        }

        // ---------------- phase 3a: three-body pair angles ----------------------------------------
        if (__any_sync(kFullMask, valid && do3)) {
            const int K = (valid && do3) ? K3 : 0;
            const int npairs = K * (K - 1) / 2;
            for (int p0 = 0; __any_sync(kFullMask, p0 < npairs); p0 += G) {
                const int p = p0 + gl;
                if (p < npairs) {
                    // p = b (b - 1) / 2 + a, a < b
                    int b = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
                    while (b * (b - 1) / 2 > p) --b;
                    while ((b + 1) * b / 2 <= p) ++b;
                    const int a = p - b * (b - 1) / 2;
                    const Vec4<T> va = S.ent[a], vb = S.ent[b];
                    int pos;
                    double c;
                    if (va.w == (T)0 || vb.w == (T)0) {  // coincident positions: CosAngle3 returns 0 (:690-693)
                        pos = (int)tab[P.nbins + 2];
                        c = 1.0;
                    } else {
                        const T dot = dot3<T>(va.x, va.y, va.z, vb.x, vb.y, vb.z);
                        c = (double)clamped_cos<T>(dot, va.w, vb.w);
                        pos = angle_position(c, tab, P.nbins, P.hist_lo, inv_width);
                        if (c != -1.0 && c <= tab[P.nbins + 3] && c >= tab[P.nbins + 4]) {
                            st.tet_count += 1u;
                            st.tet_cos += c;
                            st.tet_cossq += c * c;
                        }
                    }
                    st.n_angles += 1u;
                    if (pos >= 0 && pos < P.nbins) {
                        if (s_hist) atomicAdd(s_hist + pos, 1u);
                        else if (P.ang_hist)
                            atomicAdd(P.ang_hist + (size_t)(P.hist_per_frame ? f : 0) * P.nbins + pos, 1ull);
                    }
                }
            }
            if (valid && do3 && gl == 0) {
                if (P.n3) P.n3[out_index] = K3;
                st.n_neigh += (unsigned)K3;
            }
        }

        // ---------------- q search beyond the list: widen until provably complete -----------------
        // (large-capacity pass only; the fast pass queues the centre instead)
        if (__any_sync(kFullMask, valid && doq && !q_done)) {
            if (!LISTMODE) {
                if (valid && doq && !q_done && gl == 0) {
                    const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
                    P.fb_list[at] = fb_id | kFbNeedQ;
                    atomicAdd(P.counters + kCntWidened, 1u);
                }
                if (!q_done) doq = false;
            } else {
                int w = max(w_start, 2);
                while (__any_sync(kFullMask, valid && doq && !q_done)) {
                    const bool go = valid && doq && !q_done;
                    const bool full = (2 * w + 1 >= P.nc0) && (2 * w + 1 >= P.nc1) && (2 * w + 1 >= P.nc2);
                    const bool last = full || w >= P.wq_max;
                    const double rsel = last ? P.highq : fmin(P.highq, (double)w * P.rc1);
                    const T selsq = last ? highqsq : (T)fmin((double)highqsq, rsel * rsel);
                    if (go) {
                        top.reset();
                        n_sel = 0;
                        q_from_list = false;
                    }
                    const int total = build_segments(go, f, cx, cy, cz, w);
                    int seg = 0;
                    for (int t0 = 0; __any_sync(kFullMask, t0 < total); t0 += G) {
                        const int t = t0 + gl;
                        if (t < total) {
                            while (t >= S.segpref[seg + 1]) ++seg;
                            const size_t j = (size_t)(S.segj0[seg] + (t - S.segpref[seg]));
                            T px, py, pz;
                            int idx;
                            RecTraits<T>::load(P.recs, j, px, py, pz, idx);
                            const T dx = min_image_1<T, EXACT>(px, rx, Lx, iLx);
                            const T dy = min_image_1<T, EXACT>(py, ry, Ly, iLy);
                            const T dz = min_image_1<T, EXACT>(pz, rz, Lz, iLz);
                            const T s = sumsq3<T>(dx, dy, dz);
                            if ((s > lowqsq) && (s <= selsq)) {
                                ++n_sel;
                                const T ex = Ops<T>::sub(Ops<T>::add(rx, dx), rx);
                                const T ey = Ops<T>::sub(Ops<T>::add(ry, dy), ry);
                                const T ez = Ops<T>::sub(Ops<T>::add(rz, dz), rz);
                                top.insert(Ops<T>::sqrt(sumsq3<T>(ex, ey, ez)), idx, (int)j);
                            }
                        }
                    }
                    __syncwarp();
                    const int n_all = group_sum(n_sel);
                    if (go) {
                        n_sel = n_all;
                        if (n_all >= 4 || last) q_done = true;
                    }
                    ++w;
                }
            }
        }

        // ---------------- phase 3b: merge top-4, winners' vectors, q ------------------------------
        if (__any_sync(kFullMask, valid && doq)) {
            const bool go = valid && doq;
            int win_idx[4], win_pay[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                T hd = top.d[0];
                int hi = top.i[0], hp = top.p[0], owner = gl;
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) {
                    const T od = shfl_xor_t(hd, o, G);
                    const int oi = __shfl_xor_sync(kFullMask, hi, o, G);
                    const int op = __shfl_xor_sync(kFullMask, hp, o, G);
                    const int oo = __shfl_xor_sync(kFullMask, owner, o, G);
                    const bool take = key_less(od, oi, hd, hi) || (od == hd && oi == hi && oo < owner);
                    if (take) {
                        hd = od; hi = oi; hp = op; owner = oo;
                    }
                }
                win_idx[r] = hi;
                win_pay[r] = hp;
                if (owner == gl) top.pop();
            }
            const int n_found = go ? min(n_sel, 4) : 0;
            // winners' vectors as tetraCosAng sees them: it is handed the already reimaged position and
            // reimages it again (waterlib.f90:880-883) before CosAngle3 subtracts the centre (:694-695)
            if (go && gl < n_found) {
                T ex, ey, ez;
                const int pay = (gl == 0) ? win_pay[0] : (gl == 1 ? win_pay[1] : (gl == 2 ? win_pay[2] : win_pay[3]));
                if (q_from_list) {
                    const Vec4<T> v = S.ent[pay];
                    ex = v.x; ey = v.y; ez = v.z;
                } else {
                    T px, py, pz;
                    int idx;
                    RecTraits<T>::load(P.recs, (size_t)pay, px, py, pz, idx);
                    const T dx = min_image_1<T, EXACT>(px, rx, Lx, iLx);
                    const T dy = min_image_1<T, EXACT>(py, ry, Ly, iLy);
                    const T dz = min_image_1<T, EXACT>(pz, rz, Lz, iLz);
                    ex = Ops<T>::sub(Ops<T>::add(rx, dx), rx);
                    ey = Ops<T>::sub(Ops<T>::add(ry, dy), ry);
                    ez = Ops<T>::sub(Ops<T>::add(rz, dz), rz);
                }
                // second reimage of (img - ref): always the exact anint, it is only 4 per centre
                const T d2x = Ops<T>::sub(ex, Ops<T>::mul(Lx, anint_exact<T>(Ops<T>::mul(ex, iLx))));
                const T d2y = Ops<T>::sub(ey, Ops<T>::mul(Ly, anint_exact<T>(Ops<T>::mul(ey, iLy))));
                const T d2z = Ops<T>::sub(ez, Ops<T>::mul(Lz, anint_exact<T>(Ops<T>::mul(ez, iLz))));
                Vec4<T> v;
                v.x = Ops<T>::sub(Ops<T>::add(rx, d2x), rx);
                v.y = Ops<T>::sub(Ops<T>::add(ry, d2y), ry);
                v.z = Ops<T>::sub(Ops<T>::add(rz, d2z), rz);
                v.w = sumsq3<T>(v.x, v.y, v.z);
                S.win[gl] = v;
            }
            __syncwarp();
            // The six terms (cos + 1/3)^2 in the order of the reference's angle array: the real angles
            // in triu order, then the 180-degree padding (cos = -1) it appends when the centre has fewer
            // than four neighbours (water_properties.py:379-384); summed left to right like np.sum.
            double term = 0.0;
            if (go && gl < 6) {
                int pa = -1, pb = -1;
                if (n_found == 4) {
                    pa = (gl < 3) ? 0 : (gl < 5 ? 1 : 2);
                    pb = (gl < 3) ? gl + 1 : (gl < 5 ? gl - 1 : 3);
                } else if (n_found == 3) {
                    if (gl < 3) {
                        pa = (gl == 2) ? 1 : 0;
                        pb = (gl == 0) ? 1 : 2;
                    }
                } else if (n_found == 2) {
                    if (gl == 0) {
                        pa = 0;
                        pb = 1;
                    }
                }
                double c = -1.0;
                if (pa >= 0) {
                    const Vec4<T> va = S.win[pa], vb = S.win[pb];
                    if (va.w == (T)0 || vb.w == (T)0) c = 1.0;
                    else c = (double)clamped_cos<T>(dot3<T>(va.x, va.y, va.z, vb.x, vb.y, vb.z), va.w, vb.w);
                }
                const double u = c + (1.0 / 3.0);
                term = u * u;
            }
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc += shfl_t(term, k, G);
            if (go && gl == 0) {
                const double qv = (n_found == 0) ? 0.0 : 1.0 - (3.0 / 8.0) * acc;
                if (P.q) {
                    if (sizeof(T) == 8) reinterpret_cast<double *>(P.q)[out_index] = qv;
                    else reinterpret_cast<float *>(P.q)[out_index] = (float)qv;
                }
                if (P.nn_idx) {
                    int4 o;
                    o.x = (n_found > 0) ? win_idx[0] : -1;
                    o.y = (n_found > 1) ? win_idx[1] : -1;
                    o.z = (n_found > 2) ? win_idx[2] : -1;
                    o.w = (n_found > 3) ? win_idx[3] : -1;
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
    }
};

// ------------------------------------------------------------------------------------------------

constexpr int kFastThreads = 128;
constexpr int kFastG = 8;
constexpr int kFastCap = 32;
constexpr int kFastMaxSeg = 18;
constexpr int kBigThreads = 64;
constexpr int kBigCap = 1024;
constexpr int kLightCap = 32;
constexpr int kBigMaxSeg = 512;

// Fast pass: every centre once.  Persistent blocks walk contiguous tiles of kFastThreads / G centres so
// that the block histogram is flushed once per frame the block touches.
template <typename T, bool EXACT>
__global__ void __launch_bounds__(kFastThreads) q3b_fast_kernel(const __grid_constant__ Q3bParams P) {
    typedef GroupWorker<T, kFastG, kFastCap, kFastMaxSeg, EXACT, false> Worker;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kGroups = kFastThreads / kFastG;
    typename Worker::Smem *gs = reinterpret_cast<typename Worker::Smem *>(smem_raw);
    unsigned char *after = smem_raw + sizeof(typename Worker::Smem) * kGroups;
    const bool smem_hist = P.do_3b && P.ang_hist && P.nbins <= kMaxSmemBins;
    double *s_tab = reinterpret_cast<double *>(after);
    const int tab_len = P.do_3b ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
    const bool smem_tab = P.nbins <= kMaxSmemBins;
    unsigned *s_hist = reinterpret_cast<unsigned *>(after + (smem_tab ? sizeof(double) * tab_len : 0));
    if (smem_tab)
        for (int i = threadIdx.x; i < tab_len; i += blockDim.x) s_tab[i] = P.table[i];
    if (smem_hist)
        for (int i = threadIdx.x; i < P.nbins; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const double *tab = smem_tab ? s_tab : P.table;

    const int group = threadIdx.x / kFastG;
    Worker worker(P, gs[group]);
    LaneStats st;
    st.reset();

    const long long chunk = (P.total_tiles + gridDim.x - 1) / gridDim.x;
    const long long t_begin = chunk * blockIdx.x;
    const long long t_end = min(P.total_tiles, t_begin + chunk);
    int cur_f = -1;
    for (long long tile = t_begin; tile < t_end; ++tile) {
        const int f = (int)(tile / P.tiles_per_frame);
        const int m = (int)(tile - (long long)f * P.tiles_per_frame) * kGroups + group;
        if (f != cur_f) {
            if (cur_f >= 0) {
                flush_stats(P, cur_f, st);
                if (smem_hist && P.hist_per_frame) {
                    __syncthreads();
                    for (int i = threadIdx.x; i < P.nbins; i += blockDim.x) {
                        const unsigned v = s_hist[i];
                        if (v) atomicAdd(P.ang_hist + (size_t)cur_f * P.nbins + i, (unsigned long long)v);
                        s_hist[i] = 0u;
                    }
                    __syncthreads();
                }
            }
            cur_f = f;
        }
        const bool valid = m < P.n_centres && (P.n_valid == nullptr || m < __ldg(P.n_valid + f));
        T rx = 0, ry = 0, rz = 0;
        int cx = 0, cy = 0, cz = 0;
        size_t out_index = 0;
        uint32_t fb_id = 0;
        if (valid) {
            if (P.centres == nullptr) {
                const size_t j = (size_t)f * P.n_pos + m;  // m-th atom of the frame in cell order
                int idx;
                RecTraits<T>::load(P.recs, j, rx, ry, rz, idx);
                out_index = (size_t)f * P.n_pos + idx;
                fb_id = (uint32_t)j;
            } else {
                load_centre<T>(P, f, m, rx, ry, rz);
                out_index = (size_t)f * P.n_centres + m;
                fb_id = (uint32_t)out_index;
            }
            const double *bx = P.box + (size_t)f * 3;
            cx = cell_coord((double)rx, __ddiv_rn(1.0, bx[0]), P.nc0);
            cy = cell_coord((double)ry, __ddiv_rn(1.0, bx[1]), P.nc1);
            cz = cell_coord((double)rz, __ddiv_rn(1.0, bx[2]), P.nc2);
        }
        worker.run(valid, f, rx, ry, rz, cx, cy, cz, out_index, P.do_3b != 0, P.do_q != 0, 1, fb_id,
                   smem_hist ? s_hist : nullptr, tab, st);
    }
    if (cur_f >= 0) flush_stats(P, cur_f, st);
    if (smem_hist) {
        __syncthreads();
        if (cur_f >= 0)
            for (int i = threadIdx.x; i < P.nbins; i += blockDim.x) {
                const unsigned v = s_hist[i];
                if (v) atomicAdd(P.ang_hist + (size_t)(P.hist_per_frame ? cur_f : 0) * P.nbins + i, (unsigned long long)v);
            }
    }
}

// Large-capacity pass over the queued centres: one warp per centre.
// CAP = kBigCap handles the centres that need their three-body list redone (overflow of the fast path);
// CAP = kLightCap handles the q-only ones (the common case) with a small footprint and high occupancy.
template <typename T, bool EXACT, int CAP>
__global__ void __launch_bounds__(kBigThreads) q3b_big_kernel(const __grid_constant__ Q3bParams P) {
    typedef GroupWorker<T, 32, CAP, kBigMaxSeg, EXACT, true> Worker;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename Worker::Smem *gs = reinterpret_cast<typename Worker::Smem *>(smem_raw);
    const int warp = threadIdx.x >> 5;
    Worker worker(P, gs[warp]);
    LaneStats st;
    const uint32_t n_items = P.counters[P.list_counter];
    if (CAP == kBigCap && P.counters[kCntOverflow] == 0u) return;  // nothing overflowed: no list to redo
    const uint32_t warps_total = gridDim.x * (kBigThreads / 32);
    auto process = [&](uint32_t e) {
        st.reset();
        const uint32_t id = e & kFbIdMask;
        const bool do3 = (e & kFbNeed3b) != 0, doq = (e & kFbNeedQ) != 0;
        T rx, ry, rz;
        size_t out_index;
        int f;
        if (P.centres == nullptr) {
            f = (int)(id / (uint32_t)P.n_pos);
            int idx;
            RecTraits<T>::load(P.recs, (size_t)id, rx, ry, rz, idx);
            out_index = (size_t)f * P.n_pos + idx;
        } else {
            f = (int)(id / (uint32_t)P.n_centres);
            load_centre<T>(P, f, (int)(id - (uint32_t)f * P.n_centres), rx, ry, rz);
            out_index = id;
        }
        const double *bx = P.box + (size_t)f * 3;
        const int cx = cell_coord((double)rx, __ddiv_rn(1.0, bx[0]), P.nc0);
        const int cy = cell_coord((double)ry, __ddiv_rn(1.0, bx[1]), P.nc1);
        const int cz = cell_coord((double)rz, __ddiv_rn(1.0, bx[2]), P.nc2);
        // a centre queued only for q already failed at half-width 1; an overflowed one starts over
        worker.run(true, f, rx, ry, rz, cx, cy, cz, out_index, do3, doq, do3 ? 1 : P.list_w_start, id, nullptr, P.table, st);
        flush_stats(P, f, st);
    };
    if (CAP == kBigCap) {
        // overflowed centres are rare entries of a queue that is mostly q-only work: the lanes read 32 entries at a
        // time and the warp handles the matches one by one (a dependent load per entry would cost ~0.5 us each)
        const int lane = threadIdx.x & 31;
        for (uint32_t base = (blockIdx.x * (kBigThreads / 32) + warp) * 32u; base < n_items; base += warps_total * 32u) {
            const uint32_t it = base + lane;
            const uint32_t e_l = it < n_items ? P.list[it] : 0u;
            unsigned m = __ballot_sync(kFullMask, it < n_items && (e_l & kFbNeed3b) != 0u);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                process(__shfl_sync(kFullMask, e_l, src));
            }
        }
    } else {
        for (uint32_t it = blockIdx.x * (kBigThreads / 32) + warp; it < n_items; it += warps_total) {
            const uint32_t e = P.list[it];
            if ((e & kFbNeed3b) != 0u) continue;
            process(e);
        }
    }
}

// the fatal counter is sticky (cleared by wol_status), the others describe the last evaluation
__global__ void reset_counters_kernel(uint32_t *counters) {
    if (threadIdx.x < kNumCounters && threadIdx.x != kCntFatal) counters[threadIdx.x] = 0u;
}

template <typename T, bool EXACT, int CAP>
static int launch_big(const Q3bParams &P, cudaStream_t stream) {
    typedef GroupWorker<T, 32, CAP, kBigMaxSeg, EXACT, true> BigWorker;
    const size_t big_smem = sizeof(typename BigWorker::Smem) * (kBigThreads / 32);
    cudaError_t e = cudaFuncSetAttribute(q3b_big_kernel<T, EXACT, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(big)", e);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, q3b_big_kernel<T, EXACT, CAP>, kBigThreads, big_smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    q3b_big_kernel<T, EXACT, CAP><<<(unsigned)(sm_count() * per_sm), kBigThreads, big_smem, stream>>>(P);
    add_launches(1);
    return WOL_OK;
}

template <typename T, bool EXACT>
static int launch_typed(const Q3bParams &P, cudaStream_t stream, bool use_tpc, double brick_box_max = 0.0) {
    typedef GroupWorker<T, kFastG, kFastCap, kFastMaxSeg, EXACT, false> FastWorker;
    reset_counters_kernel<<<1, 32, 0, stream>>>(P.counters);
    add_launches(1);
    if (P.ev_begin) cudaEventRecord((cudaEvent_t)P.ev_begin, stream);
    if (brick_box_max > 0.0) {
        // warp-specialised kernel where its shapes allow (batch-wide histograms), the single-role one otherwise
        int rc = q3b_brick_ws_supported(P, false) ? q3b_brick_ws_launch(P, brick_box_max, stream) : q3b_brick_launch(P, brick_box_max, stream);
        if (rc != WOL_OK) return rc;
    } else if (use_tpc) {
        // fp32 mode: the brick kernel for large batches where every atom is a centre, else thread per centre
        int rc = (sizeof(T) == 8) ? q3b_tpc_launch(P, stream, EXACT)
                                  : (!EXACT && q3b_brick32_supported(P) ? q3b_brick32_launch(P, stream) : q3b_tpc32_launch(P, stream));
        if (rc != WOL_OK) return rc;
    } else {
        const int tab_len = P.do_3b ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
        size_t fast_smem = sizeof(typename FastWorker::Smem) * (kFastThreads / kFastG);
        if (P.nbins <= kMaxSmemBins) fast_smem += sizeof(double) * tab_len;
        if (P.do_3b && P.ang_hist && P.nbins <= kMaxSmemBins) fast_smem += sizeof(unsigned) * P.nbins;
        cudaError_t e = cudaFuncSetAttribute(q3b_fast_kernel<T, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem);
        if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(fast)", e);
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, q3b_fast_kernel<T, EXACT>, kFastThreads, fast_smem);
        if (e != cudaSuccess || per_sm < 1) per_sm = 1;
        long long grid = (long long)sm_count() * per_sm;
        if (grid > P.total_tiles) grid = P.total_tiles;
        if (grid > 0) {
            q3b_fast_kernel<T, EXACT><<<(unsigned)grid, kFastThreads, fast_smem, stream>>>(P);
            add_launches(1);
        }
    }
    if (P.ev_end) cudaEventRecord((cudaEvent_t)P.ev_end, stream);
    // queued centres.  List overflows: large-capacity group pass.  q-only: thread-per-centre widened pass
    // at half-width 2 when available (what it cannot settle goes to the second-level queue), then the
    // light group pass, which widens without bound.
    int rc = WOL_OK;
    Q3bParams Q = P;
    Q.list = P.fb_list;
    Q.list_counter = kCntFallback;
    Q.list_w_start = 2;
    rc = launch_big<T, EXACT, kBigCap>(Q, stream);
    if (rc != WOL_OK) return rc;
    if (P.do_q) {
        if (use_tpc && q3b_tpc_widen_supported(P)) {
            rc = q3b_tpc_widen_launch(Q, stream, EXACT, sizeof(T) == 4);
            if (rc != WOL_OK) return rc;
            Q.list = P.list2;
            Q.list_counter = kCntLevel2;
            Q.list_w_start = 3;
        }
        rc = launch_big<T, EXACT, kLightCap>(Q, stream);
        if (rc != WOL_OK) return rc;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("q3b launch", e);
    return WOL_OK;
}

int q3b_launch(const wol_q3b_args &a, const WorkspaceLayout &lay, cudaStream_t stream) {
    char *ws = reinterpret_cast<char *>(a.workspace);
    Q3bParams P;
    P.recs = ws + lay.off_recs;
    P.cell_start = reinterpret_cast<const uint32_t *>(ws + lay.off_cell_start);
    P.box = a.box;
    P.centres = a.centres;
    P.n_valid = a.centres ? a.n_valid : nullptr;
    P.centre_dtype = a.centre_dtype;
    P.n_frames = a.n_frames;
    P.n_pos = a.n_pos;
    P.n_centres = a.centres ? a.n_centres : a.n_pos;
    P.nc0 = a.nc[0];
    P.nc1 = a.nc[1];
    P.nc2 = a.nc[2];
    P.low3sq = a.low3 * a.low3;
    P.high3sq = a.high3 * a.high3;
    P.lowqsq = a.lowq * a.lowq;
    P.highqsq = a.highq * a.highq;
    P.highq = a.highq;
    P.rc1 = a.edge_min * (1.0 - 1e-9);
    int wq = 1;
    while ((double)wq * P.rc1 < a.highq && wq < 64) ++wq;
    P.wq_max = wq;
    P.do_q = a.do_q;
    P.do_3b = a.do_3body;
    P.nbins = a.nbins;
    P.q_nbins = a.q_nbins;
    P.hist_lo = a.hist_lo;
    P.hist_hi = a.hist_hi;
    P.table = a.angle_table;
    P.q = a.q;
    P.nn_idx = a.nn_idx;
    P.n3 = a.n3;
    P.ang_hist = reinterpret_cast<unsigned long long *>(a.ang_hist);
    P.q_hist = reinterpret_cast<unsigned long long *>(a.q_hist);
    P.stats = a.frame_stats;
    P.hist_per_frame = a.hist_per_frame;
    P.counters = reinterpret_cast<uint32_t *>(ws + lay.off_counters);
    P.fb_list = reinterpret_cast<uint32_t *>(ws + lay.off_fb_list);
    const int groups = kFastThreads / kFastG;
    P.tiles_per_frame = (P.n_centres + groups - 1) / groups;
    P.total_tiles = (long long)P.tiles_per_frame * a.n_frames;

    // widest stencil the q search can need must fit the segment table of the large-capacity pass
    {
        int w = P.wq_max;
        const int cz = (2 * w + 1 >= P.nc2) ? P.nc2 : 2 * w + 1;
        const int cy = (2 * w + 1 >= P.nc1) ? P.nc1 : 2 * w + 1;
        if ((long long)cz * cy * 2 > kBigMaxSeg && a.do_q)
            return set_error(WOL_ERR_UNSUPPORTED,
                             "q cutoff %.3f needs a stencil of half-width %d cells; plan the grid with a larger r_cell",
                             a.highq, w);
    }
    // exact anint only matters when a cutoff can reach L/2 (see min_image_1)
    const double reach = fmax(a.do_3body ? a.high3 : 0.0, a.do_q ? a.highq : 0.0);
    const bool exact = reach > 0.49 * a.edge_min * (double)(P.nc0 < P.nc1 ? (P.nc0 < P.nc2 ? P.nc0 : P.nc2)
                                                                          : (P.nc1 < P.nc2 ? P.nc1 : P.nc2));
    // thread-per-centre fast path: fp64 mode, >= 4 cells per axis
    P.wrapped = reinterpret_cast<const float4 *>(ws + lay.off_wrapped);
    P.cellpack = reinterpret_cast<const uint32_t *>(ws + lay.off_recs + (size_t)lay.n_atoms_total * sizeof(RecF));
    P.ev_begin = a.timing_event_begin;
    P.ev_end = a.timing_event_end;
    {
        // edge_min * nc underestimates L for the larger frames of an NPT batch; the caller-provided
        // box_max (>= every edge of every frame) is what the rounding margin needs
        const double lmax = a.box_max > 0.0 ? a.box_max : 0.0;
        const bool last1 = P.wq_max <= 1;
        const double rsel = a.do_q ? (last1 ? a.highq : fmin(a.highq, P.rc1)) : 0.0;
        const double rthr = fmax(a.do_3body ? a.high3 : 0.0, rsel);
        const double margin = 16.0 * ldexp(1.0, -24) * lmax;
        const double thr2 = (rthr + margin) * (rthr + margin) * (1.0 + 1e-6);
        P.pre_thr2 = nextafterf((float)thr2, INFINITY);
        // half-width-2 widened pass
        const bool last2 = P.wq_max <= 2;
        const double rsel2 = last2 ? a.highq : fmin(a.highq, 2.0 * P.rc1);
        const double t2 = (rsel2 + margin) * (rsel2 + margin) * (1.0 + 1e-6);
        P.pre_thr2_w2 = nextafterf((float)t2, INFINITY);
        const double cst = (4.0 * margin * (rsel2 + margin) + 4.0 * margin * margin) * (1.0 + 1e-6) + 1e-6 * rsel2 * rsel2;
        P.pre_cst_w2 = nextafterf((float)cst, INFINITY);
        const double lo = (a.lowq + margin) * (a.lowq + margin) * (1.0 + 1e-6);
        P.lowq_hi2 = nextafterf((float)lo, INFINITY);
        P.cell_eps = (float)(4.0 * margin);
    }
    P.list = P.fb_list;
    P.list2 = P.fb_list + lay.n_centres_total;
    P.list_counter = kCntFallback;
    P.list_w_start = 2;
    // WOL_NO_TPC / WOL_NO_TPC32 (environment, read per call): route to the generic group-per-centre kernels that
    // otherwise only serve boxes below four cells per edge -- a test switch (tests/test_gpu_edges.py), not a tuning knob
    const bool use_tpc = q3b_tpc_supported(P) && a.box_max > 0.0 && getenv("WOL_NO_TPC") == nullptr;
    if (a.precision == WOL_PREC_FP64) {
        // large batches where every atom is a centre: brick path (wol_q3b_brick.cu); it feeds the same queues
        if (use_tpc && (q3b_brick_ws_supported(P, exact) || q3b_brick_supported(P, exact))) return launch_typed<double, false>(P, stream, true, a.box_max);
        return exact ? launch_typed<double, true>(P, stream, use_tpc) : launch_typed<double, false>(P, stream, use_tpc);
    }
    const bool use_tpc32 = use_tpc && getenv("WOL_NO_TPC32") == nullptr;
    return exact ? launch_typed<float, true>(P, stream, false) : launch_typed<float, false>(P, stream, use_tpc32);
}

}  // namespace wol
