// K2, warp-specialised brick path (sm_100a): the dominant kernel for large frames, fp64 mode, every atom a centre,
// histograms accumulated over the batch.  Same bricks, same phases and the same certified decisions as
// wol_q3b_brick.cu (whose header describes them); what differs is WHO runs the phases.
//
// The single-role kernel keeps 15 consumer warps of 124 registers each resident per SM, and every one of them is
// latency-bound (dependent fp64 chains, shared-memory round trips): 57 % of the issue slots busy.  The register file
// is what limits the warp count, yet only the exact phases need that many registers.  Here one persistent CTA per SM
// holds three kinds of warps with register budgets moved between warpgroups by setmaxnreg:
//
//   producer (1 warp)        stages brick + halo rows with bulk async copies (bk_producer, shared with the other kernel)
//   sweep warps (S)          phase 1 (float prefilter over the nine stencil rows, operands in shared memory) and
//                            phase 1b (which survivors can matter), a few dozen registers each.  The compacted
//                            survivor list of a chunk of 32 centres goes into a slot of a shared-memory ring.
//   exact warps (E)          take ring slots in ticket order: phase 2 (fp64 re-evaluation from the records in L2),
//                            phase 3a (three-body pairs flattened over the warp), phase 3b (q) -- the register-hungry
//                            part -- and hand the slot back as soon as phase 2 has read it.
//
// Ring protocol: a sweep warp takes ticket t (atomic on q_tail), waits until slot t % R has been handed back
// (ring_empty, phase t / R - 1), fills it and arrives on ring_full; an exact warp takes ticket h (q_head), waits for
// ring_full of slot h % R (phase h / R), reads it and arrives on ring_empty.  Tickets are claimed in order on both
// sides, so every filled slot is consumed exactly once.  At the end the sweep warps publish one poison slot per exact
// warp (frame = -1); an exact warp leaves at its first poison.
#include <stdio.h>

#include "wol_q3b_brick.cuh"

namespace wol {

constexpr int kWsSGroups = 3;                         // warpgroups of the low-register side (producer + sweep warps)
constexpr int kWsEGroups = 3;                         // warpgroups of exact warps
constexpr int kWsProducers = 1;                        // producer warps: 1, or kBkStages = one per stage (measured: three
                                                      // producers + 9 sweep warps lose 9 % against one + 11 -- every warp
                                                      // is then busy, but instruction-fetch stalls go up eightfold)
constexpr int kWsSWarps = kWsSGroups * 4 - kWsProducers;  // sweep warps
constexpr int kWsEWarps = kWsEGroups * 4;
constexpr int kWsFirstE = kWsSGroups * 4;             // first exact warp
constexpr int kWsThreads = (kWsSGroups + kWsEGroups) * 128;
constexpr int kWsRegsS = 56;                          // (12 x 56 + 12 x 104) x 32 = 61 440 = 768 threads x 80 registers
constexpr int kWsRegsE = 104;
constexpr int kWsRing = 24;                           // ring slots (chunks of 32 centres in flight between S and E)
constexpr int kWsAtomCap = 1280;                      // atoms of brick + halo per stage
constexpr int kWsWantCentres = 480;                   // centres per brick the plan aims at
static_assert(kWsAtomCap <= 2048, "slot must fit 11 bits");

struct alignas(16) WsSlot {
    unsigned lj[kBkListCap + 1][32];   // phase 1: survivors (distance^2 | stage slot); after 1b: (mark | place in the sorted arrays)
    int gj[32];                        // the centre's place in the cell-sorted arrays, -1: no centre in this lane
    unsigned char nk[32];              // kept survivors
    double L[3], iL[3];
    int frame;                         // -1: poison
    int pad[3];
};

struct alignas(16) WsExact {           // private to one exact warp
    double ent[kBkEntCap][3][32];      // unit vectors of the kept neighbours
    unsigned ent_g[kBkEntCap][32];     // their places in the sorted arrays
    int cgj[32];
    int woff[36];
};

struct WsSmem {
    static constexpr int kAtomCap = kWsAtomCap;
    static constexpr int kProducerRegs = kWsRegsS;      // registers the producer warp may use
    float4 loc[kBkStages][kWsAtomCap];
    WsSlot ring[kWsRing];
    WsExact ex[kWsEWarps];
    unsigned short cs[kBkStages][kBkRowCap * kBkCsW];
    BkItem item[kBkStages];
    BkRow prow[kWsProducers][kBkRowCap];         // producer scratch, one set per producer warp
    double pbox[kWsProducers][6];
    unsigned long long bar_full[kBkStages], bar_raw[kBkStages], bar_empty[kBkStages];
    unsigned long long ring_full[kWsRing], ring_empty[kWsRing];
    unsigned q_tail, q_head;
    unsigned char pair_ab[kBkMaxPairs + 4];
    __device__ __forceinline__ BkRow *prow_of(int pid) { return prow[pid]; }
    __device__ __forceinline__ double *pbox_of(int pid) { return pbox[pid]; }
};

#ifdef WOL_WS_PROF
#define WS_PROF_DECL long long prof_wait_a = 0, prof_wait_b = 0, prof_t0 = clock64(); int prof_n = 0;
#define WS_PROF_WAIT(acc, stmt) { const long long t_ = clock64(); stmt; acc += clock64() - t_; }
#define WS_PROF_PRINT(role, id) if (blockIdx.x == 3 && lane == 0) printf("%s %2d: total %lld wait_a %lld wait_b %lld chunks %d\n", role, id, clock64() - prof_t0, prof_wait_a, prof_wait_b, prof_n);
#else
#define WS_PROF_DECL
#define WS_PROF_WAIT(acc, stmt) stmt;
#define WS_PROF_PRINT(role, id)
#endif

static_assert(sizeof(WsSmem) + sizeof(double) * (500 + 1 + WOL_TABLE_EXTRA) + 2 * sizeof(unsigned) * 500 <= 227 * 1024,
              "the default histogram sizes must fit beside the ring");

__device__ __forceinline__ void exact_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kWsEWarps * 32) : "memory"); }

// ---- sweep warps ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ WsSlot &ws_acquire_slot(WsSmem &S, int lane, int &slot_out, long long &prof_wait_b) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(&S.q_tail, 1u);
    t = __shfl_sync(kFullMask, t, 0);
    const int slot = (int)(t % kWsRing);
    const unsigned round = t / kWsRing;
    if (round >= 1) WS_PROF_WAIT(prof_wait_b, mbar_wait(&S.ring_empty[slot], (round + 1u) & 1u))
    slot_out = slot;
    return S.ring[slot];
}

__device__ void ws_sweep(const Q3bParams &P, const BrickPlan &B, WsSmem &S, int lane, int sweep_id) {
    const bool do3 = P.do_3b != 0, doq = P.do_q != 0;
    const float pre_thr2 = P.pre_thr2, pre_thr3 = B.pre_thr3, pre_cst1 = B.pre_cst1, lowq_hi2 = P.lowq_hi2;
    const float kInf = __int_as_float(0x7f800000);
#ifdef WOL_WS_PROF
    WS_PROF_DECL
#else
    long long prof_wait_b = 0;
#endif
    // the stages are visited round robin: the order a single producer fills them in; with one producer per stage every
    // stage is its own channel and the loop runs until all of them have said "done"
    unsigned live = (1u << kBkStages) - 1u, parity = 0u;
    for (int s = 0; live != 0u; s = (s + 1 == kBkStages) ? 0 : s + 1) {
        if (!((live >> s) & 1u)) continue;
        WS_PROF_WAIT(prof_wait_a, mbar_wait(&S.bar_full[s], (parity >> s) & 1u))
        parity ^= 1u << s;
        BkItem &I = S.item[s];
        if (I.done) {
            if (kWsProducers == 1) break;  // a single producer fills the stages in turn: its end marker ends everything
            live &= ~(1u << s);
            continue;
        }
        const float4 *loc = S.loc[s];
        const unsigned short *cst = S.cs[s];
        const int nbx = I.nbx, rstride = I.nby + 2, n_centres = I.n_centres, n_chunks = I.n_chunks, n_crows = I.n_crows;
        for (;;) {
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&I.next, 1);
            chunk = __shfl_sync(kFullMask, chunk, 0);
            if (chunk >= n_chunks) break;
            int ring_slot;
            WsSlot &Q = ws_acquire_slot(S, lane, ring_slot, prof_wait_b);
#ifdef WOL_WS_PROF
            ++prof_n;
#endif
            unsigned *const my_list = &Q.lj[0][lane];
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

            // ---------------- phase 1: float prefilter over the 9 rows of the stencil ---------------------
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
                            my_list[min(nl, kBkListCap) * 32] = (__float_as_uint(r2) & ~kBkSlotMask) | (unsigned)j;
                            ++nl;
                        }
                        w = wn;
                        ++j;
                    }
                }
            }
            const bool overflow = nl > kBkListCap;

            // ---------------- phase 1b: which survivors can matter (see wol_q3b_brick.cu) ---------------------
            int nk = 0;
            if (valid && !overflow) {
                float a0 = kInf, a1 = kInf, a2 = kInf, a3 = kInf;
#pragma unroll 2
                for (int k = 0; k < nl; ++k) {
                    const unsigned e = my_list[k * 32];
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
                    const unsigned e = my_list[k * 32];
                    const int j = (int)(e & kBkSlotMask);
                    const float r2 = __uint_as_float(e & ~kBkSlotMask);
                    if (j != slot && r2 <= thr_keep) {
                        my_list[nk * 32] = (unsigned)__float_as_int(loc[j].w) | (r2 <= thr_q ? 0x80000000u : 0u);
                        ++nk;
                    }
                }
            }
            if (valid && overflow) {
                const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
                P.fb_list[at] = (uint32_t)gj | (do3 ? kFbNeed3b : 0u) | (doq ? kFbNeedQ : 0u);
                atomicAdd(P.counters + kCntOverflow, 1u);
            }
            Q.gj[lane] = (valid && !overflow) ? gj : -1;
            Q.nk[lane] = (unsigned char)nk;
            if (lane < 3) {
                Q.L[lane] = I.L[lane];
                Q.iL[lane] = I.iL[lane];
            }
            if (lane == 3) Q.frame = I.frame;
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.ring_full[ring_slot]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.bar_empty[s]);
    }
    WS_PROF_PRINT("sweep", sweep_id)
    // one poison slot per exact warp, dealt over the sweep warps
    const int n_poison = kWsEWarps / kWsSWarps + (sweep_id < kWsEWarps % kWsSWarps ? 1 : 0);
    for (int k = 0; k < n_poison; ++k) {
        int ring_slot;
        WsSlot &Q = ws_acquire_slot(S, lane, ring_slot, prof_wait_b);
        if (lane == 0) Q.frame = -1;
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.ring_full[ring_slot]);
    }
}

// ---- exact warps ---------------------------------------------------------------------------------------------------

// the rare exact re-evaluation of a pair: box edges straight from the frame's box
static __device__ __noinline__ double ws_exact_pair(const Q3bParams &P, int f, int gc, int ga, int gb) {
    double L[3], iL[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        L[k] = P.box[(size_t)f * 3 + k];
        iL[k] = __ddiv_rn(1.0, L[k]);
    }
    return bk_exact_pair(reinterpret_cast<const RecD *>(P.recs), gc, ga, gb, L, iL);
}

__device__ void ws_exact(const Q3bParams &P, const BrickPlan &B, WsSmem &S, WsExact &E, const double *tab, unsigned *s_hist,
                         unsigned *s_qhist, int lane, int etid) {
    const bool do3 = P.do_3b != 0, doq = P.do_q != 0;
    const bool use_hist = do3 && P.ang_hist, use_qhist = doq && P.q_hist;
    const int nbins = P.nbins;
    const double inv_width = (double)nbins / (P.hist_hi - P.hist_lo);
    const float hist_lo_f = (float)P.hist_lo, inv_width_f = (float)inv_width;
    const double tet_c_hi = do3 ? tab[nbins + 3] : 0.0, tet_c_lo = do3 ? tab[nbins + 4] : 0.0;
    const int tet_pos_hi = do3 ? angle_position(tet_c_hi, tab, nbins, hist_lo_f, inv_width_f) : -2;
    const int tet_pos_lo = do3 ? angle_position(tet_c_lo, tab, nbins, hist_lo_f, inv_width_f) : -2;
    const double low3sq = P.low3sq, high3sq = P.high3sq, lowqsq = P.lowqsq, highqsq = P.highqsq;
    const bool last1 = P.wq_max <= 1;
    const double selsq1 = last1 ? highqsq : fmin(highqsq, fmin(P.highq, P.rc1) * fmin(P.highq, P.rc1));
    const HistSpec qhs = hist_spec(0.0, 1.0, P.q_nbins);

    LaneStats st;
    st.reset();
    int cur_f = -1;
    WS_PROF_DECL
    for (;;) {
        unsigned h = 0;
        if (lane == 0) h = atomicAdd(&S.q_head, 1u);
        h = __shfl_sync(kFullMask, h, 0);
        const int ring_slot = (int)(h % kWsRing);
        WS_PROF_WAIT(prof_wait_a, mbar_wait(&S.ring_full[ring_slot], (h / kWsRing) & 1u))
#ifdef WOL_WS_PROF
        ++prof_n;
#endif
        WsSlot &Q = S.ring[ring_slot];
        const int f = Q.frame;
        if (f < 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.ring_empty[ring_slot]);
            break;
        }
        if (f != cur_f) {
            if (cur_f >= 0) bk_flush_stats(P, cur_f, st);
            cur_f = f;
        }
        const int gj = Q.gj[lane];
        const bool valid = gj >= 0;
        const int nk = valid ? (int)Q.nk[lane] : 0;
        const unsigned *my_list = &Q.lj[0][lane];
        double rx = 0, ry = 0, rz = 0;
        int my_idx = 0;
        if (valid) bk_load_rec(P.recs, gj, rx, ry, rz, my_idx);
        E.cgj[lane] = gj;

        // ---------------- phase 2: exact fp64 re-evaluation, unit vectors ---------------------------------
        Top4S top;
        top.reset();
        double rej_min = Ops<double>::inf();  // smallest distance^2 among the marked candidates that did not make the four
        int K3 = 0, Kb = 0, nq = 0;
        float bmax = 0.f;
        bool overflow = false;
        if (valid) bmax = __double2float_ru(fmax(fmax(fabs(rx), fabs(ry)), fabs(rz)));
        if (valid) {
            const double Lx = Q.L[0], Ly = Q.L[1], Lz = Q.L[2], iLx = Q.iL[0], iLy = Q.iL[1], iLz = Q.iL[2];
            double nx = 0, ny = 0, nz = 0;
            int nidx = 0;
            unsigned ne = 0;
            if (nk > 0) {
                ne = my_list[0];
                bk_load_rec(P.recs, (int)(ne & 0x7fffffffu), nx, ny, nz, nidx);
            }
            for (int k = 0; k < nk; ++k) {
                const bool marked = (ne >> 31) != 0u;
                const unsigned g_this = ne & 0x7fffffffu;
                const double px = nx, py = ny, pz = nz;
                if (k + 1 < nk) {  // next survivor's record is in flight while this one is evaluated
                    ne = my_list[(k + 1) * 32];
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
                        const int te = (lane + 3 * e) & 31;  // rotated column: see wol_q3b_brick.cu
                        E.ent[e][0][te] = dx * rs;
                        E.ent[e][1][te] = dy * rs;
                        E.ent[e][2][te] = dz * rs;
                        E.ent_g[e][lane] = g_this;
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
        // the slot has been read: hand it back to the sweep warps
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.ring_empty[ring_slot]);

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
            E.woff[lane] = inc - npair;
            if (lane == 31) E.woff[32] = total;
            __syncwarp();
            const int *woff = E.woff;
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
                const int ea = ab & 15, eb = ab >> 4;
                const int ca = (t + 3 * ea) & 31, cb = (t + 3 * eb) & 31;
                double c = fma(E.ent[ea][0][ca], E.ent[eb][0][cb], fma(E.ent[ea][1][ca], E.ent[eb][1][cb], E.ent[ea][2][ca] * E.ent[eb][2][cb]));
                // Bin of the fast value with its certificate (see wol_q3b_brick.cu)
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
                        c = ws_exact_pair(P, f, E.cgj[t], (int)E.ent_g[ea][t], (int)E.ent_g[eb][t]);
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
                const int te = (lane + 3 * e) & 31;
                ux[k] = E.ent[e][0][te];
                uy[k] = E.ent[e][1][te];
                uz[k] = E.ent[e][2][te];
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
                    o.x = (nf > 0) ? __ldg(&recs[E.ent_g[top.p[0]][lane]].idx) : -1;
                    o.y = (nf > 1) ? __ldg(&recs[E.ent_g[top.p[1]][lane]].idx) : -1;
                    o.z = (nf > 2) ? __ldg(&recs[E.ent_g[top.p[2]][lane]].idx) : -1;
                    o.w = (nf > 3) ? __ldg(&recs[E.ent_g[top.p[3]][lane]].idx) : -1;
                    reinterpret_cast<int4 *>(P.nn_idx)[out_index] = o;
                }
                if (bin >= 0) atomicAdd(s_qhist + bin, 1u);
                st.q_sum += qv;
                st.q_sumsq += qv * qv;
                st.n_centres += 1u;
            }
        }
        __syncwarp();  // E.ent / E.woff are rewritten by the next chunk
    }
    WS_PROF_PRINT("exact", etid >> 5)
    if (cur_f >= 0) bk_flush_stats(P, cur_f, st);
    if (use_hist || use_qhist) {
        exact_bar();
        if (use_hist)
            for (int i = etid; i < nbins; i += kWsEWarps * 32) {
                const unsigned v = s_hist[i];
                if (v) atomicAdd(P.ang_hist + i, (unsigned long long)v);
            }
        if (use_qhist)
            for (int i = etid; i < P.q_nbins; i += kWsEWarps * 32) {
                const unsigned v = s_qhist[i];
                if (v) atomicAdd(P.q_hist + i, (unsigned long long)v);
            }
    }
}

__global__ void __launch_bounds__(kWsThreads, 1) q3b_brick_ws_kernel(const __grid_constant__ Q3bParams P, const __grid_constant__ BrickPlan B) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WsSmem &S = *reinterpret_cast<WsSmem *>(smem_raw);
    unsigned char *after = smem_raw + sizeof(WsSmem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do3 = P.do_3b != 0, doq = P.do_q != 0;
    const bool use_hist = do3 && P.ang_hist, use_qhist = doq && P.q_hist;
    const int tab_len = do3 ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
    double *s_tab = reinterpret_cast<double *>(after);
    unsigned *s_hist = reinterpret_cast<unsigned *>(after + sizeof(double) * tab_len);
    unsigned *s_qhist = s_hist + (use_hist ? P.nbins : 0);
    for (int i = tid; i < tab_len; i += kWsThreads) s_tab[i] = P.table[i];
    if (use_hist)
        for (int i = tid; i < P.nbins; i += kWsThreads) s_hist[i] = 0u;
    if (use_qhist)
        for (int i = tid; i < P.q_nbins; i += kWsThreads) s_qhist[i] = 0u;
    if (tid < kBkMaxPairs) {
        int b = 1;  // p = b (b - 1) / 2 + a, a < b
        while ((b + 1) * b / 2 <= tid) ++b;
        S.pair_ab[tid] = (unsigned char)((tid - b * (b - 1) / 2) | (b << 4));
    }
    if (tid == 0) {
        for (int s = 0; s < kBkStages; ++s) {
            mbar_init(&S.bar_full[s], 32);
            mbar_init(&S.bar_raw[s], 1);
            mbar_init(&S.bar_empty[s], kWsSWarps);
        }
        for (int s = 0; s < kWsRing; ++s) {
            mbar_init(&S.ring_full[s], 1);
            mbar_init(&S.ring_empty[s], 1);
        }
        S.q_tail = 0u;
        S.q_head = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp < kWsFirstE) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsRegsS));
        if (warp < kWsProducers) bk_producer<WsSmem, kWsProducers>(P, B, S, lane, warp);
        else ws_sweep(P, B, S, lane, warp - kWsProducers);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsRegsE));
        ws_exact(P, B, S, S.ex[warp - kWsFirstE], s_tab, s_hist, s_qhist, lane, tid - kWsFirstE * 32);
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------

static size_t ws_smem_bytes(const Q3bParams &P) {
    const int tab_len = P.do_3b ? P.nbins + 1 + WOL_TABLE_EXTRA : 0;
    size_t smem = sizeof(WsSmem) + sizeof(double) * tab_len;
    if (P.do_3b && P.ang_hist) smem += sizeof(unsigned) * P.nbins;
    if (P.do_q && P.q_hist) smem += sizeof(unsigned) * P.q_nbins;
    return smem;
}

// The shapes the single-role brick kernel takes, minus per-frame histograms (the exact warps work on several frames
// at once) and histograms too long for what the ring leaves of the shared memory.  OPT-IN (WOL_BRICK=3, read per
// call): measured on 8 x 1M-water frames this kernel takes 1.675 ms against 1.725 ms of the single-role kernel on
// jittered ice and 2.072 against 2.051 ms on the liquid-like box (profiles/r2_ws_experiment.md) -- the SM is
// throughput-bound on issue slots, the shared-memory pipe and instruction fetch, not on warp latency, so more resident
// warps buy almost nothing.  It stays in the tree as the measured alternative and is parity-checked by
// tests/tools/brick_check.py and tests/test_gpu_q3b.py.
bool q3b_brick_ws_supported(const Q3bParams &P, bool exact) {
    const char *env = getenv("WOL_BRICK");
    if (!env || env[0] != '3') return false;
    if (P.centres != nullptr || P.n_valid != nullptr || P.wrapped == nullptr || exact) return false;
    if (P.nc0 < 4 || P.nc1 < 4 || P.nc2 < 4) return false;
    if (P.hist_per_frame && (P.ang_hist || P.q_hist) && P.n_frames > 1) return false;
    if (ws_smem_bytes(P) > 227u * 1024u) return false;
    return true;
}

int q3b_brick_ws_launch(const Q3bParams &P, double box_max, cudaStream_t stream) {
    BrickPlan B;
    int nb[3];
    brick_dims(P, nb, kWsWantCentres, kWsAtomCap);
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
    const size_t smem = ws_smem_bytes(P);
    cudaError_t e = cudaFuncSetAttribute(q3b_brick_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(brick_ws)", e);
    long long grid = sm_count();
    if (grid > total) grid = total;
    if (grid > 0) {
        q3b_brick_ws_kernel<<<(unsigned)grid, kWsThreads, smem, stream>>>(P, B);
        add_launches(1);
        e = cudaGetLastError();
        if (e != cudaSuccess) return set_cuda_error("brick_ws kernel launch", e);
        if (getenv("WOL_DEBUG_SYNC")) {  // debugging aid: surface device-side faults at the launch that caused them
            e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) return set_cuda_error("brick_ws kernel", e);
        }
    }
    return WOL_OK;
}

}  // namespace wol
