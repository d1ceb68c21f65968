// Pieces shared by the two brick kernels of K2 (wol_q3b_brick.cu: every consumer warp runs all phases;
// wol_q3b_brick_ws.cu: sweep warps and exact-arithmetic warps with different register budgets): the brick plan,
// the mbarrier / bulk-copy primitives, the producer warp that stages a brick + halo into shared memory, and the
// certified-decision helpers.  See wol_q3b_brick.cu for the description of the phases.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "wol_q3b_common.cuh"

namespace wol {

constexpr int kBkStages = 3;
constexpr int kBkRowCap = 49;                   // (y, z) rows of brick + halo: (nby + 2) (nbz + 2), nby, nbz <= 5
constexpr int kBkCsW = 32;                      // cell starts per row: nbx + 3 <= 32
constexpr int kBkMaxBx = kBkCsW - 3;
constexpr int kBkMaxByz = 5;
constexpr int kBkCRowCap = kBkMaxByz * kBkMaxByz;
constexpr int kBkListCap = 16;                  // prefilter survivors per centre (self included)
constexpr int kBkEntCap = 8;                    // unit vectors per centre (three-body neighbours from the front,
                                                // q-only candidates from the back)
constexpr int kBkMaxPairs = kBkEntCap * (kBkEntCap - 1) / 2;
constexpr unsigned kBkSlotMask = 2047u;         // list entry = float bits of distance^2 with the low 11 bits = slot

struct BrickPlan {
    int nb0, nb1, nb2;        // bricks per axis; brick i covers cells [i nc / nb, (i + 1) nc / nb)
    int bricks_per_frame;
    unsigned total;           // bricks in the batch
    unsigned m_bpf, m_nb0, m_nb1, m_nb2;  // floor(2^32 / d) of the four divisors the producer decodes a brick id with
    float pre_thr3;           // prefilter threshold of the three-body cutoff (< 0: no three-body)
    float pre_cst1;           // slack added to the running 4th-smallest float distance^2
    double eps_a, eps_b;      // eps_c = eps_a * (max |coordinate|) + eps_b
    double floor2;            // neighbours closer than this (squared) send the centre to the exact path
};

struct BkItem {
    int done, frame, n_centres, n_chunks;
    int nbx, nby, n_crows, next;             // next: chunk counter
    double L[3], iL[3];
    int crow_off[kBkCRowCap + 1];            // centres before centre row r
    int crow_g0[kBkCRowCap];                 // place in the cell-sorted arrays of the first centre of row r
    unsigned short crow_slot[kBkCRowCap];    // stage slot of the first centre of row r
    unsigned short crow_hrow[kBkCRowCap];    // its row among brick + halo rows
};

struct BkRow {        // one (y, z) row of brick + halo: producer scratch, kept in shared memory (the producer warp
                      // lives in a low-register warpgroup of the warp-specialised kernel)
    int gA, gM, gB;   // first atom (place in the cell-sorted arrays) of: the image cell at x - L, the main run, the image cell at x + L
    int cA, cM, cB;   // their atom counts
    int cc;           // centres of the row (0 for halo rows)
    int base;         // cell_start index of the row's cell x = 0
    float sy, sz;     // shift of the row when it comes from the other side of the box
    int off;          // stage slot of the row's first atom
    int delta;        // stage slot - cell-sorted index, for the atoms of the main run
};

// ---- mbarrier / bulk-copy primitives (PTX) -------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t a, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    return done != 0u;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const uint32_t a = smem_addr(bar);
    if (mbar_try(a, parity)) return;
    unsigned polls = 0;
    while (!mbar_try(a, parity)) {
        __nanosleep(40);  // leave the issue slots to the warps that have work (the producer shares a scheduler with three consumers)
        // a wait that never completes becomes a launch failure the host sees, not a hung device
        if (++polls > (1u << 24)) __trap();
    }
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// ---- producer ------------------------------------------------------------------------------------------------------

// n / d for a divisor known on the host (m = min(floor(2^32 / d), 2^32 - 1)): one multiply and at most one correction
__device__ __forceinline__ unsigned bk_div(unsigned n, unsigned d, unsigned m, unsigned &rem) {
    unsigned q = __umulhi(n, m);
    unsigned r = n - q * d;
    if (r >= d) { ++q; r -= d; }
    rem = r;
    return q;
}
static inline unsigned bk_div_magic(unsigned d) {
    const unsigned long long m = (1ull << 32) / (d ? d : 1u);
    return m > 0xffffffffull ? 0xffffffffu : (unsigned)m;
}

__device__ __forceinline__ void bk_measure_row(const Q3bParams &P, int f, int rr, int nby, int nbz, int by0, int bz0, int xa, int w,
                                               float Lyf, float Lzf, BkRow &R) {
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
    const int hz = (rr * (65536 / (nby + 2) + 1)) >> 16, hy = rr - hz * (nby + 2);  // rr / (nby + 2), exact for rr < 64, nby + 2 <= 7
    int y = by0 - 1 + hy, z = bz0 - 1 + hz;
    R.sy = 0.f;
    R.sz = 0.f;
    if (y < 0) { y += nc1; R.sy = -Lyf; } else if (y >= nc1) { y -= nc1; R.sy = Lyf; }
    if (z < 0) { z += nc2; R.sz = -Lzf; } else if (z >= nc2) { z -= nc2; R.sz = Lzf; }
    const uint32_t *cs = P.cell_start;
    const int base = (int)(((size_t)f * nc2 + z) * nc1 + y) * nc0;  // < 2^31 (checked on the host)
    R.base = base;
    const int x0 = xa - 1, x1 = xa + w;  // inclusive cell range of the row, may leave [0, nc0)
    R.cA = R.cB = 0;
    R.gA = R.gB = 0;
    if (x0 < 0) {
        R.gA = (int)__ldg(cs + base + nc0 + x0);
        R.cA = (int)__ldg(cs + base + nc0) - R.gA;
    }
    const int m0 = max(x0, 0), m1 = min(x1, nc0 - 1);
    R.gM = (int)__ldg(cs + base + m0);
    R.cM = (int)__ldg(cs + base + m1 + 1) - R.gM;
    if (x1 >= nc0) {
        R.gB = (int)__ldg(cs + base);
        R.cB = (int)__ldg(cs + base + x1 - nc0 + 1) - R.gB;
    }
    R.cc = 0;
    if (hy >= 1 && hy <= nby && hz >= 1 && hz <= nbz) R.cc = (int)__ldg(cs + base + xa + w) - (int)__ldg(cs + base + xa);
    R.off = 0;
    R.delta = 0;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total) {
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(kFullMask, inc, o);
        if (lane >= o) inc += n;
    }
    total = __shfl_sync(kFullMask, inc, 31);
    return inc - v;
}

// centres of a sub-brick that cannot be staged even one cell wide: hand them to the large-capacity pass
static __device__ void bk_route_to_fallback(const Q3bParams &P, int cc, int gc) {
    const uint32_t flags = (P.do_3b ? kFbNeed3b : 0u) | (P.do_q ? kFbNeedQ : 0u);
    const int g0 = (int)__ldg(P.cell_start + gc);
    for (int k = 0; k < cc; ++k) {
        const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
        P.fb_list[at] = (uint32_t)(g0 + k) | flags;
        atomicAdd(P.counters + kCntOverflow, 1u);
    }
}

// SM: the kernel's shared-memory layout (loc, cs, item, bar_full / bar_raw / bar_empty, the producer scratch prow_of /
// pbox_of; SM::kAtomCap atoms per stage).  NP = 1: one producer warp fills the kBkStages stages in turn.  NP = kBkStages:
// producer `pid` owns stage `pid` -- its own sequence of bricks, its own barrier phases -- and the consumers visit the
// stages round robin (the staging of a brick is a long serial piece of code for one warp on a busy SM: about 17 000
// cycles, which a single producer cannot hide behind the consumers once they are fast enough).
// Lane <-> rows lane and lane + 32 of brick + halo; what a lane learns about its rows goes to the shared-memory scratch,
// so the warp needs few registers.
template <class SM, int NP>
__device__ void bk_producer(const Q3bParams &P, const BrickPlan &B, SM &S, int lane, int pid) {
    static_assert(NP == 1 || NP == kBkStages, "one producer, or one per stage");
    BkRow *const prow = S.prow_of(pid);
    double *const pbox = S.pbox_of(pid);
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
    unsigned it = 0;
    int xa = 0, xb = 0, w = 0, f = 0, by0 = 0, nby = 0, bz0 = 0, nbz = 0;
    int box_f = -1;     // frame whose box edges are in S.pbox
    unsigned next_id = 0;  // the next brick id is fetched one brick ahead: the atomic's round trip is off the critical path
    if (lane == 0) next_id = atomicAdd(P.counters + kCntBrick, 1u);
#ifdef WOL_WS_PROF
    long long pp_empty = 0, pp_raw = 0, pp_t0 = clock64();
    int pp_n = 0;
#endif
    for (;;) {
        if (xa >= xb) {  // next brick
            const unsigned id = __shfl_sync(kFullMask, next_id, 0);
            if (id >= B.total) break;
            if (lane == 0) next_id = atomicAdd(P.counters + kCntBrick, 1u);
            unsigned r, ibx, iby, ibz, rem;
            f = (int)bk_div(id, (unsigned)B.bricks_per_frame, B.m_bpf, r);
            const unsigned ryz = bk_div(r, (unsigned)B.nb0, B.m_nb0, ibx);
            ibz = bk_div(ryz, (unsigned)B.nb1, B.m_nb1, iby);
            xa = (int)bk_div(ibx * (unsigned)nc0, (unsigned)B.nb0, B.m_nb0, rem);  // nc <= 1024
            xb = (int)bk_div((ibx + 1u) * (unsigned)nc0, (unsigned)B.nb0, B.m_nb0, rem);
            by0 = (int)bk_div(iby * (unsigned)nc1, (unsigned)B.nb1, B.m_nb1, rem);
            nby = (int)bk_div((iby + 1u) * (unsigned)nc1, (unsigned)B.nb1, B.m_nb1, rem) - by0;
            bz0 = (int)bk_div(ibz * (unsigned)nc2, (unsigned)B.nb2, B.m_nb2, rem);
            nbz = (int)bk_div((ibz + 1u) * (unsigned)nc2, (unsigned)B.nb2, B.m_nb2, rem) - bz0;
            w = xb - xa;
            if (f != box_f) {  // box edges and their reciprocals, once per frame
                __syncwarp();
                if (lane < 3) {
                    const double Lk = P.box[(size_t)f * 3 + lane];
                    pbox[lane] = Lk;
                    pbox[3 + lane] = __ddiv_rn(1.0, Lk);
                }
                __syncwarp();
                box_f = f;
            }
            if (w <= 0 || nby <= 0 || nbz <= 0) { xa = xb; continue; }
        }
        const int nrows = (nby + 2) * (nbz + 2);
        // ---- measure the sub-brick [xa, xa + w): atoms per row, centres per row ------------------------------------
        int n0 = 0, n1 = 0, c0 = 0, c1 = 0;
        __syncwarp();
#pragma unroll (SM::kProducerRegs >= 96 ? 2 : 1)  // both rows' loads in flight where the producer has the registers
        for (int h = 0; h < 2; ++h) {
            const int rr = lane + 32 * h;
            if (rr < nrows) {
                BkRow R;
                bk_measure_row(P, f, rr, nby, nbz, by0, bz0, xa, w, (float)pbox[1], (float)pbox[2], R);
                const int n = R.cA + R.cM + R.cB;
                if (h == 0) { n0 = n; c0 = R.cc; } else { n1 = n; c1 = R.cc; }
                prow[rr] = R;
            }
        }
        int tot0, tot1, ctot0, ctot1;
        const int off0 = warp_excl_scan(n0, lane, tot0);
        const int off1 = tot0 + warp_excl_scan(n1, lane, tot1);
        const int coff0 = warp_excl_scan(c0, lane, ctot0);
        const int coff1 = ctot0 + warp_excl_scan(c1, lane, ctot1);
        const int n_atoms = tot0 + tot1, n_centres = ctot0 + ctot1;
        if (n_centres == 0) { xa += w; w = min(w, xb - xa); continue; }
        if (n_atoms > SM::kAtomCap - 1) {
            if (w > 1) { w = (w + 1) / 2; continue; }
            if (c0 > 0) bk_route_to_fallback(P, c0, prow[lane].base + xa);
            if (c1 > 0) bk_route_to_fallback(P, c1, prow[lane + 32].base + xa);
            if (lane == 0) atomicAdd(P.counters + kCntBrickFb, 1u);
            xa += 1;
            w = min(w, xb - xa);
            continue;
        }
        // ---- stage it ---------------------------------------------------------------------------------------
        const int s = NP == 1 ? (int)(it % kBkStages) : pid;
        const unsigned round = NP == 1 ? it / kBkStages : it;
#ifdef WOL_WS_PROF
        { const long long t_ = clock64(); if (round >= 1) mbar_wait(&S.bar_empty[s], (round + 1u) & 1u); pp_empty += clock64() - t_; ++pp_n; }
#else
        if (round >= 1) mbar_wait(&S.bar_empty[s], (round + 1u) & 1u);
#endif
        BkItem &I = S.item[s];
        unsigned short *cst = S.cs[s];
        float4 *stage = S.loc[s];
        if (lane == 0) {
            I.done = 0;
            I.frame = f;
            I.n_centres = n_centres;
            I.n_chunks = (n_centres + 31) >> 5;
            I.nbx = w;
            I.nby = nby;
            I.n_crows = nby * nbz;
            I.next = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                I.L[k] = pbox[k];
                I.iL[k] = pbox[3 + k];
            }
            I.crow_off[nby * nbz] = n_centres;
            mbar_arrive_expect_tx(&S.bar_raw[s], (unsigned)n_atoms * 16u);
        }
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int rr = lane + 32 * h;
            if (rr >= nrows) continue;
            BkRow &R = prow[rr];
            const int off = h ? off1 : off0;
            const int cA = R.cA, cM = R.cM, cB = R.cB, gM = R.gM;
            if (cA > 0) bulk_g2s(stage + off, P.wrapped + R.gA, (unsigned)cA * 16u, &S.bar_raw[s]);
            if (cM > 0) bulk_g2s(stage + off + cA, P.wrapped + gM, (unsigned)cM * 16u, &S.bar_raw[s]);
            if (cB > 0) bulk_g2s(stage + off + cA + cM, P.wrapped + R.gB, (unsigned)cB * 16u, &S.bar_raw[s]);
            R.off = off;
            R.delta = off + cA - gM;
            const int hz = (rr * (65536 / (nby + 2) + 1)) >> 16, hy = rr - hz * (nby + 2);
            if (hy >= 1 && hy <= nby && hz >= 1 && hz <= nbz) {
                const int r = (hz - 1) * nby + (hy - 1);
                I.crow_off[r] = h ? coff1 : coff0;
                const int g0 = (int)__ldg(P.cell_start + R.base + xa);
                I.crow_g0[r] = g0;
                I.crow_slot[r] = (unsigned short)(off + cA + g0 - gM);
                I.crow_hrow[r] = (unsigned short)rr;
            }
        }
        __syncwarp();
        // ---- while the copies fly: stage slot of every cell start.  Entry i of a row <-> cell xa - 1 + i, entry
        // w + 2 = end of the row.  Lane <-> cell, one coalesced load per row, 12 or 18 rows in flight; the entries of
        // the image cells and the row ends are patched afterwards by the lane that measured the row.
        {
            const int gx = min(max(xa - 1 + lane, 0), nc0 - 1);
            const bool act = lane <= w + 1;
            constexpr int kBatch = SM::kProducerRegs >= 96 ? 18 : 12;  // rows in flight
            for (int r0 = 0; r0 < nrows; r0 += kBatch) {
                int v[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u)
                    if (r0 + u < nrows && act) v[u] = (int)__ldg(P.cell_start + prow[r0 + u].base + gx);
#pragma unroll
                for (int u = 0; u < kBatch; ++u)
                    if (r0 + u < nrows && act) cst[(r0 + u) * kBkCsW + lane] = (unsigned short)(v[u] + prow[r0 + u].delta);
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = lane + 32 * h;
                if (rr >= nrows) continue;
                const BkRow &R = prow[rr];
                unsigned short *row = cst + rr * kBkCsW;
                row[w + 2] = (unsigned short)(R.off + R.cA + R.cM + R.cB);
                if (xa - 1 < 0) row[0] = (unsigned short)R.off;                          // the single cell at x - L
                if (xa + w >= nc0) row[w + 1] = (unsigned short)(R.off + R.cA + R.cM);   // the single cell at x + L
            }
        }
#ifdef WOL_WS_PROF
        { const long long t_ = clock64(); mbar_wait(&S.bar_raw[s], round & 1u); pp_raw += clock64() - t_; }
#else
        mbar_wait(&S.bar_raw[s], round & 1u);
#endif
        // ---- periodic images: rows (or row ends) that come from the other side of the box are shifted in place, so
        // the sweep has no image logic.  Interior bricks have none.  The single image cell at either end of a row is a
        // couple of atoms: every lane shifts those of its own rows; rows that come from across y or z are shifted as
        // a whole, by the whole warp.
        {
            const float Lxf = (float)pbox[0];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = lane + 32 * h;
                bool whole = false;
                if (rr < nrows) {
                    const BkRow &R = prow[rr];
                    whole = R.sy != 0.f || R.sz != 0.f;
                    if (!whole) {
                        const int off = R.off, cA = R.cA, cB = R.cB, offB = off + cA + R.cM;
                        for (int k = 0; k < cA; ++k) stage[off + k].x -= Lxf;
                        for (int k = 0; k < cB; ++k) stage[offB + k].x += Lxf;
                    }
                }
                unsigned todo = __ballot_sync(kFullMask, whole);
                while (todo) {
                    const int src = __ffs(todo) - 1 + 32 * h;
                    todo &= todo - 1;
                    const BkRow &R = prow[src];
                    const int off = R.off, cA = R.cA, cM = R.cM;
                    const float sy = R.sy, sz = R.sz;
                    const int n = cA + cM + R.cB;
                    for (int k = lane; k < n; k += 32) {
                        float4 v = stage[off + k];
                        v.x += k < cA ? -Lxf : (k >= cA + cM ? Lxf : 0.f);
                        v.y += sy;
                        v.z += sz;
                        stage[off + k] = v;
                    }
                }
            }
        }
        mbar_arrive(&S.bar_full[s]);  // 32 arrivals: every lane's metadata and rewritten rows are published
        ++it;
        xa += w;
        w = min(w, xb - xa);
    }
#ifdef WOL_WS_PROF
    if (blockIdx.x == 3 && lane == 0) printf("producer: total %lld wait_empty %lld wait_raw %lld items %d\n", clock64() - pp_t0, pp_empty, pp_raw, pp_n);
#endif
    // no more bricks: publish the end marker in the next stage
    const int s = NP == 1 ? (int)(it % kBkStages) : pid;
    const unsigned round = NP == 1 ? it / kBkStages : it;
    if (round >= 1) mbar_wait(&S.bar_empty[s], (round + 1u) & 1u);
    if (lane == 0) S.item[s].done = 1;
    mbar_arrive(&S.bar_full[s]);
}

// ---- consumers -----------------------------------------------------------------------------------------------------

// Sorted four smallest squared distances with the column entry each belongs to.
struct Top4S {
    double d[4];
    int p[4];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            d[k] = Ops<double>::inf();
            p[k] = 0;
        }
    }
    // strict <: equal keys keep their arrival order; ties are caught by the gap test afterwards
    __device__ __forceinline__ void insert(double dd, int pp) {
        d[3] = dd;
        p[3] = pp;
#pragma unroll
        for (int k = 3; k > 0; --k) {
            if (d[k] < d[k - 1]) {
                const double td = d[k]; d[k] = d[k - 1]; d[k - 1] = td;
                const int tp = p[k]; p[k] = p[k - 1]; p[k - 1] = tp;
            }
        }
    }
};

// The reference's clamped cosine for one pair, from the fp64 records, every operation as the Fortran
// performs it (waterlib.f90:880-883 second reimage is the identity here: |v| < L / 2).
static __device__ __noinline__ double bk_exact_pair(const RecD *recs, int gc, int ga, int gb, const double *L, const double *iL) {
    const RecD c = recs[gc], a = recs[ga], b = recs[gb];
    double va[3], vb[3];
    const double r[3] = {c.x, c.y, c.z}, pa[3] = {a.x, a.y, a.z}, pb[3] = {b.x, b.y, b.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double da = min_image_1<double, false>(pa[k], r[k], L[k], iL[k]);
        const double db = min_image_1<double, false>(pb[k], r[k], L[k], iL[k]);
        va[k] = __dsub_rn(__dadd_rn(r[k], da), r[k]);
        vb[k] = __dsub_rn(__dadd_rn(r[k], db), r[k]);
    }
    const double wa = sumsq3<double>(va[0], va[1], va[2]), wb = sumsq3<double>(vb[0], vb[1], vb[2]);
    if (wa == 0.0 || wb == 0.0) return 1.0;  // (cannot happen: such centres never reach the pair phase)
    return clamped_cos<double>(dot3<double>(va[0], va[1], va[2], vb[0], vb[1], vb[2]), wa, wb);
}

// (the statistics stay in registers: the out-of-line flush gets a copy)
static __device__ __noinline__ void bk_flush_stats_copy(const Q3bParams &P, int f, LaneStats st) { flush_stats(P, f, st); }
__device__ __forceinline__ void bk_flush_stats(const Q3bParams &P, int f, LaneStats &st) {
    bk_flush_stats_copy(P, f, st);
    st.reset();
}
static __device__ __noinline__ int bk_exact_position(double c, const double *tab, int nbins, float lo_f, float invw_f) {
    return angle_position(c, tab, nbins, lo_f, invw_f);
}

// float seed of the bin of cosine c, always a valid bin index (see angle_position, which it mirrors)
__device__ __forceinline__ int bk_seed_position(double c, int nbins, float lo, float inv_width) {
    const float x = (float)c, ax = fabsf(x);
    const float t = fmaxf(1.0f - ax, 1e-30f);
    float r = fmaf(fmaf(fmaf(-0.0187293f, ax, 0.0742610f), ax, -0.2121144f), ax, 1.5707288f) * (t * rsqrtf(t));
    r = x < 0.f ? 3.14159265f - r : r;
    const int k = (int)((r * 57.29577951f - lo) * inv_width);
    return min(max(k, 0), nbins - 1);
}

__device__ __forceinline__ void bk_push_q(const Q3bParams &P, uint32_t fb_id) {
    const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
    P.fb_list[at] = fb_id | kFbNeedQ;
    atomicAdd(P.counters + kCntWidened, 1u);
}

// one 256-bit load (LDG.256) per 32-byte record
__device__ __forceinline__ void bk_load_rec(const void *recs, int g, double &x, double &y, double &z, int &idx) {
    const RecD *p = reinterpret_cast<const RecD *>(recs) + g;
    long long a, b, c, d;
    asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    x = __longlong_as_double(a);
    y = __longlong_as_double(b);
    z = __longlong_as_double(c);
    idx = (int)d;
}

// ---- host side ------------------------------------------------------------------------------------------------------

// Bricks per axis: the largest bricks (<= 29 x 5 x 5 cells) whose expected brick + halo population fits a stage of
// `atom_cap` atoms with some head-room and whose centres number about `want_centres`.  A brick that turns out denser
// than expected is split by the producer, so this is a throughput choice, not a correctness one.
static inline void brick_dims(const Q3bParams &P, int nb[3], int want_centres, int atom_cap) {
    const double occ = (double)P.n_pos / ((double)P.nc0 * P.nc1 * P.nc2);  // atoms per cell
    int by = P.nc1 < 4 ? P.nc1 : 4, bz = P.nc2 < 4 ? P.nc2 : 4;
    const double want_cells = 0.97 * want_centres / (occ > 1e-9 ? occ : 1e-9);
    int bx = (int)(want_cells / (by * bz));
    if (bx > kBkMaxBx) bx = kBkMaxBx;
    if (bx > P.nc0) bx = P.nc0;
    if (bx < 1) bx = 1;
    auto halo = [&](int x, int y, int z) { return (double)(x + 2) * (y + 2) * (z + 2) * occ; };
    const double room = 0.92 * (atom_cap - 1);
    while (bx > 1 && halo(bx, by, bz) > room) --bx;
    while (by > 1 && halo(bx, by, bz) > room) --by;
    while (bz > 1 && halo(bx, by, bz) > room) --bz;
    // even split: nb bricks of floor / ceil (nc / nb) cells, none larger than the limits above
    auto count = [](int nc, int b, int bmax) {
        int n = (nc + b - 1) / b;
        while ((nc + n - 1) / n > bmax) ++n;
        return n;
    };
    nb[0] = count(P.nc0, bx, kBkMaxBx);
    nb[1] = count(P.nc1, by, kBkMaxByz);
    nb[2] = count(P.nc2, bz, kBkMaxByz);
}

// Thresholds and error bounds of the plan (everything but the brick counts)
static inline void brick_plan_bounds(const Q3bParams &P, double box_max, BrickPlan &B) {

    // float thresholds, same rounding margin as the thread-per-centre path (see q3b_launch)
    const double margin = 16.0 * ldexp(1.0, -24) * box_max;
    const bool last1 = P.wq_max <= 1;
    const double high3 = sqrt(P.high3sq);
    const double rsel = P.do_q ? (last1 ? P.highq : fmin(P.highq, P.rc1)) : 0.0;
    const double rthr = fmax(P.do_3b ? high3 : 0.0, rsel);
    B.pre_thr3 = P.do_3b ? nextafterf((float)((high3 + margin) * (high3 + margin) * (1.0 + 1e-6)), INFINITY) : -1.0f;
    const double cst = (4.0 * margin * (rthr + margin) + 4.0 * margin * margin) * (1.0 + 1e-6) + 1e-6 * rthr * rthr;
    // + what dropping 11 mantissa bits of a survivor's distance^2 can hide (phase 1b)
    B.pre_cst1 = nextafterf((float)(cst + 2.0 * ldexp(1.0, -12) * (rthr + margin) * (rthr + margin) * (1.0 + 1e-6)), INFINITY);
    // eps_c (bk header): the reference's vectors (r + d) - r differ from d by at most delta = 2^-51 (|r| + reach) per
    // component; two such vectors of length >= r_floor turn the cosine by at most 2 sqrt(3) delta / r_floor; the
    // roundings of either evaluation add less than 2^-48.  Factor 4 of safety on the first term.
    const double r_floor = 0.25, reach = rthr + margin + 1.0;
    B.floor2 = r_floor * r_floor;
    B.eps_a = 4.0 * 2.0 * sqrt(3.0) * ldexp(1.0, -51) / r_floor;
    B.eps_b = B.eps_a * reach + ldexp(1.0, -46);  // eps_c = eps_a (max |coordinate| + reach) + 2^-46
}

}  // namespace wol
