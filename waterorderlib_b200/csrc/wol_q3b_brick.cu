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
#include <math.h>
#include <stdlib.h>

#include "wol_q3b_common.cuh"

namespace wol {

constexpr int kBkWarps = 15;                    // consumer warps (16 warps with the producer: 128 registers each;
                                                // a 17th warp would be charged as 20 -- warps are allocated in fours)
constexpr int kBkConsumers = kBkWarps * 32;
constexpr int kBkThreads = kBkConsumers + 32;   // + one producer warp
constexpr int kBkStages = 3;
constexpr int kBkAtomCap = 1536;                // atoms of brick + halo per stage (slot fits 11 bits)
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
static_assert(kBkAtomCap <= 2048, "slot must fit 11 bits");

struct BrickPlan {
    int nb0, nb1, nb2;        // bricks per axis; brick i covers cells [i nc / nb, (i + 1) nc / nb)
    int bricks_per_frame;
    unsigned total;           // bricks in the batch
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
    unsigned short crow_slot[kBkCRowCap];    // stage slot of the first centre of row r
    unsigned short crow_hrow[kBkCRowCap];    // its row among brick + halo rows
};

struct BkRowInfo {    // one (y, z) row of brick + halo (producer scratch)
    int base;         // cell_start index of the row's cell x = 0
    int delta;        // stage slot - cell-sorted index, for the atoms of the main run
};

struct BkSmem {
    float4 loc[kBkStages][kBkAtomCap];
    double ent[kBkEntCap][3][kBkConsumers];
    unsigned char ent_k[kBkEntCap][kBkConsumers];  // which kept survivor the entry is (its list entry says where the record is)
    unsigned lj[kBkListCap + 1][kBkConsumers];   // survivor lists (+ one row that absorbs overflowing stores)
    unsigned short cs[kBkStages][kBkRowCap * kBkCsW];
    int cgj[kBkConsumers];                       // where the centre's fp64 record is
    int woff[kBkWarps][33];
    BkItem item[kBkStages];
    BkRowInfo rows[kBkRowCap];
    unsigned long long bar_full[kBkStages], bar_raw[kBkStages], bar_empty[kBkStages];
    unsigned char pair_ab[kBkMaxPairs + 4];
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
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kBkConsumers) : "memory"); }

// ---- producer ------------------------------------------------------------------------------------------------------

struct BkRow {        // as the lane that measures the row sees it
    int gA, gM, gB, cA, cM, cB;
    int cc, gc;       // centres of the row (0 for halo rows), cell_start index of the first centre cell
    int base;
    float sy, sz;
    int off;
};

__device__ __forceinline__ void bk_measure_row(const Q3bParams &P, int f, int rr, int nby, int nbz, int by0, int bz0, int xa, int w,
                                               float Lyf, float Lzf, BkRow &R) {
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
    const int hz = rr / (nby + 2), hy = rr - hz * (nby + 2);
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
    R.gc = base + xa;
    if (hy >= 1 && hy <= nby && hz >= 1 && hz <= nbz) R.cc = (int)__ldg(cs + base + xa + w) - (int)__ldg(cs + base + xa);
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
__device__ void bk_route_to_fallback(const Q3bParams &P, const BkRow &R) {
    const uint32_t flags = (P.do_3b ? kFbNeed3b : 0u) | (P.do_q ? kFbNeedQ : 0u);
    const int g0 = (int)__ldg(P.cell_start + R.gc);
    for (int k = 0; k < R.cc; ++k) {
        const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
        P.fb_list[at] = (uint32_t)(g0 + k) | flags;
        atomicAdd(P.counters + kCntOverflow, 1u);
    }
}

__device__ void bk_producer(const Q3bParams &P, const BrickPlan &B, BkSmem &S, int lane) {
    const int nc0 = P.nc0, nc1 = P.nc1, nc2 = P.nc2;
    unsigned it = 0;
    int xa = 0, xb = 0, w = 0, f = 0, by0 = 0, nby = 0, bz0 = 0, nbz = 0;
    double Lx = 1, Ly = 1, Lz = 1;
    int box_f = -1;
    unsigned next_id = 0;  // the next brick id is fetched one brick ahead: the atomic's round trip is off the critical path
    if (lane == 0) next_id = atomicAdd(P.counters + kCntBrick, 1u);
    for (;;) {
        if (xa >= xb) {  // next brick
            const unsigned id = __shfl_sync(kFullMask, next_id, 0);
            if (id >= B.total) break;
            if (lane == 0) next_id = atomicAdd(P.counters + kCntBrick, 1u);
            f = (int)(id / (unsigned)B.bricks_per_frame);
            const int r = (int)(id - (unsigned)f * (unsigned)B.bricks_per_frame);
            const int ibx = r % B.nb0, iby = (r / B.nb0) % B.nb1, ibz = r / (B.nb0 * B.nb1);
            xa = ibx * nc0 / B.nb0;  // nc <= 1024
            xb = (ibx + 1) * nc0 / B.nb0;
            by0 = iby * nc1 / B.nb1;
            nby = (iby + 1) * nc1 / B.nb1 - by0;
            bz0 = ibz * nc2 / B.nb2;
            nbz = (ibz + 1) * nc2 / B.nb2 - bz0;
            w = xb - xa;
            if (f != box_f) {
                Lx = P.box[(size_t)f * 3 + 0];
                Ly = P.box[(size_t)f * 3 + 1];
                Lz = P.box[(size_t)f * 3 + 2];
                box_f = f;
            }
            if (w <= 0 || nby <= 0 || nbz <= 0) { xa = xb; continue; }
        }
        const float Lxf = (float)Lx, Lyf = (float)Ly, Lzf = (float)Lz;
        const int nrows = (nby + 2) * (nbz + 2);
        // ---- measure the sub-brick [xa, xa + w): atoms per row, centres per row (lane <-> rows lane, lane + 32) -----
        BkRow R0, R1;
        R0.cA = R0.cM = R0.cB = R0.cc = 0;
        R1.cA = R1.cM = R1.cB = R1.cc = 0;
        if (lane < nrows) bk_measure_row(P, f, lane, nby, nbz, by0, bz0, xa, w, Lyf, Lzf, R0);
        if (lane + 32 < nrows) bk_measure_row(P, f, lane + 32, nby, nbz, by0, bz0, xa, w, Lyf, Lzf, R1);
        int tot0, tot1, ctot0, ctot1;
        R0.off = warp_excl_scan(R0.cA + R0.cM + R0.cB, lane, tot0);
        R1.off = tot0 + warp_excl_scan(R1.cA + R1.cM + R1.cB, lane, tot1);
        const int coff0 = warp_excl_scan(R0.cc, lane, ctot0);
        const int coff1 = ctot0 + warp_excl_scan(R1.cc, lane, ctot1);
        const int n_atoms = tot0 + tot1, n_centres = ctot0 + ctot1;
        if (n_centres == 0) { xa += w; w = min(w, xb - xa); continue; }
        if (n_atoms > kBkAtomCap - 1) {
            if (w > 1) { w = (w + 1) / 2; continue; }
            if (R0.cc > 0) bk_route_to_fallback(P, R0);
            if (R1.cc > 0) bk_route_to_fallback(P, R1);
            if (lane == 0) atomicAdd(P.counters + kCntBrickFb, 1u);
            xa += 1;
            w = min(w, xb - xa);
            continue;
        }
        // ---- stage it ---------------------------------------------------------------------------------------
        const int s = (int)(it % kBkStages);
        const unsigned round = it / kBkStages;
        if (round >= 1) mbar_wait(&S.bar_empty[s], (round + 1u) & 1u);
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
            I.L[0] = Lx; I.L[1] = Ly; I.L[2] = Lz;
            I.iL[0] = __ddiv_rn(1.0, Lx);
            I.iL[1] = __ddiv_rn(1.0, Ly);
            I.iL[2] = __ddiv_rn(1.0, Lz);
            I.crow_off[nby * nbz] = n_centres;
            mbar_arrive_expect_tx(&S.bar_raw[s], (unsigned)n_atoms * 16u);
        }
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const BkRow &R = h ? R1 : R0;
            const int rr = lane + 32 * h;
            if (rr >= nrows) continue;
            if (R.cA > 0) bulk_g2s(stage + R.off, P.wrapped + R.gA, (unsigned)R.cA * 16u, &S.bar_raw[s]);
            if (R.cM > 0) bulk_g2s(stage + R.off + R.cA, P.wrapped + R.gM, (unsigned)R.cM * 16u, &S.bar_raw[s]);
            if (R.cB > 0) bulk_g2s(stage + R.off + R.cA + R.cM, P.wrapped + R.gB, (unsigned)R.cB * 16u, &S.bar_raw[s]);
            S.rows[rr].base = R.base;
            S.rows[rr].delta = R.off + R.cA - R.gM;
            const int hz = rr / (nby + 2), hy = rr - hz * (nby + 2);
            if (hy >= 1 && hy <= nby && hz >= 1 && hz <= nbz) {
                const int r = (hz - 1) * nby + (hy - 1);
                I.crow_off[r] = h ? coff1 : coff0;
                I.crow_slot[r] = (unsigned short)(R.off + R.cA + (int)__ldg(P.cell_start + R.gc) - R.gM);
                I.crow_hrow[r] = (unsigned short)rr;
            }
        }
        __syncwarp();
        // ---- while the copies fly: stage slot of every cell start.  Entry i of a row <-> cell xa - 1 + i, entry
        // w + 2 = end of the row.  Lane <-> cell, one coalesced load per row, eight rows in flight; the entries of
        // the image cells and the row ends are patched afterwards by the lane that measured the row.
        {
            const int gx = min(max(xa - 1 + lane, 0), nc0 - 1);
            const bool act = lane <= w + 1;
            for (int r0 = 0; r0 < nrows; r0 += 8) {
                int v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (r0 + u < nrows && act) v[u] = (int)__ldg(P.cell_start + S.rows[r0 + u].base + gx);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (r0 + u < nrows && act) cst[(r0 + u) * kBkCsW + lane] = (unsigned short)(v[u] + S.rows[r0 + u].delta);
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const BkRow &R = h ? R1 : R0;
                const int rr = lane + 32 * h;
                if (rr >= nrows) continue;
                unsigned short *row = cst + rr * kBkCsW;
                row[w + 2] = (unsigned short)(R.off + R.cA + R.cM + R.cB);
                if (xa - 1 < 0) row[0] = (unsigned short)R.off;                          // the single cell at x - L
                if (xa + w >= nc0) row[w + 1] = (unsigned short)(R.off + R.cA + R.cM);   // the single cell at x + L
            }
        }
        mbar_wait(&S.bar_raw[s], round & 1u);
        // ---- periodic images: rows (or row ends) that come from the other side of the box are shifted in place, so
        // the sweep has no image logic.  Interior bricks have none.
        {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const BkRow &R = h ? R1 : R0;
                const bool mine = (lane + 32 * h < nrows) && (R.sy != 0.f || R.sz != 0.f || R.cA > 0 || R.cB > 0);
                unsigned todo = __ballot_sync(kFullMask, mine);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int off = __shfl_sync(kFullMask, R.off, src), cA = __shfl_sync(kFullMask, R.cA, src);
                    const int cM = __shfl_sync(kFullMask, R.cM, src), cB = __shfl_sync(kFullMask, R.cB, src);
                    const float sy = __shfl_sync(kFullMask, R.sy, src), sz = __shfl_sync(kFullMask, R.sz, src);
                    const int n = cA + cM + cB;
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
    // no more bricks: publish the end marker in the next stage
    const int s = (int)(it % kBkStages);
    const unsigned round = it / kBkStages;
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

__device__ __forceinline__ void bk_flush_bins(unsigned *s_bins, unsigned long long *g_bins, int nbins, bool clear, int tid) {
    for (int i = tid; i < nbins; i += kBkConsumers) {
        const unsigned v = s_bins[i];
        if (v) atomicAdd(g_bins + i, (unsigned long long)v);
        if (clear) s_bins[i] = 0u;
    }
}

__device__ __forceinline__ void bk_push_q(const Q3bParams &P, uint32_t fb_id) {
    const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
    P.fb_list[at] = fb_id | kFbNeedQ;
    atomicAdd(P.counters + kCntWidened, 1u);
}

__device__ __forceinline__ void bk_load_rec(const void *recs, int g, double &x, double &y, double &z, int &idx) {
    const int4 *p = reinterpret_cast<const int4 *>(reinterpret_cast<const RecD *>(recs) + g);
    const int4 a = __ldg(p), b = __ldg(p + 1);
    x = __hiloint2double(a.y, a.x);
    y = __hiloint2double(a.w, a.z);
    z = __hiloint2double(b.y, b.x);
    idx = b.z;
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
        bk_producer(P, B, S, lane);
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
                            S.ent[e][0][tid] = dx * rs;
                            S.ent[e][1][tid] = dy * rs;
                            S.ent[e][2][tid] = dz * rs;
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
                    double c = fma(S.ent[ea][0][col], S.ent[eb][0][col],
                                   fma(S.ent[ea][1][col], S.ent[eb][1][col], S.ent[ea][2][col] * S.ent[eb][2][col]));
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
                    ux[k] = S.ent[e][0][tid];
                    uy[k] = S.ent[e][1][tid];
                    uz[k] = S.ent[e][2][tid];
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

// Bricks per axis: the largest bricks (<= 29 x 5 x 5 cells) whose expected brick + halo population fits the stage
// with some head-room and whose centres about fill the consumer warps.  A brick that turns out denser than
// expected is split by the producer, so this is a throughput choice, not a correctness one.
static void brick_dims(const Q3bParams &P, int nb[3]) {
    const double occ = (double)P.n_pos / ((double)P.nc0 * P.nc1 * P.nc2);  // atoms per cell
    int by = P.nc1 < 4 ? P.nc1 : 4, bz = P.nc2 < 4 ? P.nc2 : 4;
    const double want_cells = 0.97 * kBkConsumers / (occ > 1e-9 ? occ : 1e-9);
    int bx = (int)(want_cells / (by * bz));
    if (bx > kBkMaxBx) bx = kBkMaxBx;
    if (bx > P.nc0) bx = P.nc0;
    if (bx < 1) bx = 1;
    auto halo = [&](int x, int y, int z) { return (double)(x + 2) * (y + 2) * (z + 2) * occ; };
    const double room = 0.92 * (kBkAtomCap - 1);
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

bool q3b_brick_supported(const Q3bParams &P, bool exact) {
    if (P.centres != nullptr || P.n_valid != nullptr || P.wrapped == nullptr || exact) return false;
    if (P.nc0 < 4 || P.nc1 < 4 || P.nc2 < 4) return false;
    if (brick_smem_bytes(P) > 227u * 1024u) return false;
    const char *env = getenv("WOL_BRICK");  // test switch: 1 = also for small batches, 0 = never
    if (env && env[0] == '0') return false;
    if (env && env[0] == '1') return true;
    int nb[3];
    brick_dims(P, nb);
    return (long long)nb[0] * nb[1] * nb[2] * P.n_frames >= 2LL * sm_count();
}

int q3b_brick_launch(const Q3bParams &P, double box_max, cudaStream_t stream) {
    BrickPlan B;
    int nb[3];
    brick_dims(P, nb);
    B.nb0 = nb[0];
    B.nb1 = nb[1];
    B.nb2 = nb[2];
    B.bricks_per_frame = nb[0] * nb[1] * nb[2];
    const long long total = (long long)B.bricks_per_frame * P.n_frames;
    if (total >= (1LL << 31)) return set_error(WOL_ERR_RANGE, "too many bricks");
    B.total = (unsigned)total;
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
