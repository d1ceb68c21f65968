// The other routines of the per-frame path (sm_100a): the value-returning variants the reference's
// Python API exposes (materialised angle lists, dense neighbour matrices, reimage / tetracosang /
// lsidists), np.histogram on the device, hydrogen-bond counting and hydration-shell selection.
// All fp64, reference operation order.  These are the drop-in paths for modest sizes; the throughput path
// is the fused kernel of wol_q3b*.cu.
//
// Reference anchors (relative to /root/reference):
//   getCosAngs            structureLibs/water_properties.py:210-250
//   tetrahedralMetrics    structureLibs/water_properties.py:314-342
//   nearNeighbors         fortran/waterlib.f90:710-743   allNearNeighbors :830-862
//   reimage               fortran/waterlib.f90:32-47     tetraCosAng :867-895   lsiDists :900-918
//   generalHbonds         fortran/waterlib.f90:1156-1210 AngBetween :954-965
//   shell selection       structureLibs/orderParam_lib.py:495-498
#include <math.h>

#include "wol_q3b_common.cuh"

namespace wol {

struct CellGrid {
    const uint32_t *cell_start;
    const void *recs;
    const float4 *wrapped;  // box-wrapped float coordinates in record order (float prefilter)
    int nc0, nc1, nc2;
};

// Visits every record of the half-width-1 stencil around cell (cx, cy, cz) of frame f exactly once
// (axes with <= 3 cells are enumerated completely), as RUNS of records: cells are ordered x-fastest, so the
// (up to) three cells of a stencil row are one contiguous run of the cell-sorted arrays -- two runs where the row wraps
// around the box.  run(j0, j1, sx, sy, sz) gets the records [j0, j1) and, per axis, the periodic image the run's cells
// are adjacent through: -1 / 0 / +1 box edges to ADD to a record's box-wrapped coordinate (only meaningful on an axis
// with more than 3 cells; with fewer the adjacency is ambiguous and the caller must take the minimum image itself).
// Cells are visited in the same order as a cell-by-cell enumeration starting at (cx - 1, cy - 1, cz - 1).
template <typename R>
__device__ __forceinline__ void sweep_stencil1_runs(const CellGrid &g, int f, int cx, int cy, int cz, R &&run) {
    const int cnty = min(3, g.nc1), cntz = min(3, g.nc2);
    const int ys = (g.nc1 <= 3) ? 0 : cy - 1, zs = (g.nc2 <= 3) ? 0 : cz - 1;
    // x: [xa, xb] and, when the row wraps, [xc, xd] after it
    int xa, xb, xc = 0, xd = -1, sxa = 0, sxc = 0;
    if (g.nc0 <= 3) {
        xa = 0;
        xb = g.nc0 - 1;
    } else if (cx == 0) {
        xa = xb = g.nc0 - 1;
        sxa = -1;
        xc = 0;
        xd = 1;
    } else if (cx == g.nc0 - 1) {
        xa = cx - 1;
        xb = cx;
        xc = xd = 0;
        sxc = 1;
    } else {
        xa = cx - 1;
        xb = cx + 1;
    }
    const size_t base = (size_t)f * g.nc0 * g.nc1 * g.nc2;
    for (int iz = 0; iz < cntz; ++iz) {
        int z = zs + iz, sz = 0;
        if (z < 0) {
            z += g.nc2;
            sz = -1;
        } else if (z >= g.nc2) {
            z -= g.nc2;
            sz = 1;
        }
        for (int iy = 0; iy < cnty; ++iy) {
            int y = ys + iy, sy = 0;
            if (y < 0) {
                y += g.nc1;
                sy = -1;
            } else if (y >= g.nc1) {
                y -= g.nc1;
                sy = 1;
            }
            const uint32_t *row = g.cell_start + base + ((size_t)z * g.nc1 + y) * g.nc0;
            const int nrun = (xd >= xc) ? 2 : 1;
            for (int r = 0; r < nrun; ++r) {  // one body for both runs: the caller's code is inlined here once
                const int a = r ? xc : xa, e = r ? xd : xb;
                run((int)__ldg(row + a), (int)__ldg(row + e + 1), r ? sxc : sxa, sy, sz);
            }
        }
    }
}

template <typename F>
__device__ __forceinline__ void sweep_stencil1(const CellGrid &g, int f, int cx, int cy, int cz, F &&fn) {
    sweep_stencil1_runs(g, f, cx, cy, cz, [&](int j0, int j1, int, int, int) {
        for (int j = j0; j < j1; ++j) fn(j);
    });
}

// The same sweep behind a float prefilter: fn(j) runs only for records whose float minimum-image distance^2 from the
// (box-wrapped) point (wx, wy, wz) is within thr2 -- a bound the caller widens by float_margin() so that no record
// inside the exact cutoff can be rejected; everything that decides a result is then recomputed exactly by fn.
struct FloatBox {
    float Lx, Ly, Lz;
};
__device__ __forceinline__ FloatBox float_box(double Lx, double Ly, double Lz) {
    FloatBox b;
    b.Lx = (float)Lx; b.Ly = (float)Ly; b.Lz = (float)Lz;
    return b;
}
// squared float acceptance threshold for an exact cutoff `cut` in a box whose largest edge is lmax: wrapped
// coordinates carry an absolute error below 2^-24 lmax each, the float arithmetic (incl. the rounding of a centre
// shifted by a box edge, below) a few more roundings of that size
__device__ __forceinline__ float float_margin_thr2(double cut, double lmax) {
    const double m = cut + 16.0 * 5.9604644775390625e-8 * lmax;
    return __double2float_ru(m * m * (1.0 + 1e-6));
}
// Two steps, because the exact work is heavy and only a few records per thread pass: the sweep appends the survivors
// to the thread's column of a shared list (a few instructions inside the divergent loop), fn runs over the list in a
// dense loop -- when the list fills up and once more at the end, so there is no capacity limit.
//
// With more than 3 cells on every axis the periodic image of a whole run is known from the cells' adjacency (every
// record within a cutoff <= the cell edge is met through that image, and it is the nearest one because the box has at
// least 4 cells), so the centre is shifted once per run and a candidate costs a 16-byte load, 3 FADD, FMUL, 2 FFMA
// and the compare; the next candidate's load is in flight meanwhile.  On a grid with <= 3 cells on some axis (boxes
// of a few cutoffs: tiny systems) the adjacency image is ambiguous, so there every record goes to fn, which decides
// exactly anyway.
constexpr int kPrefThreads = 128;
constexpr int kPrefCap = 16;
template <typename F>
__device__ __forceinline__ void sweep_stencil1_pref(const CellGrid &g, int f, int cx, int cy, int cz, float wx, float wy, float wz,
                                                    const FloatBox &fb, float thr2, int *list, F &&fn) {
    int nl = 0;
    const bool small = !(g.nc0 > 3 && g.nc1 > 3 && g.nc2 > 3);  // uniform over the grid
    sweep_stencil1_runs(g, f, cx, cy, cz, [&](int j0, int j1, int sx, int sy, int sz) {
        if (j0 >= j1) return;
        const float mx = wx - (float)sx * fb.Lx, my = wy - (float)sy * fb.Ly, mz = wz - (float)sz * fb.Lz;
        float4 w = __ldg(g.wrapped + j0);
        for (int j = j0; j < j1; ++j) {
            const float4 wn = __ldg(g.wrapped + min(j + 1, j1 - 1));
            const float dx = w.x - mx, dy = w.y - my, dz = w.z - mz;
            if (small || fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= thr2) {
                list[nl * kPrefThreads] = j;
                if (++nl == kPrefCap) {
                    for (int k = 0; k < kPrefCap; ++k) fn(list[k * kPrefThreads]);
                    nl = 0;
                }
            }
            w = wn;
        }
    });
    for (int k = 0; k < nl; ++k) fn(list[k * kPrefThreads]);
}

template <typename T>
__device__ __forceinline__ void load3(const void *p, int dtype, size_t i, T &x, T &y, T &z) {
    if (dtype == WOL_F64) {
        const double *d = reinterpret_cast<const double *>(p) + 3 * i;
        x = (T)d[0]; y = (T)d[1]; z = (T)d[2];
    } else {
        const float *d = reinterpret_cast<const float *>(p) + 3 * i;
        x = (T)d[0]; y = (T)d[1]; z = (T)d[2];
    }
}

struct BoxD {
    double Lx, Ly, Lz, iLx, iLy, iLz;
    double near;  // 0.49 x the smallest edge when all three are periodic, else 0 (min_image_3's short cut never taken)
};
// iBoxL = merge(1/BoxL, 0, BoxL >= 0)  (waterlib.f90:41); the dense / small-array routines keep the
// reference's "negative edge = not periodic" rule because they need no cell grid
__device__ __forceinline__ BoxD load_box(const double *b) {
    BoxD o;
    o.Lx = b[0]; o.Ly = b[1]; o.Lz = b[2];
    o.iLx = (o.Lx >= 0.0) ? __ddiv_rn(1.0, o.Lx) : 0.0;
    o.iLy = (o.Ly >= 0.0) ? __ddiv_rn(1.0, o.Ly) : 0.0;
    o.iLz = (o.Lz >= 0.0) ? __ddiv_rn(1.0, o.Lz) : 0.0;
    o.near = (o.Lx > 0.0 && o.Ly > 0.0 && o.Lz > 0.0) ? 0.49 * fmin(o.Lx, fmin(o.Ly, o.Lz)) : 0.0;
    return o;
}

// distvec = p - r ; distvec = distvec - BoxL * anint(distvec * iBoxL) on the three axes (waterlib.f90:43-44), with
// the common case short-cut: when every |p - r| is below 0.49 of the smallest edge, distvec * iBoxL rounds to a
// magnitude below 0.5 on every axis, anint gives 0, and distvec - BoxL * 0 is distvec bit for bit -- so the two
// products and the rounding are skipped (11 of the 12 fp64 instructions of an axis).
__device__ __forceinline__ void min_image_3(double px, double py, double pz, double rx, double ry, double rz, const BoxD &b,
                                            double &dx, double &dy, double &dz) {
    dx = __dsub_rn(px, rx);
    dy = __dsub_rn(py, ry);
    dz = __dsub_rn(pz, rz);
    if (!(fabs(dx) < b.near && fabs(dy) < b.near && fabs(dz) < b.near)) {
        dx = __dsub_rn(dx, __dmul_rn(b.Lx, anint_exact<double>(__dmul_rn(dx, b.iLx))));
        dy = __dsub_rn(dy, __dmul_rn(b.Ly, anint_exact<double>(__dmul_rn(dy, b.iLy))));
        dz = __dsub_rn(dz, __dmul_rn(b.Lz, anint_exact<double>(__dmul_rn(dz, b.iLz))));
    }
}

// A kernel's centre.  From the caller's array (thread g <-> centre g), or -- centres == NULL: every atom of the cell
// list is a centre, as in wol_q3b_frames -- the g-th record of the cell-sorted arrays: the threads of a warp then work
// on atoms of the same or adjacent cells whatever order the caller's atoms are in (overlapping stencils: the same cache
// lines, similar trip counts), and results still go to the atom's own index.
struct CentreRef {
    double x, y, z;
    int cx, cy, cz;
    size_t out;  // frame * n_centres + centre index
};
__device__ __forceinline__ CentreRef get_centre(const CellGrid &g, const void *centres, int dtype, size_t gid, int f, int n_centres,
                                                const BoxD &b) {
    CentreRef c;
    if (centres) {
        load3<double>(centres, dtype, gid, c.x, c.y, c.z);
        c.cx = cell_coord(c.x, b.iLx, g.nc0);
        c.cy = cell_coord(c.y, b.iLy, g.nc1);
        c.cz = cell_coord(c.z, b.iLz, g.nc2);
        c.out = gid;
    } else {
        const RecD *p = reinterpret_cast<const RecD *>(g.recs) + gid;
        long long a, bb, cc, d;
        asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(bb), "=l"(cc), "=l"(d) : "l"(p));
        c.x = __longlong_as_double(a);
        c.y = __longlong_as_double(bb);
        c.z = __longlong_as_double(cc);
        const int cell = (int)(d >> 32);
        c.cx = cell & 1023;
        c.cy = (cell >> 10) & 1023;
        c.cz = (cell >> 20) & 1023;
        c.out = (size_t)f * n_centres + (size_t)(int)(d & 0xffffffffLL);
    }
    return c;
}

// ---- exclusive scan of pair counts (n3 -> angle offsets) -----------------------------------------

__global__ void pair_counts_kernel(const int32_t *__restrict__ n3, size_t n, uint32_t *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t k = (uint32_t)max(n3[i], 0);
        out[i] = k * (k - (k ? 1u : 0u)) / 2u;
    } else if (i == n) {
        out[i] = 0u;
    }
}

// The offsets are 32-bit: 2^26 centres with a hundred neighbours each would wrap them silently.  The counts are summed
// in 64 bits before the scan; a total beyond 2^32 - 2 turns the last offset (the total the caller sizes its output
// by) into the sentinel 0xFFFFFFFF = "does not fit, use smaller batches".
constexpr uint32_t kOffsetsOverflow = 0xFFFFFFFFu;

__global__ void count_total_kernel(const uint32_t *__restrict__ counts, size_t n, unsigned long long *total) {
    unsigned long long s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += counts[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}
__global__ void mark_overflow_kernel(uint32_t *last_offset, const unsigned long long *total) {
    if (*total > 0xFFFFFFFEull) *last_offset = kOffsetsOverflow;
}

// exclusive scan of n + 1 counts (the last one 0) into offsets, with the 64-bit check; scratch: n / 2048 + 8 values
static int scan_offsets_checked(uint32_t *offsets, size_t n_plus_1, uint32_t *scratch, cudaStream_t stream) {
    // the scan uses n_plus_1 / kScanTile + 2 values of scratch; an aligned 8-byte slot behind them holds the total
    size_t slot = n_plus_1 / kScanTile + 3;
    slot += slot & 1;
    unsigned long long *total = reinterpret_cast<unsigned long long *>(scratch + slot);
    cudaError_t e = cudaMemsetAsync(total, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return set_cuda_error("cudaMemsetAsync(offset total)", e);
    const unsigned blocks = (unsigned)((n_plus_1 + 1023) / 1024 < 1184 ? (n_plus_1 + 1023) / 1024 : 1184);
    count_total_kernel<<<blocks ? blocks : 1, 256, 0, stream>>>(offsets, n_plus_1, total);
    exclusive_scan_u32(offsets, n_plus_1, scratch, stream);
    mark_overflow_kernel<<<1, 1, 0, stream>>>(offsets + (n_plus_1 - 1), total);
    add_launches(2);
    return WOL_OK;
}

// ---- materialised three-body angles in the reference's order ---------------------------------------

constexpr int kMatCap = 64;

struct MatParams {
    CellGrid grid;
    const double *box;
    const void *centres;
    int centre_dtype;
    int n_frames, n_pos, n_centres;
    double lowsq, highsq;
    const uint32_t *offsets;  // [n_frames * n_centres + 1]
    double *angles;
    uint32_t *counters;
};

__global__ void __launch_bounds__(kPrefThreads) angles_fill_kernel(const MatParams P) {
    __shared__ int s_list[kPrefCap * kPrefThreads];
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)P.n_frames * P.n_centres;
    if (g >= total) return;
    const int f = (int)(g / P.n_centres);
    const BoxD b = load_box(P.box + (size_t)f * 3);
    const CentreRef ctr = get_centre(P.grid, P.centres, P.centre_dtype, g, f, P.n_centres, b);
    const double rx = ctr.x, ry = ctr.y, rz = ctr.z;
    const int cx = ctr.cx, cy = ctr.cy, cz = ctr.cz;
    const size_t og = ctr.out;  // where this centre's results go
    int idx[kMatCap];
    idx[0] = 0;
    double ex[kMatCap], ey[kMatCap], ez[kMatCap], en[kMatCap];
    int K = 0;
    bool over = false;
    const FloatBox fb = float_box(b.Lx, b.Ly, b.Lz);
    const float thr2 = float_margin_thr2(sqrt(P.highsq), fmax(b.Lx, fmax(b.Ly, b.Lz)));
    sweep_stencil1_pref(P.grid, f, cx, cy, cz, wrapped_coord(rx, b.Lx, b.iLx), wrapped_coord(ry, b.Ly, b.iLy),
                        wrapped_coord(rz, b.Lz, b.iLz), fb, thr2, s_list + threadIdx.x, [&](int j) {
        double px, py, pz;
        int id;
        RecTraits<double>::load(P.grid.recs, (size_t)j, px, py, pz, id);
        double dx, dy, dz;
        min_image_3(px, py, pz, rx, ry, rz, b, dx, dy, dz);
        const double s = sumsq3<double>(dx, dy, dz);
        if (s > P.lowsq && s <= P.highsq) {
            if (K >= kMatCap) {
                over = true;
                return;
            }
            // insertion by atom index: the reference gathers neighbours with a boolean mask, i.e. in
            // ascending index order (water_properties.py:243)
            int at = K;
            while (at > 0 && idx[at - 1] > id) {
                idx[at] = idx[at - 1]; ex[at] = ex[at - 1]; ey[at] = ey[at - 1]; ez[at] = ez[at - 1]; en[at] = en[at - 1];
                --at;
            }
            idx[at] = id;
            ex[at] = __dsub_rn(__dadd_rn(rx, dx), rx);
            ey[at] = __dsub_rn(__dadd_rn(ry, dy), ry);
            ez[at] = __dsub_rn(__dadd_rn(rz, dz), rz);
            en[at] = sumsq3<double>(ex[at], ey[at], ez[at]);
            ++K;
        }
    });
    if (over) {
        atomicAdd(P.counters + kCntFatal, 1u);
        return;
    }
    size_t o = P.offsets[og];
    if ((size_t)P.offsets[og + 1] - o != (size_t)K * (K - 1) / 2) {  // counts and fill disagree: never expected
        atomicAdd(P.counters + kCntFatal, 1u);
        return;
    }
    for (int a = 0; a < K; ++a)
        for (int c = a + 1; c < K; ++c) {
            double ang;
            if (en[a] == 0.0 || en[c] == 0.0) ang = 0.0;
            else ang = angle_deg_from_cos(clamped_cos<double>(dot3<double>(ex[a], ey[a], ez[a], ex[c], ey[c], ez[c]), en[a], en[c]));
            P.angles[o++] = ang;
        }
}

// ---- neighbour lists in CSR form (allNearNeighbors / nearNeighbors without the dense matrix) -------------------------
// Count pass -> exclusive scan -> fill pass.  Within a centre's segment the atom indices ascend, the order the reference
// gets from its boolean-mask gather (water_properties.py:243,372): the cell list holds a cell's atoms in no particular
// order, so each thread sorts its own segment in place (segments are short: ~4-8 at 3.4 A, ~140 at 10 A).

struct CsrParams {
    CellGrid grid;
    const double *box;
    const void *centres;
    int centre_dtype;
    int n_frames, n_pos, n_centres;
    double lowsq, highsq;
    uint32_t *offsets;   // [n_frames * n_centres + 1]
    int32_t *indices;    // [capacity]
    long long capacity;
};

template <bool FILL>
__global__ void __launch_bounds__(kPrefThreads) neighbors_csr_kernel(const CsrParams P) {
    __shared__ int s_list[kPrefCap * kPrefThreads];
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)P.n_frames * P.n_centres;
    if (g > total) return;
    if (g == total) {  // sentinel: the scan turns it into the number of pairs
        if (!FILL) P.offsets[g] = 0u;
        return;
    }
    const int f = (int)(g / P.n_centres);
    const BoxD b = load_box(P.box + (size_t)f * 3);
    const CentreRef ctr = get_centre(P.grid, P.centres, P.centre_dtype, g, f, P.n_centres, b);
    const double rx = ctr.x, ry = ctr.y, rz = ctr.z;
    const int cx = ctr.cx, cy = ctr.cy, cz = ctr.cz;
    const size_t og = ctr.out;  // where this centre's results go
    const FloatBox fb = float_box(b.Lx, b.Ly, b.Lz);
    const float thr2 = float_margin_thr2(sqrt(P.highsq), fmax(b.Lx, fmax(b.Ly, b.Lz)));
    const size_t o = FILL ? (size_t)P.offsets[og] : 0;
    const bool fits = FILL && (long long)P.offsets[og + 1] <= P.capacity;
    uint32_t K = 0;
    sweep_stencil1_pref(P.grid, f, cx, cy, cz, wrapped_coord(rx, b.Lx, b.iLx), wrapped_coord(ry, b.Ly, b.iLy),
                        wrapped_coord(rz, b.Lz, b.iLz), fb, thr2, s_list + threadIdx.x, [&](int j) {
        double px, py, pz;
        int id;
        RecTraits<double>::load(P.grid.recs, (size_t)j, px, py, pz, id);
        double dx, dy, dz;
        min_image_3(px, py, pz, rx, ry, rz, b, dx, dy, dz);
        const double s = sumsq3<double>(dx, dy, dz);
        if (s > P.lowsq && s <= P.highsq) {
            if (fits) {  // insertion by atom index into this thread's own segment
                int32_t *seg = P.indices + o;
                int at = (int)K;
                while (at > 0 && seg[at - 1] > id) {
                    seg[at] = seg[at - 1];
                    --at;
                }
                seg[at] = id;
            }
            ++K;
        }
    });
    if (!FILL) P.offsets[og] = K;
}

// ---- water orientation (watOrient) and cube/sphere occupancy (binOnGrid) ---------------------------------------------------

// AngBetween of the unit vector v / n with the unit reference (fortran/waterlib.f90:954-965): dot product of the
// normalised components summed left to right, clamped, then the same acos -> degrees chain as CosAngle3
__device__ __forceinline__ double ang_between_unit(double vx, double vy, double vz, double n, double rx, double ry, double rz) {
    const double ux = __ddiv_rn(vx, n), uy = __ddiv_rn(vy, n), uz = __ddiv_rn(vz, n);
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(ux, rx), __dmul_rn(uy, ry)), __dmul_rn(uz, rz));
    return angle_deg_from_cos(fmin(1.0, fmax(-1.0, dot)));
}

__global__ void __launch_bounds__(256) water_orient_kernel(const double *__restrict__ opos, const double *__restrict__ hpos,
                                                           const double *__restrict__ box, int n_frames, int n_waters, double rx,
                                                           double ry, double rz, double *__restrict__ angdip,
                                                           double *__restrict__ angplane) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)n_frames * n_waters) return;
    const BoxD b = load_box(box + (g / n_waters) * 3);
    const double L[3] = {b.Lx, b.Ly, b.Lz}, iL[3] = {b.iLx, b.iLy, b.iLz};
    const double *o = opos + 3 * g, *h1 = hpos + 6 * g, *h2 = hpos + 6 * g + 3;
    double v1[3], v2[3], dip[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // vecoh = hpos - opos, minimum image; the dipole (their sum) is imaged once more (waterlib.f90:998-1004)
        double t = __dsub_rn(h1[k], o[k]);
        v1[k] = __dsub_rn(t, __dmul_rn(L[k], anint_exact<double>(__dmul_rn(t, iL[k]))));
        t = __dsub_rn(h2[k], o[k]);
        v2[k] = __dsub_rn(t, __dmul_rn(L[k], anint_exact<double>(__dmul_rn(t, iL[k]))));
        t = __dadd_rn(v1[k], v2[k]);
        dip[k] = __dsub_rn(t, __dmul_rn(L[k], anint_exact<double>(__dmul_rn(t, iL[k]))));
    }
    angdip[g] = ang_between_unit(dip[0], dip[1], dip[2], __dsqrt_rn(sumsq3<double>(dip[0], dip[1], dip[2])), rx, ry, rz);
    // crossProd3 (waterlib.f90:26-28)
    const double px = __dsub_rn(__dmul_rn(v1[1], v2[2]), __dmul_rn(v1[2], v2[1]));
    const double py = __dsub_rn(__dmul_rn(v1[2], v2[0]), __dmul_rn(v1[0], v2[2]));
    const double pz = __dsub_rn(__dmul_rn(v1[0], v2[1]), __dmul_rn(v1[1], v2[0]));
    angplane[g] = ang_between_unit(px, py, pz, __dsqrt_rn(sumsq3<double>(px, py, pz)), rx, ry, rz);
}

__global__ void __launch_bounds__(256) bin_on_grid_kernel(const double *__restrict__ opos, long long n, const double *__restrict__ xb,
                                                          const double *__restrict__ yb, const double *__restrict__ zb, int nx,
                                                          int ny, int nz, double binwidth, int32_t *__restrict__ hist) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = opos[3 * i], y = opos[3 * i + 1], z = opos[3 * i + 2];
    // thisbin = floor((pos - bins(1)) / binwidth) + 1, left edge inclusive; outside the bins: not counted
    const double fx = floor(__ddiv_rn(__dsub_rn(x, xb[0]), binwidth)), fy = floor(__ddiv_rn(__dsub_rn(y, yb[0]), binwidth)),
                 fz = floor(__ddiv_rn(__dsub_rn(z, zb[0]), binwidth));
    if (!(fx >= 0.0 && fx < (double)(nx - 1)) || !(fy >= 0.0 && fy < (double)(ny - 1)) || !(fz >= 0.0 && fz < (double)(nz - 1))) return;
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    const double half = __dmul_rn(binwidth, 0.5);
    const double vx = __dsub_rn(x, __dadd_rn(xb[ix], half)), vy = __dsub_rn(y, __dadd_rn(yb[iy], half)),
                 vz = __dsub_rn(z, __dadd_rn(zb[iz], half));
    if (sumsq3<double>(vx, vy, vz) <= __ddiv_rn(__dmul_rn(binwidth, binwidth), 4.0))
        atomicAdd(hist + ((size_t)ix * (ny - 1) + iy) * (nz - 1) + iz, 1);
}

// ---- np.histogram + tetrahedral-window sums over an array of angles --------------------------------

__global__ void __launch_bounds__(256) histogram_kernel(const double *__restrict__ x, size_t n, double lo, double hi, int nbins,
                                                        unsigned long long *__restrict__ hist, double tet_lo, double tet_hi,
                                                        double *__restrict__ tet) {
    extern __shared__ unsigned s_hist[];
    const bool use_smem = nbins <= kMaxSmemBins;
    if (use_smem) {
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
    const HistSpec hs = hist_spec(lo, hi, nbins);
    double cnt = 0.0, sc = 0.0, sc2 = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        const int b = hist_bin(hs, v);
        if (b >= 0) {
            if (use_smem) atomicAdd(s_hist + b, 1u);
            else atomicAdd(hist + b, 1ull);
        }
        if (v >= tet_lo && v <= tet_hi) {
            const double c = cos(__ddiv_rn(__dmul_rn(v, 3.141592653589793), 180.0));
            cnt += 1.0;
            sc += c;
            sc2 += c * c;
        }
    }
    cnt = warp_sum(cnt);
    sc = warp_sum(sc);
    sc2 = warp_sum(sc2);
    if (tet && (threadIdx.x & 31) == 0 && cnt != 0.0) {
        atomicAdd(tet + 0, cnt);
        atomicAdd(tet + 1, sc);
        atomicAdd(tet + 2, sc2);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += blockDim.x)
            if (s_hist[i]) atomicAdd(hist + i, (unsigned long long)s_hist[i]);
    }
}

// ---- dense neighbour matrix, reimage, tetracosang, lsidists (f2py-compatible small-array routines) ---

__global__ void neighbor_matrix_kernel(const void *sub, int sub_dtype, int m, const void *pos, int pos_dtype, int n,
                                       const double *box, double lowsq, double highsq, int32_t *out) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)m * n) return;
    const int i = (int)(g / n), j = (int)(g - (size_t)i * n);
    const BoxD b = load_box(box);
    double rx, ry, rz, px, py, pz;
    load3<double>(sub, sub_dtype, i, rx, ry, rz);
    load3<double>(pos, pos_dtype, j, px, py, pz);
    double dx, dy, dz;
    min_image_3(px, py, pz, rx, ry, rz, b, dx, dy, dz);
    const double s = sumsq3<double>(dx, dy, dz);
    out[g] = (s > lowsq && s <= highsq) ? 1 : 0;
}

// mode 0: reimaged positions (n,3); mode 1: lsiDists distances (n)
__global__ void reimage_kernel(const double *pos, int n, const double *ref, const double *box, double *out, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BoxD b = load_box(box);
    const double dx = min_image_1<double, true>(pos[3 * i + 0], ref[0], b.Lx, b.iLx);
    const double dy = min_image_1<double, true>(pos[3 * i + 1], ref[1], b.Ly, b.iLy);
    const double dz = min_image_1<double, true>(pos[3 * i + 2], ref[2], b.Lz, b.iLz);
    if (mode == 0) {
        out[3 * i + 0] = __dadd_rn(ref[0], dx);
        out[3 * i + 1] = __dadd_rn(ref[1], dy);
        out[3 * i + 2] = __dadd_rn(ref[2], dz);
    } else {
        out[i] = __dsqrt_rn(sumsq3<double>(dx, dy, dz));
    }
}

__global__ void tetracosang_kernel(const double *ref, const double *neigh, int k, const double *box, double *out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= k * k) return;
    const int a = g / k, c = g - a * k;
    if (a == c) {
        out[g] = 0.0;  // the Fortran leaves the diagonal unwritten; f2py users see zeros on fresh pages
        return;
    }
    const BoxD b = load_box(box);
    double v[2][3], nrm[2];
    const int who[2] = {a, c};
    for (int t = 0; t < 2; ++t) {
        const double *p = neigh + 3 * who[t];
        const double dx = min_image_1<double, true>(p[0], ref[0], b.Lx, b.iLx);
        const double dy = min_image_1<double, true>(p[1], ref[1], b.Ly, b.iLy);
        const double dz = min_image_1<double, true>(p[2], ref[2], b.Lz, b.iLz);
        v[t][0] = __dsub_rn(__dadd_rn(ref[0], dx), ref[0]);
        v[t][1] = __dsub_rn(__dadd_rn(ref[1], dy), ref[1]);
        v[t][2] = __dsub_rn(__dadd_rn(ref[2], dz), ref[2]);
        nrm[t] = sumsq3<double>(v[t][0], v[t][1], v[t][2]);
    }
    if (nrm[0] == 0.0 || nrm[1] == 0.0) {
        out[g] = 0.0;
        return;
    }
    out[g] = angle_deg_from_cos(clamped_cos<double>(dot3<double>(v[0][0], v[0][1], v[0][2], v[1][0], v[1][1], v[1][2]), nrm[0], nrm[1]));
}

// ---- hydrogen bonds (generalHbonds) -------------------------------------------------------------------

struct HbParams {
    CellGrid grid;  // cell list over the donor heavy atoms
    const double *box;
    const void *acc;
    int acc_dtype;
    const void *donh;
    int donh_dtype;
    int n_frames, n_acc, n_don;
    double cutsq, tiny;
    double cos_thr;        // bond <=> clamped cosine <= cos_thr (host bisection of AngBetween, wol_angle_threshold)
    int minus_one_bonds;   // whether the -180 degrees AngBetween returns for cosine -1 passes angCut
    int32_t *acc_count;    // [F][n_acc]
    int32_t *don_count;    // [F][n_don]  (atomic)
    int32_t *dense;        // optional [F][n_acc][n_don]
    int2 *pairs;           // optional (acceptor, donor) list
    uint32_t pair_capacity;
    uint32_t *pair_counter;
};

__global__ void __launch_bounds__(kPrefThreads) hbond_kernel(const HbParams P) {
    __shared__ int s_list[kPrefCap * kPrefThreads];
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)P.n_frames * P.n_acc) return;
    const int f = (int)(g / P.n_acc);
    const int i = (int)(g - (size_t)f * P.n_acc);
    const BoxD b = load_box(P.box + (size_t)f * 3);
    double ax, ay, az;
    load3<double>(P.acc, P.acc_dtype, g, ax, ay, az);
    const int cx = cell_coord(ax, b.iLx, P.grid.nc0), cy = cell_coord(ay, b.iLy, P.grid.nc1),
              cz = cell_coord(az, b.iLz, P.grid.nc2);
    int count = 0;
    const FloatBox fb = float_box(b.Lx, b.Ly, b.Lz);
    const float thr2 = float_margin_thr2(sqrt(P.cutsq), fmax(b.Lx, fmax(b.Ly, b.Lz)));
    sweep_stencil1_pref(P.grid, f, cx, cy, cz, wrapped_coord(ax, b.Lx, b.iLx), wrapped_coord(ay, b.Ly, b.iLy),
                        wrapped_coord(az, b.Lz, b.iLz), fb, thr2, s_list + threadIdx.x, [&](int j) {
        double dxp, dyp, dzp;
        int jd;
        RecTraits<double>::load(P.grid.recs, (size_t)j, dxp, dyp, dzp, jd);
        // distvec = donor - acceptor, reimaged; reject distSq > distCut^2 or distSq <= 1.0E-2 (:1184-1188)
        double ddx, ddy, ddz;
        min_image_3(dxp, dyp, dzp, ax, ay, az, b, ddx, ddy, ddz);
        const double s = sumsq3<double>(ddx, ddy, ddz);
        if (!(s > P.tiny && s <= P.cutsq)) return;
        double hx, hy, hz;
        load3<double>(P.donh, P.donh_dtype, (size_t)f * P.n_don + jd, hx, hy, hz);
        // unit vectors from the hydrogen to the acceptor and to the donor, each reimaged (:1190-1198)
        double v1x, v1y, v1z;
        min_image_3(ax, ay, az, hx, hy, hz, b, v1x, v1y, v1z);
        const double n1 = __dsqrt_rn(sumsq3<double>(v1x, v1y, v1z));
        v1x = __ddiv_rn(v1x, n1); v1y = __ddiv_rn(v1y, n1); v1z = __ddiv_rn(v1z, n1);
        double v2x, v2y, v2z;
        min_image_3(dxp, dyp, dzp, hx, hy, hz, b, v2x, v2y, v2z);
        const double n2 = __dsqrt_rn(sumsq3<double>(v2x, v2y, v2z));
        v2x = __ddiv_rn(v2x, n2); v2y = __ddiv_rn(v2y, n2); v2z = __ddiv_rn(v2z, n2);
        const double c = fmin(1.0, fmax(-1.0, dot3<double>(v1x, v1y, v1z, v2x, v2y, v2z)));
        const bool bond = (c == -1.0) ? (P.minus_one_bonds != 0) : (c <= P.cos_thr);
        if (!bond) return;
        ++count;
        if (P.don_count) atomicAdd(P.don_count + (size_t)f * P.n_don + jd, 1);
        if (P.dense) P.dense[((size_t)f * P.n_acc + i) * P.n_don + jd] = 1;
        if (P.pairs) {
            const uint32_t at = atomicAdd(P.pair_counter, 1u);
            if (at < P.pair_capacity) P.pairs[at] = make_int2(f * P.n_acc + i, jd);
        }
    });
    if (P.acc_count) P.acc_count[g] = count;
}

// H-bond locations of HBondsGeneral (structureLibs/water_properties.py:709-714): halfway between the
// acceptor and the donor hydrogen imaged around it.
__global__ void hbond_loc_kernel(const int2 *__restrict__ pairs, int n_pairs, const void *acc, int acc_dtype, int n_acc,
                                 const void *donh, int donh_dtype, int n_don, const double *box, double *__restrict__ out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pairs) return;
    const int2 pr = pairs[g];
    const int f = pr.x / n_acc;
    const BoxD b = load_box(box + (size_t)f * 3);
    double ax, ay, az, hx, hy, hz;
    load3<double>(acc, acc_dtype, (size_t)pr.x, ax, ay, az);
    load3<double>(donh, donh_dtype, (size_t)f * n_don + pr.y, hx, hy, hz);
    const double rx = __dadd_rn(ax, min_image_1<double, true>(hx, ax, b.Lx, b.iLx));
    const double ry = __dadd_rn(ay, min_image_1<double, true>(hy, ay, b.Ly, b.iLy));
    const double rz = __dadd_rn(az, min_image_1<double, true>(hz, az, b.Lz, b.iLz));
    out[3 * (size_t)g + 0] = __dmul_rn(0.5, __dadd_rn(rx, ax));
    out[3 * (size_t)g + 1] = __dmul_rn(0.5, __dadd_rn(ry, ay));
    out[3 * (size_t)g + 2] = __dmul_rn(0.5, __dadd_rn(rz, az));
}

// ---- hydration shell: atoms of the cell list within (low, cutoff] of any solute atom --------------------

struct ShellParams {
    CellGrid grid;  // cell list over the waters
    const double *box;
    const void *sol;
    int sol_dtype;
    int n_frames, n_sol, n_pos;
    double lowsq, highsq;
    int32_t *mask;  // [F][n_pos], set to 1 (caller zero-fills)
};

__global__ void __launch_bounds__(kPrefThreads) shell_kernel(const ShellParams P) {
    __shared__ int s_list[kPrefCap * kPrefThreads];
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)P.n_frames * P.n_sol) return;
    const int f = (int)(g / P.n_sol);
    const BoxD b = load_box(P.box + (size_t)f * 3);
    double sx, sy, sz;
    load3<double>(P.sol, P.sol_dtype, g, sx, sy, sz);
    const int cx = cell_coord(sx, b.iLx, P.grid.nc0), cy = cell_coord(sy, b.iLy, P.grid.nc1),
              cz = cell_coord(sz, b.iLz, P.grid.nc2);
    const FloatBox fb = float_box(b.Lx, b.Ly, b.Lz);
    const float thr2 = float_margin_thr2(sqrt(P.highsq), fmax(b.Lx, fmax(b.Ly, b.Lz)));
    sweep_stencil1_pref(P.grid, f, cx, cy, cz, wrapped_coord(sx, b.Lx, b.iLx), wrapped_coord(sy, b.Ly, b.iLy),
                        wrapped_coord(sz, b.Lz, b.iLz), fb, thr2, s_list + threadIdx.x, [&](int j) {
        double px, py, pz;
        int id;
        RecTraits<double>::load(P.grid.recs, (size_t)j, px, py, pz, id);
        double dx, dy, dz;
        min_image_3(px, py, pz, sx, sy, sz, b, dx, dy, dz);
        const double s = sumsq3<double>(dx, dy, dz);
        if (s > P.lowsq && s <= P.highsq) P.mask[(size_t)f * P.n_pos + id] = 1;
    });
}

// ---- local structure index (getLSI, structureLibs/water_properties.py:252-311) -------------------------------
// One thread per centre over a cell list of edge >= highCut + 3.7: the minimum-image distances of the neighbours
// inside (lowCut, highCut] go to a small sorted list; of the atoms in the next shell (highCut, highCut + 3.7] the one
// with the smallest NON-periodic distance to the centre (the reference compares raw coordinate differences there,
// :289; first atom index on ties) contributes its minimum-image distance too; the LSI is the population variance
// of the gaps between consecutive sorted distances.

constexpr int kLsiCap = 48;

struct LsiParams {
    CellGrid grid;
    const double *box;
    const void *centres;
    int centre_dtype;
    int n_frames, n_pos, n_centres;
    double lowsq, highsq, nextsq;
    double *lsi;       // [F][M], written where has == 1 (0 elsewhere)
    int32_t *num;      // [F][M] number of gaps (= neighbours inside highCut), 0 when the centre has no value
    uint32_t *counters;
};

__global__ void __launch_bounds__(128) lsi_kernel(const LsiParams P) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)P.n_frames * P.n_centres) return;
    const int f = (int)(g / P.n_centres);
    const BoxD b = load_box(P.box + (size_t)f * 3);
    const CentreRef ctr = get_centre(P.grid, P.centres, P.centre_dtype, g, f, P.n_centres, b);
    const double rx = ctr.x, ry = ctr.y, rz = ctr.z;
    const int cx = ctr.cx, cy = ctr.cy, cz = ctr.cz;
    const size_t og = ctr.out;  // where this centre's results go
    double dist[kLsiCap];
    dist[0] = 0.0;
    int k = 0, n_next = 0, next_idx = 0;
    double next_raw = 0.0, next_min = 0.0;
    bool over = false;
    // no float prefilter here: a sixth of the stencil lies inside the reach, so every lane does exact work most of the
    // time anyway and a uniform exact sweep beats a divergent two-step one (measured: 7 ms vs 100 ms per 1M waters)
    sweep_stencil1(P.grid, f, cx, cy, cz, [&](int j) {
        double px, py, pz;
        int id;
        RecTraits<double>::load(P.grid.recs, (size_t)j, px, py, pz, id);
        double dx, dy, dz;
        min_image_3(px, py, pz, rx, ry, rz, b, dx, dy, dz);
        const double s = sumsq3<double>(dx, dy, dz);
        if (s > P.lowsq && s <= P.highsq) {
            if (k >= kLsiCap - 1) {
                over = true;
                return;
            }
            // insertion into the ascending list (np.sort of lsidists, :300)
            const double d = __dsqrt_rn(s);
            int at = k;
            while (at > 0 && dist[at - 1] > d) {
                dist[at] = dist[at - 1];
                --at;
            }
            dist[at] = d;
            ++k;
        } else if (s > P.highsq && s <= P.nextsq) {
            // np.sqrt(np.sum((Pos[next] - apos)**2.0, axis=1)): raw differences, no minimum image (:287-289)
            const double raw = __dsqrt_rn(sumsq3<double>(__dsub_rn(px, rx), __dsub_rn(py, ry), __dsub_rn(pz, rz)));
            if (n_next == 0 || raw < next_raw || (raw == next_raw && id < next_idx)) {  // np.argmin: first index on ties
                next_raw = raw;
                next_idx = id;
                next_min = __dsqrt_rn(s);
            }
            ++n_next;
        }
    });
    if (over) {
        atomicAdd(P.counters + kCntFatal, 1u);
        return;
    }
    double val = 0.0;
    int nd = 0;
    if (k > 1 && n_next > 0) {
        int at = k;
        while (at > 0 && dist[at - 1] > next_min) {
            dist[at] = dist[at - 1];
            --at;
        }
        dist[at] = next_min;
        ++k;
        nd = k - 1;
        double mean = 0.0;
        for (int t = 0; t < nd; ++t) mean += dist[t + 1] - dist[t];
        mean /= (double)nd;
        for (int t = 0; t < nd; ++t) {
            const double x = (dist[t + 1] - dist[t]) - mean;
            val += x * x;
        }
        val /= (double)nd;
    }
    P.lsi[og] = val;
    P.num[og] = nd;
}

static CellGrid make_grid(void *workspace, const WorkspaceLayout &lay, const int32_t nc[3]) {
    char *ws = reinterpret_cast<char *>(workspace);
    CellGrid g;
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

extern "C" {

int wol_angle_offsets(const int32_t *n3, int64_t n, uint32_t *offsets, uint32_t *scratch, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || !offsets || !scratch || (!n3 && n > 0)) return set_error(WOL_ERR_INVALID, "wol_angle_offsets: null argument");
    if (n >= (1LL << 26)) return set_error(WOL_ERR_RANGE, "wol_angle_offsets: more than 2^26 centres; materialise in batches");
    const size_t cnt = (size_t)n + 1;
    pair_counts_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(n3, (size_t)n, offsets);
    add_launches(1);
    const int rc = scan_offsets_checked(offsets, cnt, scratch, stream);
    if (rc != WOL_OK) return rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_angle_offsets", e);
    return WOL_OK;
}

int wol_angles_fill(const void *centres, int32_t centre_dtype, const double *box, int32_t n_frames, int32_t n_pos,
                    int32_t n_centres, const int32_t nc[3], double edge_min, double low3, double high3, void *workspace,
                    size_t workspace_bytes, const uint32_t *offsets, double *angles, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!box || !nc || !workspace || !offsets || (!angles)) return set_error(WOL_ERR_INVALID, "wol_angles_fill: null argument");
    if (!centres && n_centres != n_pos) return set_error(WOL_ERR_INVALID, "wol_angles_fill: centres == NULL means every atom is a centre (n_centres == n_pos)");
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && high3 * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "cutoff %.6g exceeds the planned cell edge %.6g", high3, edge_min);
    const WorkspaceLayout lay = workspace_layout(n_frames, n_pos, n_centres, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    MatParams P;
    P.grid = make_grid(workspace, lay, nc);
    P.box = box;
    P.centres = centres;
    P.centre_dtype = centre_dtype;
    P.n_frames = n_frames;
    P.n_pos = n_pos;
    P.n_centres = n_centres;
    P.lowsq = low3 * low3;
    P.highsq = high3 * high3;
    P.offsets = offsets;
    P.angles = angles;
    P.counters = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(workspace) + lay.off_counters);
    const size_t total = (size_t)n_frames * n_centres;
    if (total > 0) {
        angles_fill_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_angles_fill", e);
    return WOL_OK;
}

int wol_neighbors_csr(const void *centres, int32_t centre_dtype, const double *box, int32_t n_frames, int32_t n_pos,
                      int32_t n_centres, const int32_t nc[3], double edge_min, double lowcut, double highcut, void *workspace,
                      size_t workspace_bytes, uint32_t *offsets, uint32_t *scratch, int32_t *indices, int64_t capacity,
                      void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!box || !nc || !workspace || !offsets || !scratch || capacity < 0 || (capacity > 0 && !indices))
        return set_error(WOL_ERR_INVALID, "wol_neighbors_csr: bad argument");
    if (!centres && n_centres != n_pos) return set_error(WOL_ERR_INVALID, "wol_neighbors_csr: centres == NULL means every atom is a centre (n_centres == n_pos)");
    if (n_frames < 1 || n_pos < 0 || n_centres < 0) return set_error(WOL_ERR_INVALID, "wol_neighbors_csr: negative size");
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && highcut * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "cutoff %.6g exceeds the planned cell edge %.6g", highcut, edge_min);
    const size_t total = (size_t)n_frames * n_centres;
    if (total >= (1ull << 26)) return set_error(WOL_ERR_RANGE, "wol_neighbors_csr: more than 2^26 centres; use smaller batches");
    const WorkspaceLayout lay = workspace_layout(n_frames, n_pos, n_centres, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    CsrParams P;
    P.grid = make_grid(workspace, lay, nc);
    P.box = box;
    P.centres = centres;
    P.centre_dtype = centre_dtype;
    P.n_frames = n_frames;
    P.n_pos = n_pos;
    P.n_centres = n_centres;
    P.lowsq = lowcut * lowcut;
    P.highsq = highcut * highcut;
    P.offsets = offsets;
    P.indices = indices;
    P.capacity = capacity;
    const unsigned blocks = (unsigned)((total + 1 + kPrefThreads - 1) / kPrefThreads);
    neighbors_csr_kernel<false><<<blocks, kPrefThreads, 0, stream>>>(P);
    const int rc = scan_offsets_checked(offsets, total + 1, scratch, stream);
    if (rc != WOL_OK) return rc;
    add_launches(1);
    if (capacity > 0 && total > 0) {
        neighbors_csr_kernel<true><<<blocks, kPrefThreads, 0, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_neighbors_csr", e);
    return WOL_OK;
}

int wol_water_orient(const double *opos, const double *hpos, const double *box, int32_t n_frames, int32_t n_waters,
                     const double refvec_host[3], double *angdip, double *angplane, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!box || !refvec_host || !angdip || !angplane || n_frames < 1 || n_waters < 0 || ((!opos || !hpos) && n_waters > 0))
        return set_error(WOL_ERR_INVALID, "wol_water_orient: bad argument");
    // refvecnorm = refvec / sqrt(sum(refvec * refvec))  (waterlib.f90:993)
    const double rn = sqrt((refvec_host[0] * refvec_host[0] + refvec_host[1] * refvec_host[1]) + refvec_host[2] * refvec_host[2]);
    const size_t total = (size_t)n_frames * n_waters;
    if (total > 0) {
        water_orient_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(opos, hpos, box, n_frames, n_waters, refvec_host[0] / rn,
                                                                             refvec_host[1] / rn, refvec_host[2] / rn, angdip, angplane);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_water_orient", e);
    return WOL_OK;
}

int wol_bin_on_grid(const double *opos, int64_t n, const double *xbins, const double *ybins, const double *zbins, int32_t nx, int32_t ny,
                    int32_t nz, double binwidth, int32_t *outhist, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!xbins || !ybins || !zbins || !outhist || nx < 2 || ny < 2 || nz < 2 || n < 0 || (!opos && n > 0) || !(binwidth > 0.0))
        return set_error(WOL_ERR_INVALID, "wol_bin_on_grid: bad argument");
    cudaError_t e = cudaMemsetAsync(outhist, 0, sizeof(int32_t) * (size_t)(nx - 1) * (ny - 1) * (nz - 1), stream);  // outhist = 0 (:1064)
    if (e != cudaSuccess) return set_cuda_error("wol_bin_on_grid", e);
    if (n > 0) {
        bin_on_grid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(opos, (long long)n, xbins, ybins, zbins, nx, ny, nz, binwidth,
                                                                          outhist);
        add_launches(1);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_bin_on_grid", e);
    return WOL_OK;
}

int wol_histogram(const double *x, int64_t n, double lo, double hi, int32_t nbins, int64_t *hist, double tet_lo,
                  double tet_hi, double *tet_sums, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || nbins < 1 || !hist || !(hi > lo) || (!x && n > 0)) return set_error(WOL_ERR_INVALID, "wol_histogram: bad argument");
    if (n == 0) return WOL_OK;
    const size_t smem = nbins <= kMaxSmemBins ? sizeof(unsigned) * nbins : 0;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    histogram_kernel<<<(unsigned)blocks, 256, smem, stream>>>(x, (size_t)n, lo, hi, nbins, reinterpret_cast<unsigned long long *>(hist),
                                                              tet_lo, tet_hi, tet_sums);
    add_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_histogram", e);
    return WOL_OK;
}

int wol_neighbor_matrix(const void *sub, int32_t sub_dtype, int32_t m, const void *pos, int32_t pos_dtype, int32_t n,
                        const double *box, double lowcut, double highcut, int32_t *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 0 || n < 0 || !box || !out || !sub || !pos) return set_error(WOL_ERR_INVALID, "wol_neighbor_matrix: bad argument");
    const size_t total = (size_t)m * n;
    if (total >= (1ULL << 40)) return set_error(WOL_ERR_RANGE, "dense neighbour matrix too large");
    if (total > 0) {
        neighbor_matrix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(sub, sub_dtype, m, pos, pos_dtype, n, box,
                                                                                  lowcut * lowcut, highcut * highcut, out);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_neighbor_matrix", e);
    return WOL_OK;
}

int wol_reimage(const double *pos, int32_t n, const double *ref, const double *box, double *out, int32_t mode, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || !ref || !box || !out || (!pos && n > 0) || (mode != 0 && mode != 1)) return set_error(WOL_ERR_INVALID, "wol_reimage: bad argument");
    if (n > 0) {
        reimage_kernel<<<(n + 127) / 128, 128, 0, stream>>>(pos, n, ref, box, out, mode);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_reimage", e);
    return WOL_OK;
}

int wol_tetracosang(const double *ref, const double *neigh, int32_t k, const double *box, double *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 0 || !ref || !box || !out || (!neigh && k > 0)) return set_error(WOL_ERR_INVALID, "wol_tetracosang: bad argument");
    if (k > 0) {
        tetracosang_kernel<<<(k * k + 127) / 128, 128, 0, stream>>>(ref, neigh, k, box, out);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_tetracosang", e);
    return WOL_OK;
}

int wol_lsi(const void *centres, int32_t centre_dtype, const double *box, int32_t n_frames, int32_t n_pos, int32_t n_centres,
            const int32_t nc[3], double edge_min, double lowcut, double highcut, void *workspace, size_t workspace_bytes,
            double *lsi, int32_t *num, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!box || !nc || !workspace || !lsi || !num) return set_error(WOL_ERR_INVALID, "wol_lsi: null argument");
    if (!centres && n_centres != n_pos) return set_error(WOL_ERR_INVALID, "wol_lsi: centres == NULL means every atom is a centre (n_centres == n_pos)");
    const double reach = highcut + 3.7;  // width of the next-neighbour shell, hard-coded in the reference (:269, :274)
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && reach * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "LSI search radius %.6g exceeds the planned cell edge %.6g", reach, edge_min);
    const WorkspaceLayout lay = workspace_layout(n_frames, n_pos, n_centres, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    LsiParams P;
    P.grid = make_grid(workspace, lay, nc);
    P.box = box;
    P.centres = centres;
    P.centre_dtype = centre_dtype;
    P.n_frames = n_frames;
    P.n_pos = n_pos;
    P.n_centres = n_centres;
    P.lowsq = lowcut * lowcut;
    P.highsq = highcut * highcut;
    P.nextsq = reach * reach;
    P.lsi = lsi;
    P.num = num;
    P.counters = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(workspace) + lay.off_counters);
    const size_t total = (size_t)n_frames * n_centres;
    if (total > 0 && n_pos > 0) {
        lsi_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_lsi", e);
    return WOL_OK;
}

int wol_hbond_counts(const wol_hbond_args *a, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a || a->struct_size != sizeof(wol_hbond_args)) return set_error(WOL_ERR_INVALID, "wol_hbond_counts: bad args struct");
    if (!a->acc || !a->donh || !a->box || !a->workspace) return set_error(WOL_ERR_INVALID, "wol_hbond_counts: null argument");
    if (a->n_frames < 1 || a->n_acc < 0 || a->n_don < 0) return set_error(WOL_ERR_INVALID, "wol_hbond_counts: negative size");
    for (int k = 0; k < 3; ++k)
        if (a->nc[k] > 3 && a->dist_cut * (1.0 + 1e-9) > a->edge_min)
            return set_error(WOL_ERR_INVALID, "H-bond distance cutoff %.6g exceeds the planned cell edge %.6g", a->dist_cut, a->edge_min);
    const WorkspaceLayout lay = workspace_layout(a->n_frames, a->n_don, a->n_don, a->nc);
    if (a->workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", a->workspace_bytes, lay.total);
    HbParams P;
    P.grid = make_grid(a->workspace, lay, a->nc);
    P.box = a->box;
    P.acc = a->acc;
    P.acc_dtype = a->acc_dtype;
    P.donh = a->donh;
    P.donh_dtype = a->donh_dtype;
    P.n_frames = a->n_frames;
    P.n_acc = a->n_acc;
    P.n_don = a->n_don;
    P.cutsq = a->dist_cut * a->dist_cut;
    P.tiny = (double)1.0e-2f;  // the Fortran literal 1.0E-2 is single precision (waterlib.f90:1187)
    int m1 = 0;
    P.cos_thr = angle_cos_threshold(a->ang_cut, &m1);
    P.minus_one_bonds = m1;
    P.acc_count = a->acc_count;
    P.don_count = a->don_count;
    P.dense = a->dense;
    P.pairs = reinterpret_cast<int2 *>(a->pairs);
    P.pair_capacity = a->pair_capacity;
    P.pair_counter = a->pair_counter;
    const size_t total = (size_t)a->n_frames * a->n_acc;
    if (total > 0 && a->n_don > 0) {
        hbond_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_hbond_counts", e);
    return WOL_OK;
}

int wol_hbond_locations(const int32_t *pairs, int32_t n_pairs, const void *acc, int32_t acc_dtype, int32_t n_acc,
                        const void *donh, int32_t donh_dtype, int32_t n_don, const double *box, double *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pairs < 0 || n_acc < 1 || n_don < 1 || !acc || !donh || !box || !out || (!pairs && n_pairs > 0))
        return set_error(WOL_ERR_INVALID, "wol_hbond_locations: bad argument");
    if (n_pairs > 0) {
        hbond_loc_kernel<<<(n_pairs + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const int2 *>(pairs), n_pairs, acc, acc_dtype,
                                                                    n_acc, donh, donh_dtype, n_don, box, out);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_hbond_locations", e);
    return WOL_OK;
}

int wol_shell_mask(const void *sol, int32_t sol_dtype, int32_t n_sol, const double *box, int32_t n_frames, int32_t n_pos,
                   const int32_t nc[3], double edge_min, double lowcut, double cutoff, void *workspace, size_t workspace_bytes,
                   int32_t *mask, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!sol || !box || !nc || !workspace || !mask) return set_error(WOL_ERR_INVALID, "wol_shell_mask: null argument");
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && cutoff * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "shell cutoff %.6g exceeds the planned cell edge %.6g", cutoff, edge_min);
    const WorkspaceLayout lay = workspace_layout(n_frames, n_pos, n_pos, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    ShellParams P;
    P.grid = make_grid(workspace, lay, nc);
    P.box = box;
    P.sol = sol;
    P.sol_dtype = sol_dtype;
    P.n_frames = n_frames;
    P.n_sol = n_sol;
    P.n_pos = n_pos;
    P.lowsq = lowcut * lowcut;
    P.highsq = cutoff * cutoff;
    P.mask = mask;
    const size_t total = (size_t)n_frames * n_sol;
    if (total > 0 && n_pos > 0) {
        shell_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_shell_mask", e);
    return WOL_OK;
}

}  // extern "C"
