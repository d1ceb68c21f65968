// Slab / interface routines (BASELINE config 4) and the all-Fortran triplet histogram, sm_100a, fp64 in the
// reference's operation order:
//   willard_kernel        WillardDensityField / WillardDensityPoints   fortran/waterlib.f90:1286-1341, :1351-1398
//   voxel_density_kernel  DensityField                                   fortran/waterlib.f90:1219-1268
//   iface_nearest_kernel  InterfaceWater: nearest surface point per water (+ signed depth) and nearest water per
//                         surface point                                 fortran/waterlib.f90:1431-1468
//   profile_kernel        depth-binned profile of a per-water observable (the cfg-4 composition; the reference
//                         has the ingredients, structureLibs/surface_library.py:170-210, but no such function)
//   histrr3b_kernel       histrr3b                                        fortran/waterlib.f90:1550-1593
#include "wol_mc_table.h"
#include <math.h>

#include "wol_q3b_common.cuh"

namespace wol {

struct CellGridS {
    const uint32_t *cell_start;
    const void *recs;
    int nc0, nc1, nc2;
};

struct Box3 {
    double L[3], iL[3];
};
// iBoxL = merge(1/BoxL, 0, BoxL >= 0)  (waterlib.f90:41)
__device__ __forceinline__ Box3 load_box3(const double *b) {
    Box3 o;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.L[k] = b[k];
        o.iL[k] = (b[k] >= 0.0) ? __ddiv_rn(1.0, b[k]) : 0.0;
    }
    return o;
}

// ---- Willard-Chandler density ------------------------------------------------------------------------

struct WillardParams {
    CellGridS grid;  // cell list over the waters, cell edge >= 3 smoothlen
    const double *box;
    const double *pts;                 // explicit points [n][3], or nullptr: the grid below
    const double *gx, *gy, *gz;        // grid axes
    int nx, ny, nz;
    long long n_points;
    double s2, pref, shiftterm, cut;   // smoothlen^2, (2 pi s2)^1.5, exp(-4.5)/pref, 9 smoothlen^2
    double *dens;                      // [n_points]
    double *norms;                     // [n_points][3]
};

__global__ void __launch_bounds__(128) willard_kernel(const WillardParams P) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.n_points) return;
    double ax, ay, az;
    if (P.pts) {
        ax = P.pts[3 * g + 0]; ay = P.pts[3 * g + 1]; az = P.pts[3 * g + 2];
    } else {
        const int k = (int)(g % P.nz), j = (int)((g / P.nz) % P.ny), i = (int)(g / ((long long)P.nz * P.ny));
        ax = P.gx[i]; ay = P.gy[j]; az = P.gz[k];
    }
    const Box3 b = load_box3(P.box);
    const int nc0 = P.grid.nc0, nc1 = P.grid.nc1, nc2 = P.grid.nc2;
    const int cx = cell_coord(ax, b.iL[0], nc0), cy = cell_coord(ay, b.iL[1], nc1), cz = cell_coord(az, b.iL[2], nc2);
    const int cntx = min(3, nc0), cnty = min(3, nc1), cntz = min(3, nc2);
    const int xs = (nc0 <= 3) ? 0 : (cx - 1 + nc0) % nc0, ys = (nc1 <= 3) ? 0 : (cy - 1 + nc1) % nc1,
              zs = (nc2 <= 3) ? 0 : (cz - 1 + nc2) % nc2;
    double dens = 0.0, nvx = 0.0, nvy = 0.0, nvz = 0.0;
    for (int iz = 0; iz < cntz; ++iz) {
        const int z = (zs + iz) % nc2;
        for (int iy = 0; iy < cnty; ++iy) {
            const int y = (ys + iy) % nc1;
            for (int ix = 0; ix < cntx; ++ix) {
                const int x = (xs + ix) % nc0;
                const size_t c = ((size_t)z * nc1 + y) * nc0 + x;
                const int j1 = (int)__ldg(P.grid.cell_start + c + 1);
                for (int j = (int)__ldg(P.grid.cell_start + c); j < j1; ++j) {
                    double px, py, pz;
                    int id;
                    RecTraits<double>::load(P.grid.recs, (size_t)j, px, py, pz, id);
                    // thisvec = apos - watpos, minimum image (waterlib.f90:1313-1314)
                    const double vx = min_image_1<double, true>(ax, px, b.L[0], b.iL[0]);
                    const double vy = min_image_1<double, true>(ay, py, b.L[1], b.iL[1]);
                    const double vz = min_image_1<double, true>(az, pz, b.L[2], b.iL[2]);
                    const double r2 = sumsq3<double>(vx, vy, vz);
                    if (r2 >= P.cut) continue;
                    const double expterm = __ddiv_rn(-r2, __dmul_rn(2.0, P.s2));
                    const double densfunc = __dsub_rn(__ddiv_rn(exp(expterm), P.pref), P.shiftterm);
                    const double w = __dadd_rn(densfunc, P.shiftterm);
                    nvx = __dadd_rn(nvx, __ddiv_rn(__dmul_rn(-vx, w), P.s2));
                    nvy = __dadd_rn(nvy, __ddiv_rn(__dmul_rn(-vy, w), P.s2));
                    nvz = __dadd_rn(nvz, __ddiv_rn(__dmul_rn(-vz, w), P.s2));
                    dens = __dadd_rn(dens, densfunc);
                }
            }
        }
    }
    P.dens[g] = dens;
    if (P.norms) {
        const double nn = __dsqrt_rn(sumsq3<double>(nvx, nvy, nvz));  // 0/0 = NaN far from every water, as in the reference
        P.norms[3 * g + 0] = __ddiv_rn(nvx, nn);
        P.norms[3 * g + 1] = __ddiv_rn(nvy, nn);
        P.norms[3 * g + 2] = __ddiv_rn(nvz, nn);
    }
}

// ---- DensityField: waters inside the cube of edge binwidth around each grid point ------------------------------

struct VoxelParams {
    CellGridS grid;  // cell list over the waters, cell edge >= binwidth / 2
    const double *box;
    const double *gx, *gy, *gz;
    int nx, ny, nz;
    double half, vol;  // binwidth / 2, binwidth ** 3.0
    double *dens;
};

__global__ void __launch_bounds__(128) voxel_density_kernel(const VoxelParams P) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)P.nx * P.ny * P.nz) return;
    const int k = (int)(g % P.nz), j = (int)((g / P.nz) % P.ny), i = (int)(g / ((long long)P.nz * P.ny));
    const double ax = P.gx[i], ay = P.gy[j], az = P.gz[k];
    const Box3 b = load_box3(P.box);
    const int nc0 = P.grid.nc0, nc1 = P.grid.nc1, nc2 = P.grid.nc2;
    const int cx = cell_coord(ax, b.iL[0], nc0), cy = cell_coord(ay, b.iL[1], nc1), cz = cell_coord(az, b.iL[2], nc2);
    const int cntx = min(3, nc0), cnty = min(3, nc1), cntz = min(3, nc2);
    const int xs = (nc0 <= 3) ? 0 : (cx - 1 + nc0) % nc0, ys = (nc1 <= 3) ? 0 : (cy - 1 + nc1) % nc1,
              zs = (nc2 <= 3) ? 0 : (cz - 1 + nc2) % nc2;
    const double lox = __dsub_rn(ax, P.half), hix = __dadd_rn(ax, P.half), loy = __dsub_rn(ay, P.half), hiy = __dadd_rn(ay, P.half),
                 loz = __dsub_rn(az, P.half), hiz = __dadd_rn(az, P.half);
    double dens = 0.0;
    for (int iz = 0; iz < cntz; ++iz)
        for (int iy = 0; iy < cnty; ++iy)
            for (int ix = 0; ix < cntx; ++ix) {
                const size_t c = ((size_t)((zs + iz) % nc2) * nc1 + (ys + iy) % nc1) * nc0 + (xs + ix) % nc0;
                const int j1 = (int)__ldg(P.grid.cell_start + c + 1);
                for (int jj = (int)__ldg(P.grid.cell_start + c); jj < j1; ++jj) {
                    double px, py, pz;
                    int id;
                    RecTraits<double>::load(P.grid.recs, (size_t)jj, px, py, pz, id);
                    // thisvec = watpos - apos, minimum image; watpos = apos + thisvec (:1247-1249)
                    const double wx = __dadd_rn(ax, min_image_1<double, true>(px, ax, b.L[0], b.iL[0]));
                    const double wy = __dadd_rn(ay, min_image_1<double, true>(py, ay, b.L[1], b.iL[1]));
                    const double wz = __dadd_rn(az, min_image_1<double, true>(pz, az, b.L[2], b.iL[2]));
                    if (wx < lox || wx > hix || wy < loy || wy > hiy || wz < loz || wz > hiz) continue;
                    dens += 1.0;
                }
            }
    P.dens[g] = __ddiv_rn(dens, P.vol);
}

// ---- InterfaceWater -----------------------------------------------------------------------------------
// Tiled brute force (the search radius, sqrt(1000) A, is most of a box): every thread owns one row entity and
// walks all column entities through shared-memory tiles in ascending index order.

constexpr int kIfaceThreads = 128;
constexpr int kIfaceTile = 512;

// Nearest column entity of one row entity.  A float pass over the shared-memory tile rejects almost every
// candidate (float minimum image, ~10 instructions); only a candidate whose float distance^2 is inside the
// current best (+ a margin that covers float rounding of coordinates up to ~2000 A) is re-evaluated with the
// reference's fp64 arithmetic, and only the fp64 value decides, with the strict '<' and ascending order of the
// Fortran loops -- so index ties resolve identically.
struct NearestScan {
    double best;     // exact distance^2 of the best candidate so far (1000 = none, waterlib.f90:1427,1435)
    float bound;     // float distance^2 above which a candidate certainly does not beat `best`
    int close;
    __device__ __forceinline__ void reset() {
        best = 1000.0;
        bound = 1000.0f * (1.0f + 1e-3f) + 5e-2f;
        close = -1;
    }
};

// row = the entity that owns the search (rx, ry, rz fp64 + float copy); columns come from `col` (fp64, global)
// through the float tile.  ROW_IS_WATER selects the operand order of distvec = watpos - gpos (:1440).
// A row's columns may be shared out among `step` threads (thread `first` of them takes columns first, first + step, ...).
template <bool ROW_IS_WATER>
__device__ __forceinline__ void nearest_tile(NearestScan &S, const float *__restrict__ s_t, int nt, int t0,
                                             const double *__restrict__ col, double rx, double ry, double rz, float fx, float fy,
                                             float fz, const Box3 &b, float Lxf, float Lyf, float Lzf, float iLxf, float iLyf,
                                             float iLzf, int first, int step) {
#pragma unroll 4
    for (int t = first; t < nt; t += step) {
        float dx = s_t[3 * t + 0] - fx, dy = s_t[3 * t + 1] - fy, dz = s_t[3 * t + 2] - fz;
        dx -= Lxf * rintf(dx * iLxf);
        dy -= Lyf * rintf(dy * iLyf);
        dz -= Lzf * rintf(dz * iLzf);
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (r2 <= S.bound) {
            const double *c = col + 3 * (size_t)(t0 + t);
            double ex, ey, ez;
            if (ROW_IS_WATER) {  // distvec = watpos - gpos
                ex = min_image_1<double, true>(rx, c[0], b.L[0], b.iL[0]);
                ey = min_image_1<double, true>(ry, c[1], b.L[1], b.iL[1]);
                ez = min_image_1<double, true>(rz, c[2], b.L[2], b.iL[2]);
            } else {
                ex = min_image_1<double, true>(c[0], rx, b.L[0], b.iL[0]);
                ey = min_image_1<double, true>(c[1], ry, b.L[1], b.iL[1]);
                ez = min_image_1<double, true>(c[2], rz, b.L[2], b.iL[2]);
            }
            const double s = sumsq3<double>(ex, ey, ez);
            if (s < S.best) {
                S.best = s;
                S.close = t0 + t;
                S.bound = (float)s * (1.0f + 1e-3f) + 5e-2f;
            }
        }
    }
}

// SPLIT threads (adjacent lanes) share one row entity: each scans every SPLIT-th column and the partial results are
// merged at the end -- smallest exact distance^2, smallest index among equals, which is what the Fortran's strict '<' over
// ascending indices keeps.  SPLIT > 1 is for searches with few rows (19 602 surface points are 154 blocks of 128 threads:
// one block per SM, four warps each).
template <bool ROW_IS_WATER, int SPLIT>
__global__ void __launch_bounds__(kIfaceThreads) iface_nearest_kernel(const double *__restrict__ row, int n_row,
                                                                      const double *__restrict__ col, int n_col,
                                                                      const double *__restrict__ gridnorm,
                                                                      const double *__restrict__ box, double cutoff,
                                                                      int32_t *__restrict__ closest, double *__restrict__ dists,
                                                                      int32_t *__restrict__ numwater) {
    __shared__ float s_t[kIfaceTile * 3];
    const int i = blockIdx.x * (kIfaceThreads / SPLIT) + threadIdx.x / SPLIT;
    const int part = threadIdx.x % SPLIT;
    const bool valid = i < n_row;
    const Box3 b = load_box3(box);
    // a non-periodic axis (negative edge, iBoxL = 0) simply never wraps in the float pass either
    const float Lxf = (float)b.L[0], Lyf = (float)b.L[1], Lzf = (float)b.L[2];
    const float iLxf = (float)b.iL[0], iLyf = (float)b.iL[1], iLzf = (float)b.iL[2];
    double rx = 0, ry = 0, rz = 0;
    if (valid) {
        rx = row[3 * (size_t)i + 0]; ry = row[3 * (size_t)i + 1]; rz = row[3 * (size_t)i + 2];
    }
    const float fx = (float)rx, fy = (float)ry, fz = (float)rz;
    NearestScan S;
    S.reset();
    for (int t0 = 0; t0 < n_col; t0 += kIfaceTile) {
        const int nt = min(kIfaceTile, n_col - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < nt * 3; k += kIfaceThreads) s_t[k] = (float)col[3 * (size_t)t0 + k];
        __syncthreads();
        if (valid) nearest_tile<ROW_IS_WATER>(S, s_t, nt, t0, col, rx, ry, rz, fx, fy, fz, b, Lxf, Lyf, Lzf, iLxf, iLyf, iLzf, part, SPLIT);
    }
    if (SPLIT > 1) {
#pragma unroll
        for (int o = 1; o < SPLIT; o <<= 1) {
            const double ob = __shfl_xor_sync(kFullMask, S.best, o);
            const int oc = __shfl_xor_sync(kFullMask, S.close, o);
            if (oc >= 0 && (ob < S.best || (ob == S.best && (S.close < 0 || oc < S.close)))) {
                S.best = ob;
                S.close = oc;
            }
        }
        if (part != 0) return;  // (SPLIT > 1 is only instantiated for the search that needs no warp vote below)
    }
    if (!ROW_IS_WATER) {
        if (valid) closest[i] = S.close;
        return;
    }
    bool counted = false;
    if (valid) {
        double proj = 0.0;
        if (S.close >= 0) {
            const double *g = col + 3 * (size_t)S.close, *cn = gridnorm + 3 * (size_t)S.close;
            const double dx = min_image_1<double, true>(rx, g[0], b.L[0], b.iL[0]);
            const double dy = min_image_1<double, true>(ry, g[1], b.L[1], b.iL[1]);
            const double dz = min_image_1<double, true>(rz, g[2], b.L[2], b.iL[2]);
            proj = dot3<double>(dx, dy, dz, cn[0], cn[1], cn[2]);  // sum(normvec * closenorm)  (:1464)
            counted = proj <= cutoff;
        }
        closest[i] = S.close;
        dists[i] = proj;
    }
    const unsigned m = __ballot_sync(kFullMask, counted);
    if ((threadIdx.x & 31) == 0 && m != 0u && numwater) atomicAdd(numwater, __popc(m));
}

// ---- depth-binned profile -------------------------------------------------------------------------------

constexpr int kProfileMaxBins = 2048;

__global__ void __launch_bounds__(256) profile_kernel(const double *__restrict__ value, const double *__restrict__ coord, long long n,
                                                      double lo, double width, int nbins, unsigned long long *__restrict__ count,
                                                      double *__restrict__ sum, double *__restrict__ sumsq) {
    extern __shared__ double s_acc[];  // [3][nbins] when it fits
    const bool smem = nbins <= kProfileMaxBins;
    if (smem) {
        for (int i = threadIdx.x; i < 3 * nbins; i += blockDim.x) s_acc[i] = 0.0;
        __syncthreads();
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = coord[i], v = value[i];
        const double t = floor(__ddiv_rn(__dsub_rn(d, lo), width));
        if (!(t >= 0.0) || !(t < (double)nbins)) continue;
        const int bin = (int)t;
        if (smem) {
            atomicAdd(s_acc + bin, 1.0);
            atomicAdd(s_acc + nbins + bin, v);
            atomicAdd(s_acc + 2 * nbins + bin, v * v);
        } else {
            atomicAdd(count + bin, 1ull);
            atomicAdd(sum + bin, v);
            atomicAdd(sumsq + bin, v * v);
        }
    }
    if (smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
            const double c = s_acc[i];
            if (c != 0.0) {
                atomicAdd(count + i, (unsigned long long)c);
                atomicAdd(sum + i, s_acc[nbins + i]);
                atomicAdd(sumsq + i, s_acc[2 * nbins + i]);
            }
        }
    }
}

// ---- iso-surface vertices -------------------------------------------------------------------------------
// Every grid edge node -> next node along x, y, z whose end values straddle the level carries one vertex at the
// linear interpolation -- the vertex set a marching-cubes mesh of the field has (structureLibs/surface_library.py:202).
// Two passes around an exclusive scan so that the output order is deterministic: ascending node index
// ((i * ny + j) * nz + k), then axis.

struct IsoParams {
    const double *dens;  // [nx][ny][nz]
    const double *gx, *gy, *gz;
    int nx, ny, nz;
    long long n_nodes;
    double level;
    uint32_t *offs;      // [n_nodes + 1]
    long long capacity;  // vertices `points` holds
    double *points;      // [capacity][3]
    int32_t *n_total;
};

__device__ __forceinline__ bool iso_above(double v, double level) { return v > level; }

template <bool FILL>
__global__ void __launch_bounds__(256) iso_kernel(const IsoParams P) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > P.n_nodes) return;
    if (g == P.n_nodes) {  // sentinel: the scan turns it into the total
        if (!FILL) P.offs[g] = 0u;
        else *P.n_total = (int32_t)P.offs[g];
        return;
    }
    const int k = (int)(g % P.nz), j = (int)((g / P.nz) % P.ny), i = (int)(g / ((long long)P.nz * P.ny));
    const double va = P.dens[g];
    const bool a = iso_above(va, P.level);
    const long long step[3] = {(long long)P.ny * P.nz, (long long)P.nz, 1ll};
    const bool has[3] = {i + 1 < P.nx, j + 1 < P.ny, k + 1 < P.nz};
    uint32_t n = 0;
    long long o = FILL ? (long long)P.offs[g] : 0ll;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if (!has[ax]) continue;
        const double vb = P.dens[g + step[ax]];
        if (iso_above(vb, P.level) == a) continue;
        ++n;
        if (FILL) {
            if (o < P.capacity) {
                const double t = __ddiv_rn(__dsub_rn(P.level, va), __dsub_rn(vb, va));
                double p[3] = {P.gx[i], P.gy[j], P.gz[k]};
                const double nxt = ax == 0 ? P.gx[i + 1] : (ax == 1 ? P.gy[j + 1] : P.gz[k + 1]);
                p[ax] = __dadd_rn(p[ax], __dmul_rn(t, __dsub_rn(nxt, p[ax])));
                P.points[3 * o + 0] = p[0];
                P.points[3 * o + 1] = p[1];
                P.points[3 * o + 2] = p[2];
            }
            ++o;
        }
    }
    if (!FILL) P.offs[g] = n;
}

// ---- iso-surface faces ------------------------------------------------------------------------------------------
// Triangles over the vertices above: one thread per grid cube, corner configuration -> kMcTri (wol_mc_table.h, generated
// from first principles by scripts/make_mc_table.py: faces joined so that inside regions never connect across a face
// diagonal, which makes the surface watertight; normals point towards lower values).  A triangle corner is the vertex of a
// cube edge = (lower node of the edge, axis) = voffs[node] + the number of crossed edges of that node along lower axes.
// Same count / scan / fill scheme, so faces come out ordered by cube index.  Optional per-face area with the reference's
// rule (fortran/imagelib.f90:254-267: (v1.v1 v2.v2)^0.5 (1 - cos^2)^0.5, i.e. |v1 x v2| -- twice the geometric area).

struct IsoFaceParams {
    const double *dens;
    int nx, ny, nz;
    long long n_cubes;
    double level;
    const uint32_t *voffs;  // [n_nodes + 1] vertex offsets (the scratch wol_iso_points filled)
    uint32_t *foffs;        // [n_cubes + 1]
    long long capacity;
    int32_t *faces;         // [capacity][3]
    const double *points;   // vertices, for the areas (may be null)
    double *areas;          // [capacity] (may be null)
    int32_t *n_total;
};

__device__ __forceinline__ int iso_vertex_id(const IsoFaceParams &P, int i, int j, int k, int axis) {
    const long long g = ((long long)i * P.ny + j) * P.nz + k;
    const bool a = iso_above(P.dens[g], P.level);
    int id = (int)P.voffs[g];
    if (axis > 0 && i + 1 < P.nx && iso_above(P.dens[g + (long long)P.ny * P.nz], P.level) != a) ++id;
    if (axis > 1 && j + 1 < P.ny && iso_above(P.dens[g + P.nz], P.level) != a) ++id;
    return id;
}

template <bool FILL>
__global__ void __launch_bounds__(128) iso_face_kernel(const IsoFaceParams P) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > P.n_cubes) return;
    if (c == P.n_cubes) {
        if (!FILL) P.foffs[c] = 0u;
        else *P.n_total = (int32_t)P.foffs[c];
        return;
    }
    const int cz = P.nz - 1, cy = P.ny - 1;
    const int k = (int)(c % cz), j = (int)((c / cz) % cy), i = (int)(c / ((long long)cz * cy));
    int cfg = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const long long g = ((long long)(i + (q & 1)) * P.ny + (j + ((q >> 1) & 1))) * P.nz + (k + (q >> 2));
        if (iso_above(P.dens[g], P.level)) cfg |= 1 << q;
    }
    int n = 0;
    while (n < 5 && kMcTri[cfg][3 * n] >= 0) ++n;
    if (!FILL) {
        P.foffs[c] = (uint32_t)n;
        return;
    }
    long long o = (long long)P.foffs[c];
    for (int t = 0; t < n; ++t, ++o) {
        if (o >= P.capacity) break;
        int v[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const int e = kMcTri[cfg][3 * t + m];
            const int a = kMcEdge[e][0], ax = kMcEdge[e][2];
            v[m] = iso_vertex_id(P, i + (a & 1), j + ((a >> 1) & 1), k + (a >> 2), ax);
        }
        P.faces[3 * o + 0] = v[0];
        P.faces[3 * o + 1] = v[1];
        P.faces[3 * o + 2] = v[2];
        if (P.areas && P.points) {
            double v1[3], v2[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                v1[d] = __dsub_rn(P.points[3 * (size_t)v[1] + d], P.points[3 * (size_t)v[0] + d]);
                v2[d] = __dsub_rn(P.points[3 * (size_t)v[2] + d], P.points[3 * (size_t)v[0] + d]);
            }
            const double s1 = sumsq3<double>(v1[0], v1[1], v1[2]), s2 = sumsq3<double>(v2[0], v2[1], v2[2]);
            const double root = __dsqrt_rn(__dmul_rn(s1, s2));
            // (a degenerate face -- two vertices on the same grid node -- is 0 / 0 = NaN in the Fortran rule: 0 here; rounding
            // can push 1 - cos^2 of a sliver below zero: clamped)
            const double cth = root > 0.0 ? __ddiv_rn(dot3<double>(v1[0], v1[1], v1[2], v2[0], v2[1], v2[2]), root) : 1.0;
            P.areas[o] = __dmul_rn(root, __dsqrt_rn(fmax(__dsub_rn(1.0, __dmul_rn(cth, cth)), 0.0)));
        }
    }
}

// ---- histrr3b ---------------------------------------------------------------------------------------------
// One warp per centre.  The lanes walk the candidates of the 27-cell stencil (cell edge >= dNum * distWidth),
// keep those whose distance bin is inside the histogram in a shared list, then spread the list's pairs over
// the lanes: the pair's lower ATOM INDEX supplies the first distance bin (the reference's j < k loops), the
// angle bin comes from the clamped cosine through the ceiling-rule threshold table.

constexpr int kRR3Threads = 128;
constexpr int kRR3Cap = 192;  // neighbours per centre inside the histogram's distance range

struct RR3Params {
    CellGridS grid;
    const double *box;
    int n_pos;
    double dwidth, awidth;
    int dnum, anum;
    const double *table;  // ceiling-rule angle table (wol_angle_table_ceil)
    unsigned long long *hist;  // [dnum][dnum][anum]
    uint32_t *counters;
};

struct RR3Smem {
    Vec4<double> v[kRR3Cap];
    int bin[kRR3Cap];
    int idx[kRR3Cap];
    int n;
};

__global__ void __launch_bounds__(kRR3Threads) histrr3b_kernel(const RR3Params P) {
    __shared__ RR3Smem S[kRR3Threads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RR3Smem &W = S[warp];
    const Box3 b = load_box3(P.box);
    const int nc0 = P.grid.nc0, nc1 = P.grid.nc1, nc2 = P.grid.nc2;
    const int cntx = min(3, nc0), cnty = min(3, nc1), cntz = min(3, nc2);
    const int ncell27 = cntx * cnty * cntz;
    const double inv_aw = 1.0 / P.awidth;
    for (int i = blockIdx.x * (kRR3Threads / 32) + warp; i < P.n_pos; i += gridDim.x * (kRR3Threads / 32)) {
        // centre i = record i (cell order); the histogram does not depend on the order of the centres
        double rx, ry, rz;
        int self_idx;
        RecTraits<double>::load(P.grid.recs, (size_t)i, rx, ry, rz, self_idx);
        const int cx = cell_coord(rx, b.iL[0], nc0), cy = cell_coord(ry, b.iL[1], nc1), cz = cell_coord(rz, b.iL[2], nc2);
        const int xs = (nc0 <= 3) ? 0 : (cx - 1 + nc0) % nc0, ys = (nc1 <= 3) ? 0 : (cy - 1 + nc1) % nc1,
                  zs = (nc2 <= 3) ? 0 : (cz - 1 + nc2) % nc2;
        __syncwarp();
        if (lane == 0) W.n = 0;
        __syncwarp();
        for (int c27 = lane; c27 < ncell27; c27 += 32) {
            const int ix = c27 % cntx, iy = (c27 / cntx) % cnty, iz = c27 / (cntx * cnty);
            const size_t c = ((size_t)((zs + iz) % nc2) * nc1 + (ys + iy) % nc1) * nc0 + (xs + ix) % nc0;
            const int j1 = (int)__ldg(P.grid.cell_start + c + 1);
            for (int j = (int)__ldg(P.grid.cell_start + c); j < j1; ++j) {
                if (j == i) continue;
                double px, py, pz;
                int id;
                RecTraits<double>::load(P.grid.recs, (size_t)j, px, py, pz, id);
                Vec4<double> v;
                v.x = min_image_1<double, true>(px, rx, b.L[0], b.iL[0]);
                v.y = min_image_1<double, true>(py, ry, b.L[1], b.iL[1]);
                v.z = min_image_1<double, true>(pz, rz, b.L[2], b.iL[2]);
                v.w = sumsq3<double>(v.x, v.y, v.z);
                const double t = ceil(__ddiv_rn(__dsqrt_rn(v.w), P.dwidth));  // dbin = ceiling(dist / distWidth)
                if (!(t <= (double)P.dnum) || t < 1.0) continue;                // bin 0 (coincident atoms) is out of bounds in the Fortran
                const int at = atomicAdd(&W.n, 1);
                if (at < kRR3Cap) {
                    W.v[at] = v;
                    W.bin[at] = (int)t - 1;
                    W.idx[at] = id;
                }
            }
        }
        __syncwarp();
        const int n = W.n;
        if (n > kRR3Cap) {
            if (lane == 0) atomicAdd(P.counters + kCntFatal, 1u);
            continue;
        }
        const int npairs = n * (n - 1) / 2;
        for (int p = lane; p < npairs; p += 32) {
            int hi = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
            while (hi * (hi - 1) / 2 > p) --hi;
            while ((hi + 1) * hi / 2 <= p) ++hi;
            const int lo = p - hi * (hi - 1) / 2;
            const Vec4<double> va = W.v[lo], vb = W.v[hi];
            const bool lo_first = W.idx[lo] < W.idx[hi];
            const int d1 = lo_first ? W.bin[lo] : W.bin[hi], d2 = lo_first ? W.bin[hi] : W.bin[lo];
            // CosAngle3(distvec1, 0, distvec2) (:1583): Vec21 = distvec1 - 0, Vec23 = distvec2 - 0
            const double c = clamped_cos<double>(dot3<double>(va.x, va.y, va.z, vb.x, vb.y, vb.z), va.w, vb.w);
            const int pos = angle_position(c, P.table, P.anum, 0.0, inv_aw);
            if (pos >= 0 && pos < P.anum) atomicAdd(P.hist + ((size_t)d1 * P.dnum + d2) * P.anum + pos, 1ull);
        }
    }
}

static CellGridS make_grid_s(void *workspace, const WorkspaceLayout &lay, const int32_t nc[3]) {
    char *ws = reinterpret_cast<char *>(workspace);
    CellGridS g;
    g.cell_start = reinterpret_cast<const uint32_t *>(ws + lay.off_cell_start);
    g.recs = ws + lay.off_recs;
    g.nc0 = nc[0];
    g.nc1 = nc[1];
    g.nc2 = nc[2];
    return g;
}

}  // namespace wol

using namespace wol;

extern "C" {

int wol_willard_density(const double *points, int64_t n_points, const double *gridx, const double *gridy, const double *gridz,
                        int32_t nx, int32_t ny, int32_t nz, const double *box, int32_t n_pos, const int32_t nc[3], double edge_min,
                        double smoothlen, void *workspace, size_t workspace_bytes, double *densvals, double *densnorms,
                        void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!box || !nc || !workspace || !densvals) return set_error(WOL_ERR_INVALID, "wol_willard_density: null argument");
    if (!(smoothlen > 0.0)) return set_error(WOL_ERR_INVALID, "wol_willard_density: smoothlen must be positive");
    long long n = n_points;
    if (!points) {
        if (!gridx || !gridy || !gridz || nx < 0 || ny < 0 || nz < 0) return set_error(WOL_ERR_INVALID, "wol_willard_density: bad grid");
        n = (long long)nx * ny * nz;
    }
    if (n < 0) return set_error(WOL_ERR_INVALID, "wol_willard_density: negative size");
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && 3.0 * smoothlen * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "Gaussian cut-off %.6g exceeds the planned cell edge %.6g", 3.0 * smoothlen, edge_min);
    const WorkspaceLayout lay = workspace_layout(1, n_pos, n_pos, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    WillardParams P;
    P.grid = make_grid_s(workspace, lay, nc);
    P.box = box;
    P.pts = points;
    P.gx = gridx; P.gy = gridy; P.gz = gridz;
    P.nx = nx; P.ny = ny; P.nz = nz;
    P.n_points = n;
    const double pi = 3.1415926535897931;
    P.s2 = smoothlen * smoothlen;
    P.pref = pow(2.0 * pi * (smoothlen * smoothlen), 1.5);   // host libm, like the reference's Fortran runtime
    P.shiftterm = exp(-9.0 / 2.0) / P.pref;
    P.cut = 9.0 * (smoothlen * smoothlen);
    P.dens = densvals;
    P.norms = densnorms;
    if (n > 0) {
        willard_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_willard_density", e);
    return WOL_OK;
}

int wol_density_field(const double *gridx, const double *gridy, const double *gridz, int32_t nx, int32_t ny, int32_t nz, double binwidth,
                      const double *box, int32_t n_pos, const int32_t nc[3], double edge_min, void *workspace, size_t workspace_bytes,
                      double *densvals, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!gridx || !gridy || !gridz || !box || !nc || !workspace || !densvals || nx < 0 || ny < 0 || nz < 0)
        return set_error(WOL_ERR_INVALID, "wol_density_field: bad argument");
    if (!(binwidth > 0.0)) return set_error(WOL_ERR_INVALID, "wol_density_field: binwidth (gridx[1] - gridx[0]) must be positive");
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && 0.5 * binwidth * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "half bin width %.6g exceeds the planned cell edge %.6g", 0.5 * binwidth, edge_min);
    const WorkspaceLayout lay = workspace_layout(1, n_pos, n_pos, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    VoxelParams P;
    P.grid = make_grid_s(workspace, lay, nc);
    P.box = box;
    P.gx = gridx; P.gy = gridy; P.gz = gridz;
    P.nx = nx; P.ny = ny; P.nz = nz;
    P.half = binwidth / 2.0;
    P.vol = pow(binwidth, 3.0);  // binwidth**3.0 through the host libm, like the Fortran runtime
    P.dens = densvals;
    const long long n = (long long)nx * ny * nz;
    if (n > 0) {
        voxel_density_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_density_field", e);
    return WOL_OK;
}

int wol_interface_water(const double *pos, int32_t n_pos, const double *gridpos, const double *gridnorm, int32_t n_grid,
                        double cutoff, const double *box, int32_t *watclose, int32_t *surfclose, int32_t *numwater,
                        double *allwatdists, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pos < 0 || n_grid < 0 || !box || (!pos && n_pos > 0) || ((!gridpos || !gridnorm) && n_grid > 0) || !watclose || !allwatdists)
        return set_error(WOL_ERR_INVALID, "wol_interface_water: bad argument");
    if (n_pos > 0) {
        iface_nearest_kernel<true, 1><<<(n_pos + kIfaceThreads - 1) / kIfaceThreads, kIfaceThreads, 0, stream>>>(
            pos, n_pos, gridpos, n_grid, gridnorm, box, cutoff, watclose, allwatdists, numwater);
        add_launches(1);
    }
    if (surfclose && n_grid > 0) {
        // few surface points: eight threads per point, so that the grid fills the SMs
        constexpr int kSplit = 8;
        if ((n_grid + kIfaceThreads - 1) / kIfaceThreads < 8 * sm_count())
            iface_nearest_kernel<false, kSplit><<<(n_grid + kIfaceThreads / kSplit - 1) / (kIfaceThreads / kSplit), kIfaceThreads, 0, stream>>>(
                gridpos, n_grid, pos, n_pos, nullptr, box, cutoff, surfclose, nullptr, nullptr);
        else
            iface_nearest_kernel<false, 1><<<(n_grid + kIfaceThreads - 1) / kIfaceThreads, kIfaceThreads, 0, stream>>>(
                gridpos, n_grid, pos, n_pos, nullptr, box, cutoff, surfclose, nullptr, nullptr);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_interface_water", e);
    return WOL_OK;
}

int wol_iso_points(const double *densvals, const double *gridx, const double *gridy, const double *gridz, int32_t nx, int32_t ny,
                   int32_t nz, double level, uint32_t *scratch, size_t scratch_bytes, double *points, int64_t capacity,
                   int32_t *n_total, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!densvals || !gridx || !gridy || !gridz || !scratch || !n_total) return set_error(WOL_ERR_INVALID, "wol_iso_points: null argument");
    if (nx < 1 || ny < 1 || nz < 1 || capacity < 0 || (capacity > 0 && !points)) return set_error(WOL_ERR_INVALID, "wol_iso_points: bad size");
    const long long n = (long long)nx * ny * nz;
    if (3 * n > 0x7fffffffll) return set_error(WOL_ERR_INVALID, "wol_iso_points: grid of %lld nodes is too large", n);
    const size_t need = wol_iso_scratch_bytes(nx, ny, nz);
    if (scratch_bytes < need) return set_error(WOL_ERR_WORKSPACE, "scratch holds %zu bytes, %zu needed", scratch_bytes, need);
    IsoParams P;
    P.dens = densvals;
    P.gx = gridx; P.gy = gridy; P.gz = gridz;
    P.nx = nx; P.ny = ny; P.nz = nz;
    P.n_nodes = n;
    P.level = level;
    P.offs = scratch;
    P.capacity = capacity;
    P.points = points;
    P.n_total = n_total;
    const unsigned blocks = (unsigned)((n + 1 + 255) / 256);
    iso_kernel<false><<<blocks, 256, 0, stream>>>(P);
    exclusive_scan_u32(scratch, (size_t)n + 1, scratch + (((size_t)n + 1 + 3) & ~(size_t)3), stream);
    iso_kernel<true><<<blocks, 256, 0, stream>>>(P);
    add_launches(2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_iso_points", e);
    return WOL_OK;
}

size_t wol_iso_face_scratch_bytes(int32_t nx, int32_t ny, int32_t nz) {
    if (nx < 2 || ny < 2 || nz < 2) return 16;
    const size_t n = (size_t)(nx - 1) * (ny - 1) * (nz - 1) + 1;
    return sizeof(uint32_t) * (((n + 3) & ~(size_t)3) + n / kScanTile + 2);
}

int wol_iso_faces(const double *densvals, int32_t nx, int32_t ny, int32_t nz, double level, const uint32_t *vertex_scratch,
                  uint32_t *face_scratch, size_t face_scratch_bytes, int32_t *faces, int64_t capacity, const double *points,
                  double *areas, int32_t *n_total, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!densvals || !vertex_scratch || !face_scratch || !n_total) return set_error(WOL_ERR_INVALID, "wol_iso_faces: null argument");
    if (nx < 1 || ny < 1 || nz < 1 || capacity < 0 || (capacity > 0 && !faces)) return set_error(WOL_ERR_INVALID, "wol_iso_faces: bad size");
    if (face_scratch_bytes < wol_iso_face_scratch_bytes(nx, ny, nz))
        return set_error(WOL_ERR_WORKSPACE, "face scratch holds %zu bytes, %zu needed", face_scratch_bytes, wol_iso_face_scratch_bytes(nx, ny, nz));
    IsoFaceParams P;
    P.dens = densvals;
    P.nx = nx; P.ny = ny; P.nz = nz;
    P.n_cubes = (nx < 2 || ny < 2 || nz < 2) ? 0 : (long long)(nx - 1) * (ny - 1) * (nz - 1);
    P.level = level;
    P.voffs = vertex_scratch;
    P.foffs = face_scratch;
    P.capacity = capacity;
    P.faces = faces;
    P.points = points;
    P.areas = areas;
    P.n_total = n_total;
    const unsigned blocks = (unsigned)((P.n_cubes + 1 + 127) / 128);
    iso_face_kernel<false><<<blocks, 128, 0, stream>>>(P);
    exclusive_scan_u32(face_scratch, (size_t)P.n_cubes + 1, face_scratch + (((size_t)P.n_cubes + 1 + 3) & ~(size_t)3), stream);
    iso_face_kernel<true><<<blocks, 128, 0, stream>>>(P);
    add_launches(2);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_iso_faces", e);
    return WOL_OK;
}

size_t wol_iso_scratch_bytes(int32_t nx, int32_t ny, int32_t nz) {
    if (nx < 1 || ny < 1 || nz < 1) return 0;
    const size_t n = (size_t)nx * ny * nz + 1;
    return sizeof(uint32_t) * (((n + 3) & ~(size_t)3) + n / kScanTile + 2);
}

int wol_profile_bins(const double *value, const double *coord, int64_t n, double lo, double width, int32_t nbins, int64_t *count,
                     double *sum, double *sumsq, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || nbins < 1 || !(width > 0.0) || !count || !sum || !sumsq || ((!value || !coord) && n > 0))
        return set_error(WOL_ERR_INVALID, "wol_profile_bins: bad argument");
    if (n == 0) return WOL_OK;
    const size_t smem = nbins <= kProfileMaxBins ? sizeof(double) * 3 * nbins : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(profile)", e);
    }
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    profile_kernel<<<(unsigned)blocks, 256, smem, stream>>>(value, coord, n, lo, width, nbins, reinterpret_cast<unsigned long long *>(count),
                                                            sum, sumsq);
    add_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_profile_bins", e);
    return WOL_OK;
}

int wol_histrr3b(const double *box, int32_t n_pos, const int32_t nc[3], double edge_min, double dist_width, int32_t d_num,
                 double ang_width, int32_t a_num, const double *angle_table, void *workspace, size_t workspace_bytes, int64_t *hist,
                 void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!box || !nc || !workspace || !hist || !angle_table) return set_error(WOL_ERR_INVALID, "wol_histrr3b: null argument");
    if (d_num < 1 || a_num < 1 || !(dist_width > 0.0) || !(ang_width > 0.0)) return set_error(WOL_ERR_INVALID, "wol_histrr3b: bad bin spec");
    const double reach = dist_width * d_num;
    for (int k = 0; k < 3; ++k)
        if (nc[k] > 3 && reach * (1.0 + 1e-9) > edge_min)
            return set_error(WOL_ERR_INVALID, "histogram range %.6g exceeds the planned cell edge %.6g", reach, edge_min);
    const WorkspaceLayout lay = workspace_layout(1, n_pos, n_pos, nc);
    if (workspace_bytes < lay.total) return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay.total);
    RR3Params P;
    P.grid = make_grid_s(workspace, lay, nc);
    P.box = box;
    P.n_pos = n_pos;
    P.dwidth = dist_width;
    P.awidth = ang_width;
    P.dnum = d_num;
    P.anum = a_num;
    P.table = angle_table;
    P.hist = reinterpret_cast<unsigned long long *>(hist);
    P.counters = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(workspace) + lay.off_counters);
    if (n_pos > 0) {
        long long blocks = ((long long)n_pos + 3) / 4;
        const long long cap = (long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        histrr3b_kernel<<<(unsigned)blocks, kRR3Threads, 0, stream>>>(P);
        add_launches(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_histrr3b", e);
    return WOL_OK;
}

}  // extern "C"
