/*
 * wol_capi.h -- C ABI of the B200-native water-structure backend (libwol.so).
 *
 * This is the drop-in boundary for the per-frame hot path of WaterOrderLib.  In the reference the
 * seam is the f2py extension module `waterlib` (built by fortran/buildWrappers.sh:3-5, imported at
 * structureLibs/water_properties.py:42-43 and structureLibs/orderParam_lib.py:28-30); each entry
 * point below names the f2py routine(s) and Python loop it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, enums; no torch / CUDA types in any signature (`stream` is a
 *     cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - every array pointer is a DEVICE pointer unless the name ends in `_host`.
 *   - arrays are dense row-major: positions [n_frames][n][3], boxes [n_frames][3] (orthorhombic edge
 *     lengths, what the reference takes from frame.box.values[:3], orderParam_lib.py:1315).
 *   - the caller owns every buffer including scratch (`wol_workspace_bytes`); the library never
 *     allocates device memory, never synchronises the host (except wol_status, which exists to do so)
 *     and keeps no state besides a thread-local error string.  All work is enqueued on `stream`.
 *   - return value: WOL_OK or a negative WOL_ERR_* code; wol_last_error() describes the failure.
 *     Nothing aborts the process (the reference's Fortran `stop`s do, waterlib.f90:1171-1174).
 *   - indices are 0-based int32; "no neighbour" is -1.
 *   - the library never synchronises the host except in wol_status and wol_effective_box (open axes only).
 *
 * Limits (reported, never silent)
 *   - n_frames * max(n_pos, n_centres) < 2^30 per call (WOL_ERR_RANGE): batch longer trajectories.
 *   - at most 1024 cells per axis (wol_plan_grid coarsens the grid beyond that).
 *   - more than 1024 neighbours of one centre inside a cutoff: WOL_ERR_CAPACITY from wol_status.
 *   - wol_angles_fill materialises at most 64 neighbours per centre (the histogram path has no such limit).
 *   - materialising calls address at most 2^32 - 2 angles / pairs (WOL_ERR_RANGE; offsets[last] = 0xFFFFFFFF).
 *   - box edges must be finite and non-zero; negative (= non-periodic) edges go through wol_effective_box first.
 */
#ifndef WOL_CAPI_H
#define WOL_CAPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bumped when a struct layout or an existing signature changes (2: wol_q3b_args.n_valid); added entry points do not bump it */
#define WOL_ABI_VERSION 2

enum {
    WOL_OK = 0,
    WOL_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unknown enum)          */
    WOL_ERR_UNSUPPORTED = -2, /* valid in the reference but not implemented by this entry point     */
    WOL_ERR_WORKSPACE = -3,   /* workspace too small                                                */
    WOL_ERR_CUDA = -4,        /* a CUDA runtime call failed; message holds cudaGetErrorString       */
    WOL_ERR_RANGE = -5,       /* sizes overflow the 32-bit indexing of the kernels                  */
    WOL_ERR_CAPACITY = -6     /* a device-side list overflowed (reported by wol_status)             */
};

enum { WOL_F64 = 0, WOL_F32 = 1 };            /* storage dtype of a position array               */
enum { WOL_SUM_I64 = 0, WOL_SUM_F64 = 1 };    /* element type of wol_hist_allreduce              */
enum { WOL_PREC_FP64 = 0, WOL_PREC_FP32 = 1 }; /* arithmetic mode of the evaluation kernels       */

/* Number of doubles per frame in `frame_stats` (see wol_q3b_args). */
#define WOL_NSTATS 8
enum {
    WOL_STAT_Q_SUM = 0,     /* sum of q over centres                                   */
    WOL_STAT_Q_SUMSQ = 1,   /* sum of q^2                                              */
    WOL_STAT_N_CENTRES = 2, /* number of centres accumulated                           */
    WOL_STAT_TET_COUNT = 3, /* angles with 100 <= theta <= 120 (water_properties.py:330) */
    WOL_STAT_TET_COS = 4,   /* sum of cos(theta) over that window                      */
    WOL_STAT_TET_COSSQ = 5, /* sum of cos^2                                            */
    WOL_STAT_N_ANGLES = 6,  /* all angles produced (incl. the -180 deg ones)           */
    WOL_STAT_N_NEIGH = 7    /* sum of 3-body neighbour counts                          */
};

/* Extra doubles that follow the nbins + 1 thresholds of an angle-bin table (wol_angle_table). */
#define WOL_TABLE_EXTRA 8

const char *wol_version(void);
const char *wol_last_error(void);
int wol_abi_version(void);

/*
 * Cell-grid plan for a batch of frames (host-only, cheap).  Picks nc[3] (cells per axis, identical
 * for all frames of the batch) such that every frame's cell edge L/nc is >= r_cell (1 + 1e-9), and
 * reports the smallest cell edge of the batch, which the evaluation kernels use as the radius inside
 * which a 27-cell sweep is complete, and the largest box edge.  Replaces nothing in the reference (its search is the O(N^2)
 * double loop of waterlib.f90:846-861).
 */
int wol_plan_grid(const double *box_host, int32_t n_frames, double r_cell, int32_t nc_out[3],
                  double *edge_min_out, double *box_max_out);

/*
 * Non-periodic axes.  The reference marks an axis as not periodic with a NEGATIVE box edge (fortran/waterlib.f90:41,
 * :840: iBoxL = merge(1/BoxL, 0, BoxL >= 0), so the minimum-image step leaves that component untouched).  The cell-list
 * entry points need a positive period on every axis; this call replaces every negative edge by an equivalent period
 *     L' = 2 (extent + reach) + 1,   extent = max - min of that coordinate over the frame's atoms (and centres),
 * for which the kernels' arithmetic is bit-identical to "no wrap" (every difference d has |d| < L'/2, so
 * d - L' anint(d / L') = d) and no periodic image comes within `reach` of any atom.  Pass the result to wol_plan_grid,
 * wol_cell_build and the evaluation calls in place of the original box.
 *   box_host      [n_frames][3] as the reference takes it (negative = open axis)
 *   reach         the largest cutoff / search radius any later call on these frames uses
 *   scratch_dev   n_frames * 48 bytes of device memory (only touched when some edge is negative; may be NULL otherwise)
 *   box_out_host  [n_frames][3], all positive
 * Positive edges are copied unchanged and nothing is launched when no edge is negative; otherwise one small reduction
 * kernel per array runs on `stream` and the call synchronises it.  Zero, NaN or infinite edges: WOL_ERR_INVALID.
 */
int wol_effective_box(const void *pos, int32_t pos_dtype, int32_t n_frames, int32_t n_pos, const void *centres,
                      int32_t centre_dtype, int32_t n_centres, const double *box_host, double reach, void *scratch_dev,
                      double *box_out_host, void *stream);

/* Bytes of scratch needed by wol_cell_build + the evaluation kernels for this batch shape.
 * n_centres_max = the largest number of centres per frame any later call on this workspace passes. */
size_t wol_workspace_bytes(int32_t n_frames, int32_t n_pos, int32_t n_centres_max, const int32_t nc[3]);

/*
 * K1: cell-list build (counting sort of atoms by cell).
 *   pos        [n_frames][n_pos][3] of pos_dtype
 *   box        [n_frames][3] double, all > 0
 *   precision  WOL_PREC_FP64: 32-byte records (x, y, z double + index + cell)
 *              WOL_PREC_FP32: 16-byte records (x, y, z float + index)
 * Afterwards `workspace` holds the cell list consumed by the entry points below.
 */
int wol_cell_build(const void *pos, int32_t pos_dtype, const double *box, int32_t n_frames, int32_t n_pos,
                   const int32_t nc[3], int32_t precision, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Angle-bin table (host-only).  The reference bins the angle acos(c) in degrees with np.histogram
 * (water_properties.py:328, CosAngle3 waterlib.f90:696-702).  Because that chain is monotone in the
 * clamped cosine c, bin membership is decided on the device by comparing c with nbins + 1 thresholds
 * that this routine finds by bisection over the doubles with the host libm -- the same acos the
 * reference's Fortran calls -- so no device transcendental sits between c and its bin.
 *   table_host[k], k = 0..nbins : largest c whose angle lies at or beyond bin k (k = nbins: beyond hi)
 *   table_host[nbins+1]         : bin position of the angle the reference returns for c == -1 (-180)
 *   table_host[nbins+2]         : bin position of a 0-degree angle (coincident-position rule, :690)
 *   table_host[nbins+3..+4]     : largest c with angle >= tet_lo, smallest c with angle <= tet_hi
 *   table_host[nbins+5]         : 1.0 if the table passed its monotonicity self-check
 * The caller uploads the nbins + 1 + WOL_TABLE_EXTRA doubles to the device and passes them as
 * wol_q3b_args.angle_table.
 */
int wol_angle_table(double hist_lo, double hist_hi, int32_t nbins, double tet_lo, double tet_hi,
                    double *table_host);

/*
 * K2: fused neighbour sweep -> 4 nearest neighbours -> tetrahedral q, plus all three-body angles among
 * the neighbours inside (low3, high3] -> histogram and tetrahedral-window sums.
 *
 * Replaces, per frame: wl.allnearneighbors / wl.nearneighbors (fortran/waterlib.f90:830-862, :710-743),
 * the per-water Python loops of getOrderParamq and getCosAngs with their wl.reimage / np.argsort /
 * wl.tetracosang calls (structureLibs/water_properties.py:369-388, :241-248; waterlib.f90:32-47,
 * :867-895, :683-703) and the np.histogram of tetrahedralMetrics (water_properties.py:328).
 *
 * Centres: `centres` == NULL means "every atom of pos" (the subPos == Pos branch,
 * water_properties.py:363-364); otherwise [n_frames][n_centres][3] positions of centre_dtype (the
 * sub-population branch, :366) -- they need not be members of pos.
 *
 * Outputs (any may be NULL):
 *   q           [n_frames][n_centres] double (FP64 mode) or float (FP32 mode)
 *   nn_idx      [n_frames][n_centres][4] int32, indices into pos of the selected neighbours in
 *               selection order (distance, then index), -1 padded
 *   n3          [n_frames][n_centres] int32 neighbour count inside (low3, high3]  (== numAngs)
 *   ang_hist    [n_frames or 1][nbins] int64, ACCUMULATED (+=): np.histogram(angles, nbins, [hist_lo, hist_hi])
 *   q_hist      [n_frames or 1][q_nbins] int64, ACCUMULATED: np.histogram(q, q_nbins, [0, 1])
 *   frame_stats [n_frames][WOL_NSTATS] double, ACCUMULATED
 */
typedef struct wol_q3b_args {
    uint32_t struct_size; /* sizeof(wol_q3b_args), for forward compatibility */
    int32_t precision;    /* WOL_PREC_FP64 | WOL_PREC_FP32; must match wol_cell_build */
    int32_t n_frames;
    int32_t n_pos;
    int32_t n_centres; /* ignored when centres == NULL (then n_centres = n_pos) */
    int32_t centre_dtype;
    const void *centres;
    const double *box;
    void *workspace; /* as filled by wol_cell_build for the same pos/box/nc */
    size_t workspace_bytes;
    int32_t nc[3];
    int32_t hist_per_frame; /* 1: one histogram row per frame, 0: a single shared row */
    double edge_min;        /* from wol_plan_grid */
    double box_max;         /* from wol_plan_grid: largest box edge of the batch (bounds float rounding) */
    double low3, high3;     /* getCosAngs lowCut/highCut (default 0, 3.413) */
    double lowq, highq;     /* getOrderParamq lowCut/highCut (default 0, 10) */
    int32_t do_q;           /* evaluate the q branch */
    int32_t do_3body;       /* evaluate the three-body branch */
    int32_t nbins;          /* angle histogram bins (default 500) */
    int32_t q_nbins;        /* q histogram bins (default 500) */
    double hist_lo, hist_hi; /* angle histogram range (default 0, 180); must match angle_table */
    const double *angle_table; /* device copy of wol_angle_table output; required when do_3body */
    void *q;
    int32_t *nn_idx;
    int32_t *n3;
    int64_t *ang_hist;
    int64_t *q_hist;
    double *frame_stats;
    /* Optional cudaEvent_t pair (as void*) recorded on `stream` immediately before and after the launch
     * of the dominant evaluation kernel, so a caller can time that kernel alone (bench roofline). */
    void *timing_event_begin;
    void *timing_event_end;
    /* Optional ragged centres: device array [n_frames]; frame f evaluates only its first n_valid[f] centres (the
     * drivers' sub-populations change size from frame to frame, structureLibs/orderParam_lib.py:1343-1346); outputs of
     * the padded slots are left untouched.  NULL: all n_centres of every frame. */
    const int32_t *n_valid;
} wol_q3b_args;

int wol_q3b_frames(const wol_q3b_args *args, void *stream);

/*
 * Device-side status of the last evaluation on this workspace (same shape arguments as
 * wol_workspace_bytes).  Synchronises `stream`.
 *   status_host[0]  number of centres that took the widened-search path (q with < 4 neighbours in 27 cells)
 *   status_host[1]  number of centres whose neighbour list overflowed the fast path
 *   status_host[2]  non-zero: a neighbour list overflowed even the large-capacity path -> results invalid;
 *                   this one is sticky across evaluations on the workspace and cleared by this call
 *   status_host[3]  reserved
 * Returns WOL_ERR_CAPACITY when status_host[2] != 0.  The first 256 bytes of a fresh workspace must be zero (the rest
 * may hold anything).  The counters live in those first 256 bytes of the workspace whatever the batch shape, so one workspace can serve batches
 * of different sizes (a shorter last batch) without the flag moving.
 */
int wol_status(const void *workspace, int32_t n_frames, int32_t n_pos, int32_t n_centres_max, const int32_t nc[3],
               void *stream, int32_t status_host[4]);

/* ------------------------------------------------------------------------------------------------
 * Value-returning routines of the same path: what the reference's Python API hands back as arrays.
 * All fp64 in the reference's operation order.  They share the cell list of wol_cell_build.
 * ------------------------------------------------------------------------------------------------ */

/*
 * Offsets of each centre's block inside the flat angle array getCosAngs returns
 * (structureLibs/water_properties.py:247: np.hstack of the triu entries, centre by centre):
 *   offsets[i] = sum_{k<i} n3[k] (n3[k] - 1) / 2,   i = 0..n   (offsets[n] = number of angles)
 * n3 as written by wol_q3b_frames.  scratch: at least (n / 2048 + 8) uint32, 8-byte aligned.  The offsets are 32-bit: when the
 * total does not fit (> 2^32 - 2 angles) offsets[n] is set to 0xFFFFFFFF -- split the batch.
 */
int wol_angle_offsets(const int32_t *n3, int64_t n, uint32_t *offsets, uint32_t *scratch, void *stream);

/*
 * The angle VALUES (degrees) of getCosAngs in the reference's order: centres ascending, neighbours in
 * ascending atom index, pairs in np.triu_indices(k=1) order (water_properties.py:241-247; CosAngle3
 * fortran/waterlib.f90:683-703 incl. the -180 it returns for an exactly antiparallel pair).
 * centres == NULL: every atom of pos is a centre (the subPos == Pos branch; n_centres must equal n_pos): the
 * kernel then walks the atoms in cell order, which is faster than passing pos itself when the atoms are not
 * spatially ordered -- same results, written at each atom's own index.  Needs the cell list of
 * wol_cell_build(FP64) over pos; at most 64 neighbours per centre (more -> wol_status reports it).
 */
int wol_angles_fill(const void *centres, int32_t centre_dtype, const double *box, int32_t n_frames, int32_t n_pos,
                    int32_t n_centres, const int32_t nc[3], double edge_min, double low3, double high3, void *workspace,
                    size_t workspace_bytes, const uint32_t *offsets, double *angles, void *stream);

/*
 * Neighbour lists in CSR form: what allNearNeighbors / nearNeighbors (fortran/waterlib.f90:830-862, :710-743) mark in
 * their dense N x N / M x N logical matrices, for sizes where that matrix cannot exist (3.6 TiB at 10^6 waters).
 * For centre g = f * n_centres + i its neighbours j with lowcut^2 < r^2 <= highcut^2 (minimum image) are
 * indices[offsets[g] .. offsets[g + 1]), frame-local atom indices in ASCENDING order (the order of the reference's
 * boolean-mask gathers, structureLibs/water_properties.py:243,372).
 *   centres   : NULL = every atom of pos is a centre (allNearNeighbors; n_centres must equal n_pos), walked in cell order
 *   workspace : cell list of wol_cell_build(FP64) over pos with r_cell >= highcut
 *   offsets   : [n_frames * n_centres + 1] uint32; offsets[last] = number of pairs
 *   scratch   : at least (n_frames * n_centres / 2048 + 8) uint32, 8-byte aligned; offsets[last] = 0xFFFFFFFF when the total
 *               number of pairs does not fit 32 bits (split the batch)
 *   indices   : [capacity] int32; a centre whose segment ends beyond capacity is not written (call with capacity 0
 *               and indices NULL to get the offsets only, read offsets[last], allocate, call again)
 */
int wol_neighbors_csr(const void *centres, int32_t centre_dtype, const double *box, int32_t n_frames, int32_t n_pos,
                      int32_t n_centres, const int32_t nc[3], double edge_min, double lowcut, double highcut, void *workspace,
                      size_t workspace_bytes, uint32_t *offsets, uint32_t *scratch, int32_t *indices, int64_t capacity,
                      void *stream);

/*
 * np.histogram(x, bins=nbins, range=[lo, hi]) counts (ACCUMULATED into hist) plus the sums
 * tetrahedralMetrics takes over the window tet_lo <= x <= tet_hi (water_properties.py:328-335):
 * tet_sums[0] += count, [1] += sum cos(x pi/180), [2] += sum cos^2.  tet_sums may be NULL.
 */
int wol_histogram(const double *x, int64_t n, double lo, double hi, int32_t nbins, int64_t *hist, double tet_lo,
                  double tet_hi, double *tet_sums, void *stream);

/*
 * Dense neighbour matrix out[m][n] (int32 0/1) of nearNeighbors / allNearNeighbors
 * (fortran/waterlib.f90:710-743, :830-862): lowcut^2 < r^2 <= highcut^2 under the minimum image.
 * box[3] on the device; a negative edge disables wrapping on that axis as in the reference (:41).
 * O(m n) by construction -- the drop-in for small systems; large ones use wol_q3b_frames.
 */
int wol_neighbor_matrix(const void *sub, int32_t sub_dtype, int32_t m, const void *pos, int32_t pos_dtype, int32_t n,
                        const double *box, double lowcut, double highcut, int32_t *out, void *stream);

/*
 * mode 0: reimage (fortran/waterlib.f90:32-47): out[n][3] = ref + minimg(pos - ref)
 * mode 1: lsiDists (fortran/waterlib.f90:900-918): out[n] = |minimg(pos - ref)|
 */
int wol_reimage(const double *pos, int32_t n, const double *ref, const double *box, double *out, int32_t mode,
                void *stream);

/* tetraCosAng (fortran/waterlib.f90:867-895): out[k][k] angles in degrees about `ref`; the diagonal,
 * which the Fortran leaves unwritten, is set to 0. */
int wol_tetracosang(const double *ref, const double *neigh, int32_t k, const double *box, double *out, void *stream);

/*
 * Local structure index, getLSI (structureLibs/water_properties.py:252-311; distances: lsiDists,
 * fortran/waterlib.f90:900-918).  For each centre with more than one neighbour inside (lowcut, highcut] and at
 * least one in the next shell (highcut, highcut + 3.7]: the population variance of the gaps between the sorted
 * minimum-image distances of those neighbours plus the next-shell atom with the smallest NON-periodic distance
 * (the reference's choice, :289).  num[f][i] = number of gaps (0: the centre has no value, lsi = 0).
 * workspace: cell list over pos with r_cell >= highcut + 3.7.  centres == NULL: every atom of pos is a centre
 * (n_centres must equal n_pos), walked in cell order.  At most 47 neighbours
 * inside highcut (more -> wol_status reports it).
 */
int wol_lsi(const void *centres, int32_t centre_dtype, const double *box, int32_t n_frames, int32_t n_pos, int32_t n_centres,
            const int32_t nc[3], double edge_min, double lowcut, double highcut, void *workspace, size_t workspace_bytes,
            double *lsi, int32_t *num, void *stream);

/*
 * K3: hydrogen bonds, generalHbonds (fortran/waterlib.f90:1156-1210) + AngBetween (:954-965), reduced to
 * the sums hbCalc takes (structureLibs/orderParam_lib.py:867-884) and, optionally, the dense matrix or
 * the bonded pair list.  The cell list in `workspace` must have been built (FP64) over the DONOR heavy
 * atoms: wol_cell_build(don, ..., n_pos = n_don, ...) with r_cell >= dist_cut.
 *   bond(i, j)  <=>  1.0E-2f < r^2(acc_i, don_j) <= dist_cut^2  and  angle(acc_i - H_j, don_j - H_j) >= ang_cut
 */
typedef struct wol_hbond_args {
    uint32_t struct_size;
    int32_t n_frames;
    int32_t n_acc;
    int32_t n_don; /* donor heavy atoms == donor hydrogens (the reference stops otherwise, :1171-1174) */
    int32_t acc_dtype;
    int32_t donh_dtype;
    const void *acc;  /* [n_frames][n_acc][3] */
    const void *donh; /* [n_frames][n_don][3], hydrogen j belongs to donor heavy atom j */
    const double *box;
    void *workspace;
    size_t workspace_bytes;
    int32_t nc[3];
    uint32_t pair_capacity;
    double edge_min;
    double dist_cut, ang_cut;
    int32_t *acc_count;     /* [n_frames][n_acc] written                                   (may be NULL) */
    int32_t *don_count;     /* [n_frames][n_don] ACCUMULATED with atomics                 (may be NULL) */
    int32_t *dense;         /* [n_frames][n_acc][n_don], bonded entries set to 1          (may be NULL) */
    int32_t *pairs;         /* [pair_capacity][2] (frame * n_acc + acceptor, donor)       (may be NULL) */
    uint32_t *pair_counter; /* number of pairs found (may exceed pair_capacity: list truncated)          */
} wol_hbond_args;

int wol_hbond_counts(const wol_hbond_args *args, void *stream);

/* H-bond locations of HBondsGeneral (structureLibs/water_properties.py:709-714): for each row of a pair
 * list from wol_hbond_counts, out[p][3] = 0.5 * (reimage(H_donor, acceptor) + acceptor). */
int wol_hbond_locations(const int32_t *pairs, int32_t n_pairs, const void *acc, int32_t acc_dtype, int32_t n_acc,
                        const void *donh, int32_t donh_dtype, int32_t n_don, const double *box, double *out, void *stream);

/*
 * K4: hydration-shell selection (structureLibs/orderParam_lib.py:495-498 over nearNeighbors,
 * fortran/waterlib.f90:710-743): mask[f][j] = 1 iff some solute atom lies within (lowcut, cutoff] of
 * atom j of the cell list (built over the waters).  The caller zero-fills mask.
 */
int wol_shell_mask(const void *sol, int32_t sol_dtype, int32_t n_sol, const double *box, int32_t n_frames, int32_t n_pos,
                   const int32_t nc[3], double edge_min, double lowcut, double cutoff, void *workspace,
                   size_t workspace_bytes, int32_t *mask, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Slab / interface routines (BASELINE config 4) and the all-Fortran triplet histogram.  One frame per call.
 * ------------------------------------------------------------------------------------------------ */

/*
 * K5: Willard-Chandler density, WillardDensityField / WillardDensityPoints (fortran/waterlib.f90:1286-1341,
 * :1351-1398; called at structureLibs/surface_library.py:197): at every point the sum over waters of a
 * Gaussian of width smoothlen truncated and shifted to zero at 3 smoothlen, and the normalised gradient.
 *   points == NULL : the nx * ny * nz grid points (gridx[i], gridy[j], gridz[k]); outputs [nx][ny][nz] and
 *                    [nx][ny][nz][3] row-major.  Otherwise n_points explicit points [n_points][3].
 *   workspace      : cell list of wol_cell_build(FP64) over the n_pos water oxygens with r_cell >= 3 smoothlen.
 * densnorms may be NULL.  A point farther than 3 smoothlen from every water gets a NaN normal (0/0), as in
 * the reference.  The terms are summed in cell order rather than atom order (differences ~1e-16 relative).
 */
int wol_willard_density(const double *points, int64_t n_points, const double *gridx, const double *gridy, const double *gridz,
                        int32_t nx, int32_t ny, int32_t nz, const double *box, int32_t n_pos, const int32_t nc[3], double edge_min,
                        double smoothlen, void *workspace, size_t workspace_bytes, double *densvals, double *densnorms,
                        void *stream);

/*
 * DensityField (fortran/waterlib.f90:1219-1268; called at structureLibs/surface_library.py:239): at every grid point the
 * number of waters inside the cube of edge binwidth (= gridx[1] - gridx[0], passed by the caller) centred on it, faces
 * included, minimum image, divided by binwidth**3.0.  densvals [nx][ny][nz] row-major.  workspace: cell list over the
 * waters with r_cell >= binwidth / 2.
 */
int wol_density_field(const double *gridx, const double *gridy, const double *gridz, int32_t nx, int32_t ny, int32_t nz, double binwidth,
                      const double *box, int32_t n_pos, const int32_t nc[3], double edge_min, void *workspace, size_t workspace_bytes,
                      double *densvals, void *stream);

/*
 * Iso-surface vertices of a scalar field on a rectilinear grid: for every grid edge from node (i, j, k) to (i+1, j, k),
 * (i, j+1, k) and (i, j, k+1) whose end values va, vb straddle `level` ((va > level) != (vb > level)) the point
 * a + t (b - a), t = (level - va) / (vb - va).  This is the vertex set of the marching-cubes mesh densityGrid takes from
 * skimage.measure.marching_cubes (structureLibs/surface_library.py:202; un-vendored, version unpinned, so parity is
 * pinned on this rule, not on skimage); together with wol_willard_density in points mode (the normals at the
 * vertices) it feeds wol_interface_water without leaving the device.  No periodic wrap of the grid, as in skimage.
 *   points[capacity][3] : vertices in ascending node index ((i * ny + j) * nz + k), then axis x, y, z; only the
 *                         first `capacity` are written (capacity 0 / points NULL: count only)
 *   n_total             : device int32, the number of crossing edges (may exceed capacity)
 *   scratch             : wol_iso_scratch_bytes(nx, ny, nz) bytes of device memory, 16-byte aligned
 */
size_t wol_iso_scratch_bytes(int32_t nx, int32_t ny, int32_t nz);
int wol_iso_points(const double *densvals, const double *gridx, const double *gridy, const double *gridz, int32_t nx, int32_t ny,
                   int32_t nz, double level, uint32_t *scratch, size_t scratch_bytes, double *points, int64_t capacity,
                   int32_t *n_total, void *stream);

/*
 * Triangles of the same iso-surface (structureLibs/surface_library.py:202 takes verts, faces, normals, values from
 * skimage.measure.marching_cubes; skimage is not vendored, so the triangulation rule is this library's own, generated by
 * scripts/make_mc_table.py: PARITY UNPINNED, the surface is pinned by its properties -- watertight, consistently oriented
 * towards lower values, vertices = wol_iso_points').  Call after wol_iso_points has filled `vertex_scratch` for the same
 * field and level (its per-node vertex offsets are reused).
 *   faces[capacity][3]  int32 vertex indices into wol_iso_points' output, ordered by cube index; capacity 0: count only
 *   areas[capacity]     optional, needs `points`: the reference's triangleArea of each face (fortran/imagelib.f90:254-267,
 *                       which returns |v1 x v2|, twice the geometric area)
 *   n_total             device int32, number of faces
 *   face_scratch        wol_iso_face_scratch_bytes(nx, ny, nz) bytes of device memory, 16-byte aligned
 */
size_t wol_iso_face_scratch_bytes(int32_t nx, int32_t ny, int32_t nz);
int wol_iso_faces(const double *densvals, int32_t nx, int32_t ny, int32_t nz, double level, const uint32_t *vertex_scratch,
                  uint32_t *face_scratch, size_t face_scratch_bytes, int32_t *faces, int64_t capacity, const double *points,
                  double *areas, int32_t *n_total, void *stream);

/*
 * InterfaceWater (fortran/waterlib.f90:1414-1469): for every water the nearest interface point (0-based,
 * first index on ties, -1 if none within distance^2 < 1000 -- the Fortran leaves that entry unwritten) and
 * its signed depth allwatdists = (water - point) . normal; for every interface point the nearest water;
 * numwater (ACCUMULATED, caller zero-fills) = waters with depth <= cutoff.  surfclose / numwater may be NULL.
 * Waters without an interface point in range get depth 0 and are not counted.
 */
int wol_interface_water(const double *pos, int32_t n_pos, const double *gridpos, const double *gridnorm, int32_t n_grid,
                        double cutoff, const double *box, int32_t *watclose, int32_t *surfclose, int32_t *numwater,
                        double *allwatdists, void *stream);

/*
 * Depth-binned profile of a per-water observable (config 4: q against the interface depth).  The reference
 * has the ingredients but no such function; bin = floor((coord - lo) / width), values
 * outside [0, nbins) are skipped.  count / sum / sumsq are ACCUMULATED.
 */
int wol_profile_bins(const double *value, const double *coord, int64_t n, double lo, double width, int32_t nbins, int64_t *count,
                     double *sum, double *sumsq, void *stream);

/*
 * Angle-bin table for the Fortran ceiling rule bin = ceiling(angle / ang_width) (fortran/waterlib.f90:1584),
 * same layout and construction as wol_angle_table (nbins + 1 + WOL_TABLE_EXTRA doubles, host-only).
 */
int wol_angle_table_ceil(double ang_width, int32_t nbins, double *table_host);

/*
 * histrr3b (fortran/waterlib.f90:1550-1593): histogram over triplets (i; j < k) of (distance i-j, distance
 * i-k, angle j-i-k) with ceiling binning; hist[d_num][d_num][a_num] int64, ACCUMULATED.  workspace: cell
 * list over the n_pos atoms with r_cell >= d_num * dist_width; angle_table: device copy of
 * wol_angle_table_ceil.  Triplets that index bin 0 in the Fortran (coincident atoms, 0 or -180 degree
 * angles; an out-of-bounds write there) are skipped.  At most 192 neighbours per atom inside the range.
 */
int wol_histrr3b(const double *box, int32_t n_pos, const int32_t nc[3], double edge_min, double dist_width, int32_t d_num,
                 double ang_width, int32_t a_num, const double *angle_table, void *workspace, size_t workspace_bytes, int64_t *hist,
                 void *stream);

/* ------------------------------------------------------------------------------------------------
 * Same-sweep observables.
 * ------------------------------------------------------------------------------------------------ */

/*
 * Pair-distance histograms with the Fortran binning nbin = ceiling(dist / binwidth):
 *   mode 0  RadialDist            (fortran/waterlib.f90:193-231)  outer = Pos2, cell list over Pos1, every pair
 *   mode 1  RadialDistSame        (fortran/waterlib.f90:316-353)  cell list over Pos, each unordered pair once (the i < j
 *                                                                 loops); the atoms are taken from the cell list, in cell
 *                                                                 order: `outer` must be that same Pos (n_outer == n_inner)
 *                                                                 and is not read
 *   mode 2  PairDistanceHistogram (fortran/waterlib.f90:358-389)  outer = Pos1, cell list over Pos2, 3-D
 * counts[totbins] int64, ACCUMULATED (the g(r) normalisation of the first two, O(totbins), is left to the caller:
 * counts(k) / (N * BulkDens * (4./3.) * pi * binwidth**3 * (k**3 - (k-1)**3)) with the Fortran's single-precision
 * 4./3. and truncated pi).  Pairs at distance exactly 0 index bin 0 in the Fortran (out of bounds) and are skipped.
 * workspace: cell list (FP64) over the n_inner atoms with r_cell >= totbins * binwidth.
 */
int wol_pair_hist(int32_t mode, const void *outer, int32_t outer_dtype, int32_t n_outer, const double *box, int32_t n_inner,
                  const int32_t nc[3], double edge_min, double binwidth, int32_t totbins, void *workspace, size_t workspace_bytes,
                  int64_t *counts, void *stream);

/*
 * getOrderParamPsi (structureLibs/water_properties.py:393-433): per centre, | mean over pairs of neighbours inside
 * (lowcut, highcut] of cos(6 theta) | -- the reference stores its complex mean of exp(6 i theta) into a real array,
 * which keeps only the real part, and that behaviour is reproduced; 0 with fewer than two neighbours.
 * workspace: cell list over pos with r_cell >= highcut.  At most 320 neighbours per centre.
 */
int wol_psi(const void *centres, int32_t centre_dtype, const double *box, int32_t n_frames, int32_t n_pos, int32_t n_centres,
            const int32_t nc[3], double edge_min, double lowcut, double highcut, void *workspace, size_t workspace_bytes, double *psi,
            void *stream);

/*
 * Connected components of a SYMMETRIC 0/1 adjacency matrix adj[n][n] (int32, as wol_neighbor_matrix and the residue
 * H-bond matrices of getHBClusterStats produce): labels[i] = smallest vertex index of i's component.  Replaces the
 * recursive depthFirstSort (fortran/sortlib.f90:26-72) behind getClusters (structureLibs/orderParam_lib.py:123-156).
 * `changed`: one int32 of device scratch.  Synchronises the stream (the sweep count depends on the graph).
 */
int wol_components(const int32_t *adj, int32_t n, int32_t *labels, int32_t *changed, void *stream);

/*
 * watOrient (fortran/waterlib.f90:973-1011; called at structureLibs/water_properties.py:605,636): per water the angle
 * in degrees (AngBetween, :954-965) between refvec and the dipole direction (sum of the two minimum-imaged O->H
 * vectors, imaged once more) and between refvec and the normal of the molecular plane (cross product of the O->H vectors).
 * opos [n_frames][n_waters][3], hpos [n_frames][2 n_waters][3] (H1, H2 of water i at rows 2i, 2i+1), box [n_frames][3]:
 * device, fp64.  refvec_host: 3 doubles on the HOST (normalised here like the Fortran does).  Out: [n_frames][n_waters].
 */
int wol_water_orient(const double *opos, const double *hpos, const double *box, int32_t n_frames, int32_t n_waters,
                     const double refvec_host[3], double *angdip, double *angplane, void *stream);

/*
 * binOnGrid (fortran/waterlib.f90:1047-1099; called at structureLibs/water_properties.py:668): number of atoms in every
 * cubic bin of the grid (left edge inclusive, atoms outside the grid ignored, no periodic wrap), counting only those inside
 * the sphere of diameter binwidth centred in the bin.  x/y/zbins: bin EDGES on the device (nx, ny, nz of them, uniform
 * spacing binwidth = xbins[1] - xbins[0] passed by the caller, who also checks that the bins are cubes -- the Fortran
 * STOPs otherwise).  outhist [nx-1][ny-1][nz-1] int32 row-major, overwritten.
 */
int wol_bin_on_grid(const double *opos, int64_t n, const double *xbins, const double *ybins, const double *zbins, int32_t nx, int32_t ny,
                    int32_t nz, double binwidth, int32_t *outhist, void *stream);

/*
 * RadialDistPlane (fortran/waterlib.f90:237-314; never called from the reference's Python): atoms of Pos2 inside the slab
 * |z'| <= 5 of the frame spanned by the three points Pos1, counted on a totbins x totbins grid of (x', y') with the
 * Fortran's ceiling binning.  pos1 [3][3], pos2 [n][3], box [3] (a negative edge = not periodic), counts
 * [totbins][totbins] int64 ACCUMULATED, counts[kx][ky] = the Fortran's rdf(kx + 1, ky + 1).  The Fortran writes out of
 * bounds for any slab atom with x' <= 0 or y' <= 0 (bin <= 0); here those atoms are skipped and their number is added to
 * *n_out_of_bounds (device int32) -- non-zero means the reference's result for this input is undefined.
 */
int wol_radial_dist_plane(const double *pos1, const double *pos2, int32_t n_pos2, const double *box, double binwidth, int32_t totbins,
                          double bulkdens, int64_t *counts, int32_t *n_out_of_bounds, void *stream);

/*
 * np.histogram2d(x, y, bins=(xedges, yedges)) on the device (numpy's histogramdd rule: searchsorted from the right, the
 * last edge belongs to the last bin, samples outside are dropped): the (theta, N_c) histogram of
 * threeBodyCalc(output2D=True), structureLibs/orderParam_lib.py:1385-1393.  x, y [n] double; edges ascending;
 * out [(n_xedges - 1)][(n_yedges - 1)] int64, ACCUMULATED.
 */
int wol_histogram2d(const double *x, const double *y, int64_t n, const double *xedges, int32_t n_xedges, const double *yedges,
                    int32_t n_yedges, int64_t *out, void *stream);

/*
 * Multi-GPU combine for frame sharding (the reference has no parallel path at all,
 * structureLibs/orderParam_lib.py:1312-1353 loops over frames in one process): in-place sum over the ranks of an NCCL
 * communicator the CALLER owns of `count` int64 histogram bins (WOL_SUM_I64: angle, q, H-bond histograms -- integer
 * sums, so the result is bit-identical for any number of ranks) or doubles (WOL_SUM_F64), enqueued on `stream`.
 * nccl_comm is an ncclComm_t.  NCCL is looked up at run time in the host process (libnccl.so.2 already loaded, e.g. by
 * torch, else the library search path): libwol.so does not link it.  WOL_ERR_UNSUPPORTED if none is found.
 */
int wol_hist_allreduce(void *nccl_comm, void *buf, size_t count, int32_t dtype, void *stream);

/*
 * Measurement aid (no counterpart in the reference): an FMA throughput probe for the FP roofline denominator
 * (SURVEY.md section 8d asks for a measured one).  Launches `blocks` blocks of 256 threads; every thread runs `iters`
 * rounds of 8 independent fused multiply-adds in the given dtype (WOL_F64 / WOL_F32) and adds its result to sink[0]
 * (a device double that only exists to keep the arithmetic alive).  The caller times the launch with events:
 * flop = 2 * 8 * iters * 256 * blocks.
 */
int wol_fma_probe(int32_t dtype, int32_t blocks, int32_t iters, double *sink, void *stream);

/* Number of kernel launches the last wol_* call on this thread enqueued (for bench bookkeeping). */
int wol_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* WOL_CAPI_H */
