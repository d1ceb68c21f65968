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
 *     allocates device memory, never synchronises the host and keeps no state besides a
 *     thread-local error string.  All work is enqueued on `stream`.
 *   - return value: WOL_OK or a negative WOL_ERR_* code; wol_last_error() describes the failure.
 *     Nothing aborts the process (the reference's Fortran `stop`s do, waterlib.f90:1171-1174).
 *   - indices are 0-based int32; "no neighbour" is -1.
 */
#ifndef WOL_CAPI_H
#define WOL_CAPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WOL_ABI_VERSION 1

enum {
    WOL_OK = 0,
    WOL_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unknown enum)          */
    WOL_ERR_UNSUPPORTED = -2, /* valid in the reference but not implemented here (e.g. BoxL <= 0)  */
    WOL_ERR_WORKSPACE = -3,   /* workspace too small                                                */
    WOL_ERR_CUDA = -4,        /* a CUDA runtime call failed; message holds cudaGetErrorString       */
    WOL_ERR_RANGE = -5        /* sizes overflow the 32-bit indexing of the kernels                  */
};

enum { WOL_F64 = 0, WOL_F32 = 1 };            /* storage dtype of a position array               */
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

const char *wol_version(void);
const char *wol_last_error(void);
int wol_abi_version(void);

/*
 * Cell-grid plan for a batch of frames.  Picks nc[3] (cells per axis, identical for all frames of the
 * batch) such that every frame's cell edge L/nc is >= r_cell, and reports the stencil half-width
 * `w_out` for which w*edge >= r_complete on every axis.  Host-only, cheap.
 * Replaces nothing in the reference (its search is O(N^2), waterlib.f90:846-861).
 */
int wol_plan_grid(const double *box_host, int32_t n_frames, double r_cell, double r_complete,
                  int32_t nc_out[3], int32_t *w_out);

/* Bytes of scratch needed by wol_cell_build + the evaluation kernels for this batch shape.
 * n_centres_max = the largest number of centres any later call on this workspace will pass. */
size_t wol_workspace_bytes(int32_t n_frames, int32_t n_pos, int32_t n_centres_max, const int32_t nc[3]);

/*
 * K1: cell-list build (counting sort of atoms by cell, fixed-point periodic coordinates).
 *   pos        [n_frames][n_pos][3] of pos_dtype
 *   box        [n_frames][3] double, all > 0
 * Afterwards `workspace` holds the cell list consumed by the entry points below.
 */
int wol_cell_build(const void *pos, int32_t pos_dtype, const double *box, int32_t n_frames, int32_t n_pos,
                   const int32_t nc[3], void *workspace, size_t workspace_bytes, void *stream);

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
 *   ang_hist    [n_frames][nbins] int64, ACCUMULATED (+=): np.histogram(angles, nbins, [hist_lo, hist_hi])
 *   q_hist      [n_frames][q_nbins] int64, ACCUMULATED: np.histogram(q, q_nbins, [0, 1])
 *   frame_stats [n_frames][WOL_NSTATS] double, ACCUMULATED
 * `hist_frame_stride0` = 1 keeps one histogram row per frame; 0 folds all frames into row 0.
 */
typedef struct wol_q3b_args {
    uint32_t struct_size; /* sizeof(wol_q3b_args), for forward compatibility */
    int32_t precision;    /* WOL_PREC_FP64 | WOL_PREC_FP32 */
    const void *pos;
    int32_t pos_dtype;
    int32_t n_frames;
    int32_t n_pos;
    int32_t n_centres; /* ignored when centres == NULL (then n_centres = n_pos) */
    const void *centres;
    int32_t centre_dtype;
    int32_t hist_per_frame; /* 1: one histogram row per frame, 0: a single shared row */
    const double *box;
    const void *workspace; /* as filled by wol_cell_build for the same pos/box/nc */
    size_t workspace_bytes;
    int32_t nc[3];
    int32_t stencil_w; /* from wol_plan_grid */
    double low3, high3; /* getCosAngs lowCut/highCut (default 0, 3.413) */
    double lowq, highq; /* getOrderParamq lowCut/highCut (default 0, 10) */
    int32_t do_q;       /* evaluate the q branch */
    int32_t do_3body;   /* evaluate the three-body branch */
    int32_t nbins;      /* angle histogram bins (default 500) */
    int32_t q_nbins;    /* q histogram bins (default 500) */
    double hist_lo, hist_hi; /* angle histogram range (default 0, 180) */
    void *q;
    int32_t *nn_idx;
    int32_t *n3;
    int64_t *ang_hist;
    int64_t *q_hist;
    double *frame_stats;
} wol_q3b_args;

int wol_q3b_frames(const wol_q3b_args *args, void *stream);

/* Number of kernel launches the last wol_* call on this thread enqueued (for bench bookkeeping). */
int wol_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* WOL_CAPI_H */
