// Entry points of libwol.so (see include/wol_capi.h): argument validation, error strings, the grid
// plan, the angle-bin table and the launches.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "wol_internal.h"

namespace wol {

static thread_local char g_error[512] = "";
static thread_local int g_launches = 0;

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int set_cuda_error(const char *what, cudaError_t e) {
    snprintf(g_error, sizeof(g_error), "%s: %s", what, cudaGetErrorString(e));
    return WOL_ERR_CUDA;
}

void add_launches(int n) { g_launches += n; }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

// ---- angle-bin table ----------------------------------------------------------------------------

// CosAngle3's tail (fortran/waterlib.f90:699-702) on the host libm.
static double ref_angle_deg(double c) {
    const double pi = 3.1415926535897931;
    volatile double phi = acos(c);
    volatile double a = fmod(phi + pi, pi * 2.0) - pi;
    if (a < -pi) a = a + pi * 2.0;
    return a * (180.0 / pi);
}

// np.histogram's uniform-bin rule as a position: -1 below lo, nbins above hi
static int ref_position(double x, double lo, double hi, int nbins) {
    if (!(x >= lo)) return -1;
    if (!(x <= hi)) return nbins;
    const double denom = hi - lo;
    volatile double f = ((x - lo) / denom) * (double)nbins;
    long idx = (long)f;
    if (idx == nbins) idx -= 1;
    const double step = denom / (double)nbins;
    volatile double e_lo = (double)idx * step + lo;
    if (x < e_lo) {
        idx -= 1;
    } else if (idx != nbins - 1) {
        volatile double e_hi = (idx + 1 == nbins) ? hi : (double)(idx + 1) * step + lo;
        if (x >= e_hi) idx += 1;
    }
    return (int)idx;
}

// order-preserving map between doubles and int64
static long long d2o(double d) {
    long long i;
    memcpy(&i, &d, 8);
    return i >= 0 ? i : (long long)0x8000000000000000ULL - i;
}
static double o2d(long long o) {
    long long i = o >= 0 ? o : (long long)0x8000000000000000ULL - o;
    double d;
    memcpy(&d, &i, 8);
    return d;
}

// largest c in (-1, 1] with pred(c) true, for a predicate that is true on a lower interval; -2 if none
template <typename F>
static double last_true(F pred) {
    long long lo = d2o(nextafter(-1.0, 0.0)), hi = d2o(1.0);
    if (!pred(o2d(lo))) return -2.0;
    if (pred(o2d(hi))) return 1.0;
    while (hi - lo > 1) {
        const long long mid = lo + (hi - lo) / 2;
        if (pred(o2d(mid))) lo = mid; else hi = mid;
    }
    return o2d(lo);
}

double angle_cos_threshold(double ang_deg, int *minus_one_passes) {
    if (minus_one_passes) *minus_one_passes = ref_angle_deg(-1.0) >= ang_deg ? 1 : 0;
    return last_true([&](double c) { return ref_angle_deg(c) >= ang_deg; });
}

// FMA throughput probe (wol_fma_probe): 8 independent chains per thread keep the pipe full at any occupancy
template <typename T>
__global__ void __launch_bounds__(256) fma_probe_kernel(int iters, double *sink) {
    T a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (T)(threadIdx.x + k) * (T)1e-3;
    const T m = (T)0.999999, c = (T)1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = a[k] * m + c;
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == (T)-1) atomicAdd(sink, (double)s);  // never true; keeps the loop alive
}

}  // namespace wol

using namespace wol;

extern "C" {

const char *wol_version(void) { return "waterorderlib_b200 0.1 (sm_100a)"; }
const char *wol_last_error(void) { return g_error; }
int wol_abi_version(void) { return WOL_ABI_VERSION; }
// ---- multi-GPU combine (frames sharded over ranks) ------------------------------------------------------
// NCCL is resolved at run time from the library the host process already uses (torch's, an MPI program's): libwol.so
// itself links only the CUDA runtime.

typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*nccl_errstr_fn)(int);

static void *nccl_symbol(const char *name) {
    static const char *const kNames[] = {"libnccl.so.2", "libnccl.so"};
    for (int pass = 0; pass < 2; ++pass)      // first a library that is already loaded, then the default search path
        for (const char *lib : kNames) {
            void *h = dlopen(lib, pass == 0 ? (RTLD_NOW | RTLD_NOLOAD) : (RTLD_NOW | RTLD_GLOBAL));
            if (!h) continue;
            if (void *sym = dlsym(h, name)) return sym;
        }
    return dlsym(RTLD_DEFAULT, name);         // statically linked into the host, or loaded under another name
}

int wol_hist_allreduce(void *nccl_comm, void *buf, size_t count, int32_t dtype, void *stream) {
    if (!nccl_comm || (!buf && count > 0)) return set_error(WOL_ERR_INVALID, "wol_hist_allreduce: null argument");
    if (dtype != WOL_SUM_I64 && dtype != WOL_SUM_F64) return set_error(WOL_ERR_INVALID, "wol_hist_allreduce: unknown dtype %d", dtype);
    if (count == 0) return WOL_OK;
    static nccl_allreduce_fn allreduce = reinterpret_cast<nccl_allreduce_fn>(nccl_symbol("ncclAllReduce"));
    if (!allreduce) return set_error(WOL_ERR_UNSUPPORTED, "wol_hist_allreduce: no NCCL library (libnccl.so.2) in this process or on the search path");
    // ncclDataType_t: ncclInt64 = 4, ncclFloat64 = 8; ncclRedOp_t: ncclSum = 0 (stable across NCCL 2.x)
    const int rc = allreduce(buf, buf, count, dtype == WOL_SUM_I64 ? 4 : 8, 0, nccl_comm, (cudaStream_t)stream);
    if (rc != 0) {
        static nccl_errstr_fn errstr = reinterpret_cast<nccl_errstr_fn>(nccl_symbol("ncclGetErrorString"));
        return set_error(WOL_ERR_CUDA, "ncclAllReduce failed: %s", errstr ? errstr(rc) : "unknown NCCL error");
    }
    return WOL_OK;
}

int wol_last_launch_count(void) { return g_launches; }

int wol_fma_probe(int32_t dtype, int32_t blocks, int32_t iters, double *sink, void *stream) {
    if (!sink || blocks < 1 || iters < 1) return set_error(WOL_ERR_INVALID, "wol_fma_probe: need a sink, blocks >= 1, iters >= 1");
    if (dtype == WOL_F64) fma_probe_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    else if (dtype == WOL_F32) fma_probe_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    else return set_error(WOL_ERR_INVALID, "wol_fma_probe: unknown dtype %d", dtype);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wol_fma_probe", e);
    g_launches = 1;
    return WOL_OK;
}

int wol_plan_grid(const double *box_host, int32_t n_frames, double r_cell, int32_t nc_out[3], double *edge_min_out,
                  double *box_max_out) {
    if (!box_host || !nc_out || n_frames < 1) return set_error(WOL_ERR_INVALID, "wol_plan_grid: null argument or no frames");
    if (!(r_cell > 0.0)) return set_error(WOL_ERR_INVALID, "wol_plan_grid: r_cell must be positive");
    double lmin[3], lmax = 0.0;
    for (int k = 0; k < 3; ++k) lmin[k] = INFINITY;
    for (int f = 0; f < n_frames; ++f)
        for (int k = 0; k < 3; ++k) {
            const double L = box_host[(size_t)f * 3 + k];
            if (!(L > 0.0) || !isfinite(L))
                return set_error(WOL_ERR_UNSUPPORTED,
                                 "frame %d: box edge %d is %g; for non-periodic (negative) axes pass the box through wol_effective_box first", f, k, L);
            if (L < lmin[k]) lmin[k] = L;
            if (L > lmax) lmax = L;
        }
    double emin = INFINITY;
    for (int k = 0; k < 3; ++k) {
        double c = floor(lmin[k] / (r_cell * (1.0 + 1e-9)));
        if (c < 1.0) c = 1.0;
        if (c > 1024.0) c = 1024.0;
        nc_out[k] = (int32_t)c;
        const double e = lmin[k] / c;
        if (e < emin) emin = e;
    }
    if (edge_min_out) *edge_min_out = emin;
    if (box_max_out) *box_max_out = lmax;
    return WOL_OK;
}

int wol_effective_box(const void *pos, int32_t pos_dtype, int32_t n_frames, int32_t n_pos, const void *centres, int32_t centre_dtype,
                      int32_t n_centres, const double *box_host, double reach, void *scratch_dev, double *box_out_host, void *stream) {
    if (!box_host || !box_out_host || n_frames < 1) return set_error(WOL_ERR_INVALID, "wol_effective_box: null argument or no frames");
    if (n_pos < 0 || n_centres < 0 || (n_pos > 0 && !pos)) return set_error(WOL_ERR_INVALID, "wol_effective_box: bad positions");
    if ((pos_dtype != WOL_F64 && pos_dtype != WOL_F32) || (centres && centre_dtype != WOL_F64 && centre_dtype != WOL_F32))
        return set_error(WOL_ERR_INVALID, "wol_effective_box: unknown dtype");
    if (!(reach >= 0.0) || isinf(reach)) return set_error(WOL_ERR_INVALID, "wol_effective_box: reach must be a finite distance");
    g_launches = 0;
    return effective_box(pos, pos_dtype, n_frames, n_pos, centres, centre_dtype, n_centres, box_host, reach, scratch_dev, box_out_host,
                         (cudaStream_t)stream);
}

size_t wol_workspace_bytes(int32_t n_frames, int32_t n_pos, int32_t n_centres_max, const int32_t nc[3]) {
    if (n_frames < 0 || n_pos < 0 || !nc) return 0;
    return workspace_layout(n_frames, n_pos, n_centres_max, nc).total;
}

static int check_shape(int32_t n_frames, int32_t n_pos, int32_t n_centres, const int32_t nc[3], const void *workspace,
                       size_t workspace_bytes, WorkspaceLayout *lay) {
    if (n_frames < 1 || n_pos < 0 || n_centres < 0) return set_error(WOL_ERR_INVALID, "negative size");
    if (!nc || nc[0] < 1 || nc[1] < 1 || nc[2] < 1) return set_error(WOL_ERR_INVALID, "cell grid must be at least 1x1x1");
    const long long big = (long long)n_frames * (n_pos > n_centres ? n_pos : n_centres);
    if (big >= (1LL << 30)) return set_error(WOL_ERR_RANGE, "n_frames * n_atoms = %lld exceeds 2^30; use smaller batches", big);
    if ((long long)nc[0] * nc[1] * nc[2] * n_frames >= (1LL << 31)) return set_error(WOL_ERR_RANGE, "too many cells");
    if (!workspace) return set_error(WOL_ERR_INVALID, "null workspace");
    if (((uintptr_t)workspace & 255u) != 0) return set_error(WOL_ERR_INVALID, "workspace must be 256-byte aligned");
    *lay = workspace_layout(n_frames, n_pos, n_centres, nc);
    if (workspace_bytes < lay->total)
        return set_error(WOL_ERR_WORKSPACE, "workspace holds %zu bytes, %zu needed", workspace_bytes, lay->total);
    return WOL_OK;
}

int wol_cell_build(const void *pos, int32_t pos_dtype, const double *box, int32_t n_frames, int32_t n_pos,
                   const int32_t nc[3], int32_t precision, void *workspace, size_t workspace_bytes, void *stream) {
    g_launches = 0;
    if (!pos && n_pos > 0) return set_error(WOL_ERR_INVALID, "wol_cell_build: null positions");
    if (!box) return set_error(WOL_ERR_INVALID, "wol_cell_build: null box");
    if (pos_dtype != WOL_F64 && pos_dtype != WOL_F32) return set_error(WOL_ERR_INVALID, "unknown position dtype %d", pos_dtype);
    if (precision != WOL_PREC_FP64 && precision != WOL_PREC_FP32) return set_error(WOL_ERR_INVALID, "unknown precision %d", precision);
    WorkspaceLayout lay;
    // the workspace may have been sized for more centres than atoms: only the prefix up to the records
    // is touched here, so validate with n_centres = 0 and the caller's byte count
    int rc = check_shape(n_frames, n_pos, 0, nc, workspace, workspace_bytes, &lay);
    if (rc != WOL_OK) return rc;
    return cell_build_launch(pos, pos_dtype, box, n_frames, n_pos, nc, precision, workspace, lay, (cudaStream_t)stream);
}

int wol_angle_table(double hist_lo, double hist_hi, int32_t nbins, double tet_lo, double tet_hi, double *table_host) {
    if (!table_host || nbins < 1 || !(hist_hi > hist_lo))
        return set_error(WOL_ERR_INVALID, "wol_angle_table: need nbins >= 1, hi > lo and an output array");
    for (int k = 0; k <= nbins; ++k)
        table_host[k] = last_true([&](double c) { return ref_position(ref_angle_deg(c), hist_lo, hist_hi, nbins) >= k; });
    double *extra = table_host + nbins + 1;
    extra[0] = (double)ref_position(ref_angle_deg(-1.0), hist_lo, hist_hi, nbins);
    extra[1] = (double)ref_position(0.0, hist_lo, hist_hi, nbins);
    extra[2] = last_true([&](double c) { return ref_angle_deg(c) >= tet_lo; });
    // smallest c with angle <= tet_hi  ==  successor of the largest c with angle > tet_hi
    {
        const double below = last_true([&](double c) { return ref_angle_deg(c) > tet_hi; });
        extra[3] = (below == -2.0) ? nextafter(-1.0, 0.0) : nextafter(below, 2.0);
    }
    // self-check: thresholds decrease, and the predicate really flips at each of them (a few ulps
    // either side), i.e. the host acos behaved monotonically where it matters
    int ok = 1;
    for (int k = 0; k <= nbins && ok; ++k) {
        const double c = table_host[k];
        if (k > 0 && c > table_host[k - 1]) ok = 0;
        if (c == -2.0 || c == 1.0) continue;
        double lo = c, hi = c;
        for (int s = 0; s < 4 && ok; ++s) {
            hi = nextafter(hi, 2.0);
            if (hi <= 1.0 && ref_position(ref_angle_deg(hi), hist_lo, hist_hi, nbins) >= k) ok = 0;
            if (ref_position(ref_angle_deg(lo), hist_lo, hist_hi, nbins) < k) ok = 0;
            lo = nextafter(lo, -2.0);
            if (lo <= -1.0) break;
        }
    }
    extra[4] = (double)ok;
    extra[5] = extra[6] = extra[7] = 0.0;
    if (!ok) return set_error(WOL_ERR_UNSUPPORTED, "wol_angle_table: host acos is not monotone near a bin edge");
    return WOL_OK;
}

int wol_angle_table_ceil(double ang_width, int32_t nbins, double *table_host) {
    if (!table_host || nbins < 1 || !(ang_width > 0.0))
        return set_error(WOL_ERR_INVALID, "wol_angle_table_ceil: need nbins >= 1, a positive width and an output array");
    // 0-based bin of the Fortran rule bin = ceiling(x / width) (waterlib.f90:1584): -1 for x <= 0, nbins beyond
    auto position = [&](double c) {
        volatile double q = ref_angle_deg(c) / ang_width;
        const double t = ceil(q);
        if (!(t >= 1.0)) return -1;
        if (t > (double)nbins) return (int)nbins;
        return (int)t - 1;
    };
    for (int k = 0; k <= nbins; ++k) table_host[k] = last_true([&](double c) { return position(c) >= k; });
    double *extra = table_host + nbins + 1;
    extra[0] = (double)position(-1.0);  // the -180 CosAngle3 returns for an antiparallel pair: no bin
    extra[1] = -1.0;                    // 0 degrees (coincident positions): bin 0 of the Fortran, out of bounds
    extra[2] = extra[3] = 0.0;
    int ok = 1;
    for (int k = 1; k <= nbins && ok; ++k)
        if (table_host[k] > table_host[k - 1]) ok = 0;
    extra[4] = (double)ok;
    extra[5] = extra[6] = extra[7] = 0.0;
    if (!ok) return set_error(WOL_ERR_UNSUPPORTED, "wol_angle_table_ceil: host acos is not monotone near a bin edge");
    return WOL_OK;
}

int wol_q3b_frames(const wol_q3b_args *a, void *stream) {
    g_launches = 0;
    if (!a) return set_error(WOL_ERR_INVALID, "wol_q3b_frames: null args");
    if (a->struct_size != sizeof(wol_q3b_args))
        return set_error(WOL_ERR_INVALID, "wol_q3b_args.struct_size is %u, this library expects %zu", a->struct_size, sizeof(wol_q3b_args));
    if (a->precision != WOL_PREC_FP64 && a->precision != WOL_PREC_FP32) return set_error(WOL_ERR_INVALID, "unknown precision %d", a->precision);
    if (!a->box) return set_error(WOL_ERR_INVALID, "null box");
    if (a->centres && a->centre_dtype != WOL_F64 && a->centre_dtype != WOL_F32) return set_error(WOL_ERR_INVALID, "unknown centre dtype");
    if (!a->do_q && !a->do_3body) return set_error(WOL_ERR_INVALID, "nothing to do: both do_q and do_3body are 0");
    const int32_t n_centres = a->centres ? a->n_centres : a->n_pos;
    WorkspaceLayout lay;
    int rc = check_shape(a->n_frames, a->n_pos, n_centres, a->nc, a->workspace, a->workspace_bytes, &lay);
    if (rc != WOL_OK) return rc;
    if (!(a->edge_min > 0.0)) return set_error(WOL_ERR_INVALID, "edge_min must come from wol_plan_grid");
    if (a->do_3body) {
        if (!(a->high3 >= 0.0) || !(a->low3 >= 0.0) || !isfinite(a->high3))
            return set_error(WOL_ERR_INVALID, "three-body cutoffs must be finite and non-negative");
        if (!a->angle_table) return set_error(WOL_ERR_INVALID, "do_3body needs angle_table (wol_angle_table)");
        if (a->nbins < 1 || !(a->hist_hi > a->hist_lo)) return set_error(WOL_ERR_INVALID, "bad angle histogram spec");
        // a half-width-1 stencil must contain every neighbour inside high3 on every axis that is not
        // fully enumerated
        for (int k = 0; k < 3; ++k)
            if (a->nc[k] > 3 && a->high3 * (1.0 + 1e-9) > a->edge_min)
                return set_error(WOL_ERR_INVALID, "three-body cutoff %.6g exceeds the planned cell edge %.6g; re-plan with r_cell >= cutoff",
                                 a->high3, a->edge_min);
    }
    if (a->do_q) {
        if (!(a->highq >= 0.0) || !(a->lowq >= 0.0) || !isfinite(a->highq))
            return set_error(WOL_ERR_INVALID, "q cutoffs must be finite and non-negative");
        if (a->q_hist && a->q_nbins < 1) return set_error(WOL_ERR_INVALID, "bad q histogram spec");
    }
    if (n_centres == 0 || a->n_pos == 0) return WOL_OK;
    return q3b_launch(*a, lay, (cudaStream_t)stream);
}

int wol_status(const void *workspace, int32_t n_frames, int32_t n_pos, int32_t n_centres_max, const int32_t nc[3],
               void *stream, int32_t status_host[4]) {
    if (!workspace || !status_host || !nc) return set_error(WOL_ERR_INVALID, "wol_status: null argument");
    const WorkspaceLayout lay = workspace_layout(n_frames, n_pos, n_centres_max, nc);
    uint32_t c[kNumCounters];
    cudaError_t e = cudaMemcpyAsync(c, reinterpret_cast<const char *>(workspace) + lay.off_counters, sizeof(c),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(const_cast<char *>(reinterpret_cast<const char *>(workspace)) + lay.off_counters + kCntFatal * 4, 0, 4,
                            (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return set_cuda_error("wol_status", e);
    status_host[0] = (int32_t)c[kCntWidened];
    status_host[1] = (int32_t)c[kCntOverflow];
    status_host[2] = (int32_t)c[kCntFatal];
    status_host[3] = (int32_t)c[kCntSlowPair];  /* brick path: pairs re-evaluated in exact arithmetic */
    if (c[kCntFatal] != 0)
        return set_error(WOL_ERR_CAPACITY, "%u centres have more neighbours than the large-capacity path holds", c[kCntFatal]);
    return WOL_OK;
}

}  // extern "C"
