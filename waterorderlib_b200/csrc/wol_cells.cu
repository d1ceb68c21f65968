// K1: cell-list build for a batch of frames (sm_100a).
//
//   count pass  : a block per tile of 1024 atoms, four atoms per thread; positions are staged through shared memory
//                 with 16-byte streaming loads (the whole 24 KB tile in flight before the first store) and binned;
//                 one atomicAdd per atom counts its cell.
//   scan_*      : exclusive prefix sum over all F*ncell counters (reduce / scan-of-sums / apply, the
//                 in-block part is a warp shuffle scan).
//   scatter pass: each atom is binned again and takes the next free place of its cell with an atomicAdd on
//                 the scanned counter, then writes its record (original coordinates, index, cell) and its
//                 box-wrapped float coordinates there with 16-byte stores.  No per-atom scratch between the
//                 passes: 24 + 24 B read and 48 B written per atom (fp64 positions and records).
//
// The reference has no counterpart: its neighbour search is the O(N^2) double loop of
// fortran/waterlib.f90:846-861 writing an N x N logical matrix.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "wol_device.cuh"
#include "wol_internal.h"
#include "wol_workspace.h"

namespace wol {

constexpr int kBuildThreads = 256;
constexpr int kBuildPerThread = 4;                                // atoms per thread
constexpr int kBuildTile = kBuildThreads * kBuildPerThread;      // atoms per block: 24 KB of fp64 positions in flight

// Stage `n_elems` consecutive position components (3 per atom) of a tile into shared memory.  16-byte vector loads
// when the tile base is 16-byte aligned (every load of the thread issued before the first store, so a block keeps its
// whole tile in flight), scalar (still coalesced) otherwise.
template <typename T>
__device__ __forceinline__ void stage_tile(const T *__restrict__ src, int n_elems, T *smem) {
    constexpr int kVec = 16 / sizeof(T);
    constexpr int kMaxVec = kBuildTile * 3 / kVec;                         // 16-byte words of a full tile
    constexpr int kPer = (kMaxVec + kBuildThreads - 1) / kBuildThreads;   // per thread
    const int tid = threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int n_vec = n_elems / kVec;
        const int4 *src4 = reinterpret_cast<const int4 *>(src);
        int4 *dst4 = reinterpret_cast<int4 *>(smem);
        int4 v[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kBuildThreads;
            if (i < n_vec) v[k] = __ldcs(src4 + i);  // streamed: read once per pass
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kBuildThreads;
            if (i < n_vec) dst4[i] = v[k];
        }
        for (int i = n_vec * kVec + tid; i < n_elems; i += kBuildThreads) smem[i] = src[i];
    } else {
        for (int i = tid; i < n_elems; i += kBuildThreads) smem[i] = src[i];
    }
}

struct BuildParams {
    const void *pos;
    const double *box;
    int n_frames, n_pos;
    int nc0, nc1, nc2;
    int tiles_per_frame;
    uint32_t *cell_start;
    void *recs;
    float4 *wrapped;
    uint32_t *cellpack;  // FP32 records only: packed cell coordinates per sorted atom (RecD carries them itself)
};

// T = storage type of the input positions, R = record type (RecD keeps doubles, RecF floats).
// SCATTER = false: bin + count into cell_start[cell + 1].  The exclusive scan over cell_start[0 .. ncells] (entry 0 is
//                  zero) then leaves the START of cell c in entry c + 1.
// SCATTER = true : bin again (same arithmetic, same cell) and take the next free place of the cell with an atomicAdd on
//                  entry c + 1, which thereby ends up as the cell's END = the start of cell c + 1: afterwards entry c is
//                  the start of cell c for every c, with no per-atom scratch (cell id, rank) written or re-read.
// A block owns a tile of kBuildTile consecutive atoms of one frame; a thread handles atoms t, t + 256, ... of it, all
// their atomics issued before the first record is written.
template <typename T, typename R, bool SCATTER>
__global__ void __launch_bounds__(kBuildThreads) cell_pass_kernel(BuildParams p) {
    __shared__ __align__(16) T s_pos[kBuildTile * 3];
    __shared__ double s_iL[3];
    __shared__ double s_L[3];
    const int f = blockIdx.x / p.tiles_per_frame;
    const int tile = blockIdx.x - f * p.tiles_per_frame;
    const int a0 = tile * kBuildTile;
    const int n_here = min(kBuildTile, p.n_pos - a0);
    const size_t frame_atom0 = (size_t)f * p.n_pos;
    if (threadIdx.x < 3) {
        const double L = p.box[(size_t)f * 3 + threadIdx.x];
        s_iL[threadIdx.x] = __ddiv_rn(1.0, L);
        s_L[threadIdx.x] = L;
    }
    stage_tile(reinterpret_cast<const T *>(p.pos) + (frame_atom0 + a0) * 3, n_here * 3, s_pos);
    __syncthreads();
    const size_t ncell = (size_t)p.nc0 * p.nc1 * p.nc2;
    uint32_t *const frame_counters = p.cell_start + (size_t)f * ncell + 1;
    const double iL0 = s_iL[0], iL1 = s_iL[1], iL2 = s_iL[2];
    // (the positions stay in shared memory and are read again after the atomics have returned: holding four atoms'
    // coordinates in registers across that wait costs the scatter pass half of its resident warps)
    int cellpack[kBuildPerThread];
    uint32_t dst[kBuildPerThread];
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) {
        const int t = threadIdx.x + k * kBuildThreads;
        cellpack[k] = -1;
        if (t < n_here) {
            const T x = s_pos[3 * t + 0], y = s_pos[3 * t + 1], z = s_pos[3 * t + 2];
            double xd = (double)x, yd = (double)y, zd = (double)z;
            if (sizeof(R) == sizeof(RecF)) {  // FP32 records: bin the value the sweep will see
                xd = (double)(float)x;
                yd = (double)(float)y;
                zd = (double)(float)z;
            }
            const int cx = cell_coord(xd, iL0, p.nc0);
            const int cy = cell_coord(yd, iL1, p.nc1);
            const int cz = cell_coord(zd, iL2, p.nc2);
            cellpack[k] = cx | (cy << 10) | (cz << 20);
            uint32_t *counter = frame_counters + (uint32_t)((cz * p.nc1 + cy) * p.nc0 + cx);
            if (!SCATTER) atomicAdd(counter, 1u);
            else dst[k] = atomicAdd(counter, 1u);
        }
    }
    if (!SCATTER) return;
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) {
        if (cellpack[k] < 0) continue;
        const int t = threadIdx.x + k * kBuildThreads;
        const T x = s_pos[3 * t + 0], y = s_pos[3 * t + 1], z = s_pos[3 * t + 2];
        if (sizeof(R) == sizeof(RecD)) {
            RecD r;
            r.x = (double)x;
            r.y = (double)y;
            r.z = (double)z;
            r.idx = a0 + t;
            r.cell = cellpack[k];
            // one 256-bit store (STG.256) per record: the 32-byte record is exactly one sector of L2, and the scatter
            // pass is bound by the number of L2 transactions, not by bytes
            {
                RecD *d = reinterpret_cast<RecD *>(p.recs) + dst[k];
                const long long tail = (long long)(unsigned)r.idx | ((long long)r.cell << 32);
                asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(d), "l"(__double_as_longlong(r.x)),
                             "l"(__double_as_longlong(r.y)), "l"(__double_as_longlong(r.z)), "l"(tail)
                             : "memory");
            }
            if (p.wrapped) {
                // box-wrapped coordinates for the float prefilter of the sweep: frac(x / L) * L
                float4 w;
                w.x = wrapped_coord((double)x, s_L[0], iL0);
                w.y = wrapped_coord((double)y, s_L[1], iL1);
                w.z = wrapped_coord((double)z, s_L[2], iL2);
                w.w = __int_as_float((int)dst[k]);  // fp64 records: the atom's place in the cell-sorted arrays (brick sweep)
                p.wrapped[dst[k]] = w;
            }
        } else {
            RecF r;
            r.x = (float)x;
            r.y = (float)y;
            r.z = (float)z;
            r.idx = a0 + t;
            *reinterpret_cast<int4 *>(reinterpret_cast<RecF *>(p.recs) + dst[k]) = *reinterpret_cast<const int4 *>(&r);
            if (p.wrapped) {
                // the FP32 sweep works on box-wrapped coordinates throughout (same binning input as the count pass)
                float4 w;
                w.x = wrapped_coord((double)r.x, s_L[0], iL0);
                w.y = wrapped_coord((double)r.y, s_L[1], iL1);
                w.z = wrapped_coord((double)r.z, s_L[2], iL2);
                w.w = __int_as_float(a0 + t);
                p.wrapped[dst[k]] = w;
                p.cellpack[dst[k]] = (uint32_t)cellpack[k];
            }
        }
    }
}

// ---- open (non-periodic) axes --------------------------------------------------------------------------------------
// The reference marks an axis as not periodic with a negative box edge (fortran/waterlib.f90:41: iBoxL = 0 there, so
// distvec - BoxL * anint(distvec * iBoxL) leaves the component untouched).  The cell-list kernels want a positive
// period on every axis, so an open axis gets an EQUIVALENT period L' = 2 (extent + reach) + 1: every difference
// between two atoms is below L' / 2, so anint(d / L') = 0 and d - L' * 0 = d bit for bit, and no periodic image comes
// closer than extent + 2 reach + 1 > reach.  extent = max - min of the coordinate over the frame's atoms (and centres).

__device__ __forceinline__ unsigned long long ordered_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);  // monotone in v
}
static double key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    memcpy(&v, &b, sizeof v);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(256) extent_kernel(const T *__restrict__ pos, int n_pos, unsigned long long *__restrict__ keys) {
    const int f = blockIdx.y;
    const T *p = pos + (size_t)f * n_pos * 3;
    double lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = Ops<double>::inf();
        hi[k] = -Ops<double>::inf();
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pos; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double v = (double)p[(size_t)i * 3 + k];
            lo[k] = fmin(lo[k], v);
            hi[k] = fmax(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(kFullMask, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(kFullMask, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(keys + (size_t)f * 6 + k, ordered_key(lo[k]));
            atomicMax(keys + (size_t)f * 6 + 3 + k, ordered_key(hi[k]));
        }
    }
}

static void launch_extent(const void *pos, int dtype, int n_frames, int n_pos, unsigned long long *keys, cudaStream_t stream) {
    if (n_pos < 1) return;
    int bx = (n_pos + 255) / 256;
    if (bx > 4 * sm_count()) bx = 4 * sm_count();
    const dim3 grid((unsigned)bx, (unsigned)n_frames);
    if (dtype == WOL_F64) extent_kernel<double><<<grid, 256, 0, stream>>>(reinterpret_cast<const double *>(pos), n_pos, keys);
    else extent_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float *>(pos), n_pos, keys);
    add_launches(1);
}

int effective_box(const void *pos, int pos_dtype, int n_frames, int n_pos, const void *centres, int centre_dtype, int n_centres,
                  const double *box_host, double reach, void *scratch_dev, double *box_out_host, cudaStream_t stream) {
    bool open = false;
    for (size_t i = 0; i < (size_t)n_frames * 3; ++i) {
        const double L = box_host[i];
        if (!(L == L) || L == 0.0 || isinf(L)) return set_error(WOL_ERR_INVALID, "box edge %zu is %g", i, L);
        box_out_host[i] = L;
        open |= L < 0.0;
    }
    if (!open) return WOL_OK;
    if (!scratch_dev) return set_error(WOL_ERR_INVALID, "wol_effective_box: open axes need %d bytes of device scratch", n_frames * 48);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(scratch_dev);
    // min slots start at the largest key, max slots at the smallest: two strided memsets
    cudaError_t e = cudaMemset2DAsync(keys, 48, 0xff, 24, (size_t)n_frames, stream);
    if (e == cudaSuccess) e = cudaMemset2DAsync(keys + 3, 48, 0x00, 24, (size_t)n_frames, stream);
    if (e != cudaSuccess) return set_cuda_error("wol_effective_box: memset", e);
    launch_extent(pos, pos_dtype, n_frames, n_pos, keys, stream);
    if (centres && n_centres > 0) launch_extent(centres, centre_dtype, n_frames, n_centres, keys, stream);
    unsigned long long *host = (unsigned long long *)malloc((size_t)n_frames * 48);
    if (!host) return set_error(WOL_ERR_INVALID, "wol_effective_box: out of host memory");
    e = cudaMemcpyAsync(host, keys, (size_t)n_frames * 48, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        free(host);
        return set_cuda_error("wol_effective_box: extent", e);
    }
    for (int f = 0; f < n_frames; ++f)
        for (int k = 0; k < 3; ++k) {
            if (!(box_host[(size_t)f * 3 + k] < 0.0)) continue;
            const double lo = key_to_double(host[(size_t)f * 6 + k]), hi = key_to_double(host[(size_t)f * 6 + 3 + k]);
            const double extent = (hi >= lo) ? hi - lo : 0.0;  // no atoms: any period will do
            if (!(extent == extent) || isinf(extent)) {
                free(host);
                return set_error(WOL_ERR_INVALID, "frame %d: coordinates along open axis %d are not finite", f, k);
            }
            box_out_host[(size_t)f * 3 + k] = 2.0 * (extent + reach) + 1.0;
        }
    free(host);
    return WOL_OK;
}

// ---- exclusive scan over n uint32 values, in place ------------------------------------------

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t *__restrict__ data, size_t n,
                                                                  uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t sum = 0;
    if (base + kScanItems <= n && ((reinterpret_cast<uintptr_t>(data + base) & 15u) == 0)) {
        const uint4 a = *reinterpret_cast<const uint4 *>(data + base);
        const uint4 b = *reinterpret_cast<const uint4 *>(data + base + 4);
        sum = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i)
            if (base + i < n) sum += data[base + i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < kScanThreads / 32; ++i) s += s_warp[i];
        block_sums[blockIdx.x] = s;
    }
}

// single block: exclusive scan of block_sums[0..nb) in place
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t *__restrict__ block_sums, int nb) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < nb) ? block_sums[i] : 0u;
        uint32_t inc = warp_inclusive_scan(v, lane);
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane];
            uint32_t winc = warp_inclusive_scan(w, lane);
            s_warp[lane] = winc - w;  // exclusive offset of each warp
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_warp[wid] + inc - v;
        if (i < nb) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t *__restrict__ data, size_t n,
                                                                 const uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    const bool vec = base + kScanItems <= n && ((reinterpret_cast<uintptr_t>(data + base) & 15u) == 0);
    if (vec) {
        const uint4 a = *reinterpret_cast<const uint4 *>(data + base);
        const uint4 b = *reinterpret_cast<const uint4 *>(data + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) v[i] = (base + i < n) ? data[base + i] : 0u;
    }
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) tsum += v[i];
    const uint32_t inc = warp_inclusive_scan(tsum, lane);
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = (lane < kScanThreads / 32) ? s_warp[lane] : 0u;
        uint32_t winc = warp_inclusive_scan(w, lane);
        if (lane < kScanThreads / 32) s_warp[lane] = winc - w;
    }
    __syncthreads();
    uint32_t run = block_sums[blockIdx.x] + s_warp[wid] + inc - tsum;
    uint32_t o[kScanItems];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        o[i] = run;
        run += v[i];
    }
    if (vec) {
        *reinterpret_cast<uint4 *>(data + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(data + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i)
            if (base + i < n) data[base + i] = o[i];
    }
}

int exclusive_scan_u32(uint32_t *data, size_t n, uint32_t *block_sums, cudaStream_t stream) {
    const int blocks = (int)((n + kScanTile - 1) / kScanTile);
    if (blocks < 1) return WOL_OK;
    scan_reduce_kernel<<<blocks, kScanThreads, 0, stream>>>(data, n, block_sums);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(block_sums, blocks);
    scan_apply_kernel<<<blocks, kScanThreads, 0, stream>>>(data, n, block_sums);
    add_launches(3);
    return WOL_OK;
}

template <typename T, typename R>
static void launch_passes(const BuildParams &p, unsigned blocks, bool scatter, cudaStream_t stream) {
    if (scatter)
        cell_pass_kernel<T, R, true><<<blocks, kBuildThreads, 0, stream>>>(p);
    else
        cell_pass_kernel<T, R, false><<<blocks, kBuildThreads, 0, stream>>>(p);
}

static void launch_pass(const BuildParams &p, unsigned blocks, bool scatter, int pos_dtype, int precision,
                        cudaStream_t stream) {
    if (pos_dtype == WOL_F64) {
        if (precision == WOL_PREC_FP64) launch_passes<double, RecD>(p, blocks, scatter, stream);
        else launch_passes<double, RecF>(p, blocks, scatter, stream);
    } else {
        if (precision == WOL_PREC_FP64) launch_passes<float, RecD>(p, blocks, scatter, stream);
        else launch_passes<float, RecF>(p, blocks, scatter, stream);
    }
}

int cell_build_launch(const void *pos, int pos_dtype, const double *box, int n_frames, int n_pos, const int32_t nc[3],
                      int precision, void *workspace, const WorkspaceLayout &lay, cudaStream_t stream) {
    char *ws = reinterpret_cast<char *>(workspace);
    BuildParams p;
    p.pos = pos;
    p.box = box;
    p.n_frames = n_frames;
    p.n_pos = n_pos;
    p.nc0 = nc[0];
    p.nc1 = nc[1];
    p.nc2 = nc[2];
    p.tiles_per_frame = (n_pos + kBuildTile - 1) / kBuildTile;
    p.cell_start = reinterpret_cast<uint32_t *>(ws + lay.off_cell_start);
    p.recs = ws + lay.off_recs;
    p.wrapped = reinterpret_cast<float4 *>(ws + lay.off_wrapped);
    // FP32 records fill only the first half of the record region; the packed cells of the sorted atoms follow
    p.cellpack = reinterpret_cast<uint32_t *>(ws + lay.off_recs + (size_t)lay.n_atoms_total * sizeof(RecF));
    uint32_t *block_sums = reinterpret_cast<uint32_t *>(ws + lay.off_block_sums);
    const size_t n_scan = (size_t)lay.n_cells_total + 1;

    cudaError_t e = cudaMemsetAsync(p.cell_start, 0, n_scan * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return set_cuda_error("cudaMemsetAsync(cell counters)", e);
    const long long blocks = (long long)p.tiles_per_frame * n_frames;
    if (blocks > 0x7fffffffLL) return set_error(WOL_ERR_RANGE, "too many atom tiles for one launch");
    if (blocks > 0) {
        launch_pass(p, (unsigned)blocks, false, pos_dtype, precision, stream);
        add_launches(1);
    }
    exclusive_scan_u32(p.cell_start, n_scan, block_sums, stream);
    if (blocks > 0) {
        launch_pass(p, (unsigned)blocks, true, pos_dtype, precision, stream);
        add_launches(1);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("cell build launch", e);
    return WOL_OK;
}

}  // namespace wol
