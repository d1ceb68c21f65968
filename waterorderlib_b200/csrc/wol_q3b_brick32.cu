// K2f, brick path of the FP32 arithmetic mode (sm_100a): large frames, every atom a centre.
//
// Same persistent CTA, producer warp and shared-memory stages as the fp64 brick kernel (wol_q3b_brick.cu /
// wol_q3b_brick.cuh), and the arithmetic of the thread-per-centre fp32 kernel (wol_q3b_tpc32.cu): the float distance of
// the sweep IS the distance of this mode, so there is no exact re-evaluation and NO global load on the hot path at
// all -- a centre's candidates, their coordinates and their atom indices all sit in the stage.
//
//   phase 1   sweep over the nine stencil rows (LDS.128 per candidate, 3 FADD + FMUL + 2 FFMA + compare); a survivor
//             (inside the reach = the larger of the three-body cutoff and the q selection radius) is one predicated
//             store of its stage slot.
//   phase 2   dense pass over the ~8 survivors: difference vector and distance again (same operations, same value),
//             cutoff tests low^2 < r^2 <= high^2 (fortran/waterlib.f90:737,855), three-body neighbours compacted into a
//             per-thread shared column (rotated inside the warp's 32 columns, see the fp64 kernel), register top-4 by
//             (distance, atom index).
//   phase 3   three-body pairs flattened over the warp (rsqrtf + acosf, CosAngle3's 0-degree / -180-degree rules),
//             q from the four winners (water_properties.py:379-388).
// North-star tolerance of the mode: 1e-4 on q and cosines (tests/test_gpu_fp32.py); neighbour selection and bin
// membership may differ from the fp64 reference where two values agree to ~1e-7.
#include "wol_q3b_brick.cuh"
#include "wol_q3b_f32.cuh"

namespace wol {

constexpr int kB32Producers = 1;           // (measured with kBkStages producers and 13 consumer warps: 1.60 ms against 1.43 ms
                                            // per 8M waters -- the two consumer warps are worth more than the producers)
constexpr int kB32Warps = 16 - kB32Producers;
constexpr int kB32Consumers = kB32Warps * 32;
constexpr int kB32Threads = kB32Consumers + 32 * kB32Producers;
constexpr int kB32AtomCap = 1536;
constexpr int kB32EntCap = 10;  // three-body neighbours per centre
constexpr int kB32MaxPairs = kB32EntCap * (kB32EntCap - 1) / 2;

struct B32Smem {
    static constexpr int kAtomCap = kB32AtomCap;
    static constexpr int kProducerRegs = 124;
    float4 loc[kBkStages][kB32AtomCap];
    float4 ent[kB32EntCap][kB32Consumers];       // (dx, dy, dz, r^2) of the three-body neighbours
    unsigned short lj[kBkListCap + 1][kB32Consumers];  // survivors' stage slots (+ one row that absorbs overflowing stores)
    unsigned short cs[kBkStages][kBkRowCap * kBkCsW];
    int woff[kB32Warps][33];
    BkItem item[kBkStages];
    BkRow prow[kB32Producers][kBkRowCap];
    double pbox[kB32Producers][6];
    unsigned long long bar_full[kBkStages], bar_raw[kBkStages], bar_empty[kBkStages];
    unsigned char pair_ab[kB32MaxPairs + 3];
    __device__ __forceinline__ BkRow *prow_of(int pid) { return prow[pid]; }
    __device__ __forceinline__ double *pbox_of(int pid) { return pbox[pid]; }
};

__device__ __forceinline__ void b32_consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kB32Consumers) : "memory"); }

__device__ __forceinline__ void b32_flush_bins(unsigned *s_bins, unsigned long long *g_bins, int nbins, bool clear, int tid) {
    for (int i = tid; i < nbins; i += kB32Consumers) {
        const unsigned v = s_bins[i];
        if (v) atomicAdd(g_bins + i, (unsigned long long)v);
        if (clear) s_bins[i] = 0u;
    }
}

__global__ void __launch_bounds__(kB32Threads, 1) q3b_brick32_kernel(const __grid_constant__ Q3bParams P, const __grid_constant__ BrickPlan B) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    B32Smem &S = *reinterpret_cast<B32Smem *>(smem_raw);
    unsigned *s_hist = reinterpret_cast<unsigned *>(smem_raw + sizeof(B32Smem));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do3 = P.do_3b != 0, doq = P.do_q != 0;
    const bool use_hist = do3 && P.ang_hist, use_qhist = doq && P.q_hist;
    unsigned *s_qhist = s_hist + (use_hist ? P.nbins : 0);
    if (use_hist)
        for (int i = tid; i < P.nbins; i += kB32Threads) s_hist[i] = 0u;
    if (use_qhist)
        for (int i = tid; i < P.q_nbins; i += kB32Threads) s_qhist[i] = 0u;
    if (tid < kB32MaxPairs) {
        int b = 1;  // p = b (b - 1) / 2 + a, a < b
        while ((b + 1) * b / 2 <= tid) ++b;
        S.pair_ab[tid] = (unsigned char)((tid - b * (b - 1) / 2) | (b << 4));
    }
    if (tid == 0) {
        for (int s = 0; s < kBkStages; ++s) {
            mbar_init(&S.bar_full[s], 32);
            mbar_init(&S.bar_raw[s], 1);
            mbar_init(&S.bar_empty[s], kB32Warps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= kB32Warps) {
        bk_producer<B32Smem, kB32Producers>(P, B, S, lane, warp - kB32Warps);
        return;
    }

    const float low3sq = (float)P.low3sq, high3sq = (float)P.high3sq, lowqsq = (float)P.lowqsq;
    const bool last1 = P.wq_max <= 1;
    const float selsq1 = (float)(last1 ? P.highqsq : fmin(P.highqsq, fmin(P.highq, P.rc1) * fmin(P.highq, P.rc1)));
    const float reach2 = fmaxf(do3 ? high3sq : 0.f, doq ? selsq1 : 0.f);
    const float hist_lo = (float)P.hist_lo, hist_hi = (float)P.hist_hi;
    const float inv_width = (float)((double)P.nbins / (P.hist_hi - P.hist_lo));
    const HistSpec qspec = hist_spec(0.0, 1.0, P.q_nbins);
    unsigned short *const my_list = &S.lj[0][tid];

    LaneStats st;
    st.reset();
    int cur_f = -1;
    // the stages are visited round robin: the order a single producer fills them in; with one producer per stage every
    // stage is its own channel and the loop runs until all of them have said "done"
    unsigned live = (1u << kBkStages) - 1u, parity = 0u;
    for (int s = 0; live != 0u; s = (s + 1 == kBkStages) ? 0 : s + 1) {
        if (!((live >> s) & 1u)) continue;
        mbar_wait(&S.bar_full[s], (parity >> s) & 1u);
        parity ^= 1u << s;
        BkItem &I = S.item[s];
        if (I.done) {
            if (kB32Producers == 1) break;
            live &= ~(1u << s);
            continue;
        }
        const int f = I.frame;
        if (f != cur_f) {
            if (cur_f >= 0) {
                bk_flush_stats(P, cur_f, st);
                if ((use_hist || use_qhist) && P.hist_per_frame) {
                    b32_consumer_bar();
                    if (use_hist) b32_flush_bins(s_hist, P.ang_hist + (size_t)cur_f * P.nbins, P.nbins, true, tid);
                    if (use_qhist) b32_flush_bins(s_qhist, P.q_hist + (size_t)cur_f * P.q_nbins, P.q_nbins, true, tid);
                    b32_consumer_bar();
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
            int slot = 0, hx1 = 1, hrow = rstride + 1, gj = 0;
            if (valid) {
                int r = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int t = r + step;
                    if (t < n_crows && I.crow_off[t] <= ci) r = t;
                }
                slot = I.crow_slot[r] + (ci - I.crow_off[r]);
                gj = I.crow_g0[r] + (ci - I.crow_off[r]);  // the centre's place in the cell-sorted arrays = its id in the queues
                hrow = I.crow_hrow[r];
                const unsigned short *row = cst + hrow * kBkCsW;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int t = hx1 + step;
                    if (t <= nbx && (int)row[t] <= slot) hx1 = t;
                }
            }
            const float4 me = loc[slot];
            const size_t out_index = (size_t)f * P.n_pos + __float_as_int(me.w);  // .w of a staged atom: its original index

            // ---------------- phase 1: sweep over the 9 rows of the stencil, final distances ------------------
            int nl = 0;
            if (valid) {
                const unsigned short *row = cst + (hrow - rstride - 1) * kBkCsW + hx1 - 1;
#pragma unroll 1
                for (int r9 = 0; r9 < 9; ++r9) {
                    int j = row[0];
                    const int jend = row[3];
                    row += (r9 == 2 || r9 == 5) ? (rstride - 2) * kBkCsW : kBkCsW;
                    float4 w = loc[j];
#pragma unroll 2
                    while (j < jend) {
                        const float4 wn = loc[j + 1];  // a stage holds one spare entry
                        const float dx = w.x - me.x, dy = w.y - me.y, dz = w.z - me.z;
                        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (r2 <= reach2) {
                            my_list[min(nl, kBkListCap) * kB32Consumers] = (unsigned short)j;
                            ++nl;
                        }
                        w = wn;
                        ++j;
                    }
                }
            }
            bool overflow = nl > kBkListCap;

            // ---------------- phase 2: classification, three-body compaction, register top-4 --------------------
            Top4F top;
            top.reset();
            int K3 = 0, nq = 0;
            if (valid && !overflow) {
#pragma unroll 2
                for (int k = 0; k < nl; ++k) {
                    const int j = my_list[k * kB32Consumers];
                    if (j == slot) continue;  // the centre itself
                    const float4 w = loc[j];
                    const float dx = w.x - me.x, dy = w.y - me.y, dz = w.z - me.z;
                    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (doq && r2 > lowqsq && r2 <= selsq1) {
                        ++nq;
                        top.insert(r2, __float_as_int(w.w), dx, dy, dz);
                    }
                    if (do3 && r2 > low3sq && r2 <= high3sq) {
                        if (K3 < kB32EntCap) S.ent[K3][(tid & ~31) | ((tid + 3 * K3) & 31)] = make_float4(dx, dy, dz, r2);
                        ++K3;
                    }
                }
                if (K3 > kB32EntCap) overflow = true;
            }
            bool q_go = valid && doq && !overflow;
            const bool b3_go = valid && do3 && !overflow;
            if (valid && overflow) {
                const uint32_t at = atomicAdd(P.counters + kCntFallback, 1u);
                P.fb_list[at] = (uint32_t)gj | (do3 ? kFbNeed3b : 0u) | (doq ? kFbNeedQ : 0u);
                atomicAdd(P.counters + kCntOverflow, 1u);
            }
            if (q_go && nq < 4 && !last1) {  // fewer than four inside the radius the stencil guarantees
                bk_push_q(P, (uint32_t)gj);
                q_go = false;
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
                    const int ea = ab & 15, eb = ab >> 4;
                    const float4 va = S.ent[ea][warp * 32 + ((t + 3 * ea) & 31)], vb = S.ent[eb][warp * 32 + ((t + 3 * eb) & 31)];
                    float th;
                    bool binned = true;
                    if (va.w == 0.f || vb.w == 0.f) {
                        th = 0.f;  // coincident positions: CosAngle3 returns 0
                    } else {
                        const float c = cos32(va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w);
                        th = acosf(c) * 57.29577951308232f;
                        if (c == -1.f) binned = false;  // the reference returns -180 here: outside every range starting at 0
                        else if (th >= 100.f && th <= 120.f) {
                            st.tet_count += 1u;
                            st.tet_cos += (double)c;
                            st.tet_cossq += (double)c * (double)c;
                        }
                    }
                    st.n_angles += 1u;
                    if (binned && th >= hist_lo && th <= hist_hi) {
                        const int pos = min((int)((th - hist_lo) * inv_width), P.nbins - 1);
                        if (use_hist) atomicAdd(s_hist + pos, 1u);
                    }
                }
                __syncwarp();
                if (b3_go) {
                    if (P.n3) P.n3[out_index] = K3;
                    st.n_neigh += (unsigned)K3;
                }
            }

            // ---------------- phase 3b: q from the four winners -------------------------------------------------
            if (q_go) finish_q32(P, f, top, min(nq, 4), out_index, st, use_qhist ? s_qhist : nullptr, qspec);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.bar_empty[s]);
    }
    if (cur_f >= 0) bk_flush_stats(P, cur_f, st);
    if (use_hist || use_qhist) {
        b32_consumer_bar();
        if (cur_f >= 0) {
            const size_t rowi = (size_t)(P.hist_per_frame ? cur_f : 0);
            if (use_hist) b32_flush_bins(s_hist, P.ang_hist + rowi * P.nbins, P.nbins, false, tid);
            if (use_qhist) b32_flush_bins(s_qhist, P.q_hist + rowi * P.q_nbins, P.q_nbins, false, tid);
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------

static size_t b32_smem_bytes(const Q3bParams &P) {
    size_t smem = sizeof(B32Smem);
    if (P.do_3b && P.ang_hist) smem += sizeof(unsigned) * P.nbins;
    if (P.do_q && P.q_hist) smem += sizeof(unsigned) * P.q_nbins;
    return smem;
}

bool q3b_brick32_supported(const Q3bParams &P) {
    if (P.centres != nullptr || P.n_valid != nullptr || P.wrapped == nullptr) return false;
    if (P.nc0 < 4 || P.nc1 < 4 || P.nc2 < 4) return false;
    if (b32_smem_bytes(P) > 227u * 1024u) return false;
    const char *env = getenv("WOL_BRICK");  // test switch, as for the fp64 kernel: 0 = never, 1 / 3 = also for small batches
    if (env && env[0] == '0') return false;
    if (env && (env[0] == '1' || env[0] == '3')) return true;
    int nb[3];
    brick_dims(P, nb, kB32Consumers, kB32AtomCap);
    return (long long)nb[0] * nb[1] * nb[2] * P.n_frames >= 2LL * sm_count();
}

int q3b_brick32_launch(const Q3bParams &P, cudaStream_t stream) {
    BrickPlan B = {};
    int nb[3];
    brick_dims(P, nb, kB32Consumers, kB32AtomCap);
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
    const size_t smem = b32_smem_bytes(P);
    cudaError_t e = cudaFuncSetAttribute(q3b_brick32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(brick32)", e);
    long long grid = sm_count();
    if (grid > total) grid = total;
    if (grid > 0) {
        q3b_brick32_kernel<<<(unsigned)grid, kB32Threads, smem, stream>>>(P, B);
        add_launches(1);
        e = cudaGetLastError();
        if (e != cudaSuccess) return set_cuda_error("brick32 kernel launch", e);
    }
    return WOL_OK;
}

}  // namespace wol
