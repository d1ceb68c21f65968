// Layout of the caller-owned scratch buffer.  Everything the kernels share between launches lives
// here; offsets are a pure function of the batch shape so the caller can size it up front
// (wol_workspace_bytes) and reuse it across calls.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace wol {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Sorted per-atom records, grouped by (frame, cell).
// FP64 mode: original coordinates (bit-identical to the input, the reference arithmetic needs them),
// original atom index and the cell coordinates packed 10 bits per axis (cx | cy << 10 | cz << 20).
// 32 bytes = one sector: one 256-bit store / load (two 16-byte loads where a kernel prefers them).
struct alignas(32) RecD {
    double x, y, z;
    int32_t idx;
    int32_t cell;
};
// FP32 mode: 16 bytes = one load.
struct alignas(16) RecF {
    float x, y, z;
    int32_t idx;
};

// Device counters (uint32) at off_counters.
enum {
    kCntFallback = 0,  // entries in the fallback list
    kCntWidened = 1,   // centres whose q search had to be widened
    kCntOverflow = 2,  // centres whose list overflowed the fast path
    kCntFatal = 3,     // centres whose list overflowed the large-capacity path
    kCntLevel2 = 4,    // entries in the second-level list (widened search at half-width 2 was not enough)
    kCntBrick = 5,     // brick path: next brick to hand out
    kCntSlowPair = 6,  // brick path: three-body pairs re-evaluated in exact arithmetic (decision too close to call)
    kCntBrickFb = 7,   // brick path: one-cell-wide sub-bricks too dense to stage (their centres went to the queue)
    kNumCounters = 8
};

// Fallback list entry: centre id (frame * n_centres + centre) in the low 30 bits, flags above.
constexpr uint32_t kFbNeedQ = 1u << 30;
constexpr uint32_t kFbNeed3b = 1u << 31;
constexpr uint32_t kFbIdMask = (1u << 30) - 1u;

struct WorkspaceLayout {
    size_t off_counters;    // uint32[kNumCounters], always at offset 0
    size_t off_cell_start;  // uint32[F*ncell + 1]  counts during the build, exclusive starts afterwards
    size_t off_block_sums;  // uint32[scan blocks + 1]
    size_t off_recs;        // RecD[F*N] (or RecF) atoms grouped by (frame, cell)
    size_t off_wrapped;     // float4[F*N]  box-wrapped float coordinates, same order as recs; .w = the atom's place in
                            // this array (fp64 records) or its original index (fp32 records)
    size_t off_fb_list;     // uint32[2*F*M]  centres the fast path handed on; second half = second level
    size_t total;
    int64_t n_cells_total;
    int64_t n_atoms_total;
    int64_t n_centres_total;
    int32_t scan_blocks;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline WorkspaceLayout workspace_layout(int32_t n_frames, int32_t n_pos, int32_t n_centres_max,
                                        const int32_t nc[3]) {
    WorkspaceLayout w;
    int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    w.n_cells_total = ncell * n_frames;
    w.n_atoms_total = (int64_t)n_pos * n_frames;
    w.n_centres_total = (int64_t)(n_centres_max > n_pos ? n_centres_max : n_pos) * n_frames;
    w.scan_blocks = (int32_t)((w.n_cells_total + 1 + kScanTile - 1) / kScanTile);
    // the counters come first: their address does not move when the batch shape changes, so the sticky overflow flag
    // survives a caller that reuses one workspace for batches of different sizes (e.g. a shorter last batch)
    size_t o = 0;
    w.off_counters = o;
    o = align_up(o + kNumCounters * 4, 256);
    w.off_cell_start = o;
    o = align_up(o + (size_t)(w.n_cells_total + 1) * 4, 256);
    w.off_block_sums = o;
    o = align_up(o + (size_t)(w.scan_blocks + 1) * 4, 256);
    w.off_recs = o;
    o = align_up(o + (size_t)w.n_atoms_total * sizeof(RecD), 256);
    w.off_wrapped = o;
    o = align_up(o + (size_t)w.n_atoms_total * 16, 256);
    w.off_fb_list = o;
    o = align_up(o + (size_t)w.n_centres_total * 8, 256);
    w.total = o;
    return w;
}

}  // namespace wol
