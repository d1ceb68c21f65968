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

// Sorted per-atom record: periodic fixed-point coordinates + original atom index.  16 bytes, loaded
// as one int4.
struct alignas(16) Rec {
    uint32_t x, y, z;
    int32_t idx;
};

struct WorkspaceLayout {
    size_t off_cell_start;  // uint32[F*ncell + 1]  counts during the build, exclusive starts afterwards
    size_t off_block_sums;  // uint32[scan blocks + 1]
    size_t off_cell_id;     // uint32[F*N]
    size_t off_slot;        // uint32[F*N]  rank of the atom inside its cell
    size_t off_recs;        // Rec[F*N]     atoms grouped by (frame, cell)
    size_t off_counters;    // uint32[8]    [0] q-fallback count, [1] three-body overflow count
    size_t off_fb_list;     // uint32[F*M]  centres whose 4-NN search must be widened
    size_t off_ov_list;     // uint32[F*M]  centres with more three-body neighbours than the fast path holds
    size_t total;
    int64_t n_cells_total;
    int64_t n_atoms_total;
    int32_t scan_blocks;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline WorkspaceLayout workspace_layout(int32_t n_frames, int32_t n_pos, int32_t n_centres_max,
                                        const int32_t nc[3]) {
    WorkspaceLayout w;
    int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    w.n_cells_total = ncell * n_frames;
    w.n_atoms_total = (int64_t)n_pos * n_frames;
    int64_t n_centres_total = (int64_t)(n_centres_max > n_pos ? n_centres_max : n_pos) * n_frames;
    w.scan_blocks = (int32_t)((w.n_cells_total + 1 + kScanTile - 1) / kScanTile);
    size_t o = 0;
    w.off_cell_start = o;
    o = align_up(o + (size_t)(w.n_cells_total + 1) * 4, 256);
    w.off_block_sums = o;
    o = align_up(o + (size_t)(w.scan_blocks + 1) * 4, 256);
    w.off_cell_id = o;
    o = align_up(o + (size_t)w.n_atoms_total * 4, 256);
    w.off_slot = o;
    o = align_up(o + (size_t)w.n_atoms_total * 4, 256);
    w.off_recs = o;
    o = align_up(o + (size_t)w.n_atoms_total * sizeof(Rec), 256);
    w.off_counters = o;
    o = align_up(o + 8 * 4, 256);
    w.off_fb_list = o;
    o = align_up(o + (size_t)n_centres_total * 4, 256);
    w.off_ov_list = o;
    o = align_up(o + (size_t)n_centres_total * 4, 256);
    w.total = o;
    return w;
}

}  // namespace wol
