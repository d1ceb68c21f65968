// Host-side internals shared by the translation units of libwol.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wol_capi.h"
#include "wol_workspace.h"

namespace wol {

int set_error(int code, const char *fmt, ...);
int set_cuda_error(const char *what, cudaError_t e);
void add_launches(int n);

int cell_build_launch(const void *pos, int pos_dtype, const double *box, int n_frames, int n_pos, const int32_t nc[3],
                      int precision, void *workspace, const WorkspaceLayout &lay, cudaStream_t stream);

int effective_box(const void *pos, int pos_dtype, int n_frames, int n_pos, const void *centres, int centre_dtype, int n_centres,
                  const double *box_host, double reach, void *scratch_dev, double *box_out_host, cudaStream_t stream);

int q3b_launch(const wol_q3b_args &a, const WorkspaceLayout &lay, cudaStream_t stream);

int sm_count();

// exclusive prefix sum over n uint32 values, in place (three launches); block_sums: n / kScanTile + 2 values
int exclusive_scan_u32(uint32_t *data, size_t n, uint32_t *block_sums, cudaStream_t stream);

// largest clamped cosine whose AngBetween angle (fortran/waterlib.f90:954-965) is >= ang_deg, found by
// bisection with the host libm; *minus_one_passes = whether the -180 it returns for cosine -1 passes
double angle_cos_threshold(double ang_deg, int *minus_one_passes);

}  // namespace wol
