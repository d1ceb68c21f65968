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

int q3b_launch(const wol_q3b_args &a, const WorkspaceLayout &lay, cudaStream_t stream);

int sm_count();

}  // namespace wol
