// Whole-GPU cooperative persistent trajectory kernel (placeholder until the first GPU measurement
// of the single-block kernel is in): reports "unsupported" so that DeviceBatch::run uses nb_traj.cu.
#include "nb_internal.h"

namespace nb {
bool grid_traj_supported(int, int, int) { return false; }
size_t grid_traj_workspace_bytes(int, int) { return 0; }
int launch_grid_traj(int, int, int, const TrajDesc*, const double*, int, void*, size_t, cudaStream_t) {
    return NB_ERR_UNSUPPORTED;
}
}  // namespace nb
