// fsv_dpx_variant.cu — one (DUAL, TBM) family of the DPX fill kernel per translation unit.
// Compiled eight times by focalsv_b200/build.py with -DFSV_VARIANT_DUAL=0|1 -DFSV_VARIANT_TBM=0|1|2|3 (in parallel:
// the kernel is large and ptxas time is what the build waits for).
#include <algorithm>
#include <string>

#include "fsv_fill_dpx.cuh"
#include "fsv_fill_ew.cuh"

#ifndef FSV_VARIANT_DUAL
#error "compile with -DFSV_VARIANT_DUAL=0|1 -DFSV_VARIANT_TBM=0|1|2|3"
#endif

#define FSV_CAT3_(a, b, c) a##b##c
#define FSV_CAT3(a, b, c) FSV_CAT3_(a, b, c)

namespace fsv {

int FSV_CAT3(dpx_grid_, FSV_VARIANT_DUAL, FSV_VARIANT_TBM)(int sm_count, int nw, int n_tasks)
{
    return dpx_grid_nw<FSV_VARIANT_DUAL != 0, FSV_VARIANT_TBM>(sm_count, nw, n_tasks);
}
int FSV_CAT3(dpx_launch_, FSV_VARIANT_DUAL, FSV_VARIANT_TBM)(cudaStream_t stream, int nw, int grid, bool excl, const DpxParams& P, std::string* err)
{
    return dpx_launch_nw<FSV_VARIANT_DUAL != 0, FSV_VARIANT_TBM>(stream, nw, grid, excl, P, err);
}

#if FSV_VARIANT_TBM == 1
int FSV_CAT3(dpx_launch_seg_, FSV_VARIANT_DUAL, )(cudaStream_t stream, int sm_count, int nw, int n_segs, const DpxParams& P, std::string* err)
{
    return dpx_launch_seg_nw<FSV_VARIANT_DUAL != 0>(stream, sm_count, nw, n_segs, P, err);
}
// the edge-warp kernel (fsv_fill_ew.cuh): left-aligned traceback only
int FSV_CAT3(ew_grid_, FSV_VARIANT_DUAL, )(int sm_count, int nw, int n_tasks) { return ew_grid_nw<FSV_VARIANT_DUAL != 0>(sm_count, nw, n_tasks); }
int FSV_CAT3(ew_launch_, FSV_VARIANT_DUAL, )(cudaStream_t stream, int nw, int grid, bool excl, const DpxParams& P, std::string* err)
{
    return ew_launch_nw<FSV_VARIANT_DUAL != 0>(stream, nw, grid, excl, P, err);
}
#endif

}  // namespace fsv
