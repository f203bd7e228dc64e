// fsv_fill_dpx.cuh — register-resident DPX fill kernel (placeholder until the kernel lands).
#pragma once
#include <string>
#include "fsv_common.cuh"

namespace fsv {

inline bool dpx_supports(const DevScoring&, const DevTask&) { return false; }

inline int dpx_launch(cudaStream_t, int, const DevScoring&, const uint8_t*, const uint8_t*, const DevTask*,
                      const int32_t*, int, int32_t*, fsv_result*, DevAux*, uint8_t*, std::string*)
{
    return FSV_OK;
}

}  // namespace fsv
