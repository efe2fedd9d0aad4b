// tcgen05 / TMEM / TMA multi-tap GEMM (bf16 operands, fp32 accumulate) -- see gemm_tc.cu.
#pragma once
#include "kernels.cuh"

namespace ptts {
void gemm_tc_init();
// fp32-in / fp32-out debug entry used by ptts_debug_linear(path=3); returns < 0 when unsupported.
int gemm_tc_debug(const LinearParams& p, bool bf16_storage, cudaStream_t s);
}  // namespace ptts
