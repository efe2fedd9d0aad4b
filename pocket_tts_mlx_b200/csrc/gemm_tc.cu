#include "gemm_tc.cuh"
namespace ptts {
void gemm_tc_init() {}
int gemm_tc_debug(const LinearParams&, bool, cudaStream_t) { return -1; }
}  // namespace ptts
