// tcgen05 flash attention (placeholder dispatch until the kernel lands in this file).
#include "common.cuh"

namespace rb200 {
bool tc_sdpa_supported(const SdpaProblem&) { return false; }
int tc_sdpa(cudaStream_t, const SdpaProblem&) { RB200_FAIL(-5, "tc_sdpa: not built"); }
}  // namespace rb200
