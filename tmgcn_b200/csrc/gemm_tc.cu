// (c) feature GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32.
// Placeholder until the kernel lands: nothing is eligible, everything takes the SIMT path.
#include "common.cuh"

namespace tmgcn {
bool gemm_tc_eligible(int64_t, int, int) { return false; }
int gemm_tc_fwd(const float *, const float *, float *, int64_t, int, int, int, bool, const float *, cudaStream_t) {
    set_error("gemm_tc: not built");
    return 1;
}
}  // namespace tmgcn
