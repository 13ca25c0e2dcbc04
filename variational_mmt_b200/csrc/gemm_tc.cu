// tcgen05 GEMM -- placeholder until the tensor-core kernel lands (next commit).
#include "common.cuh"
#include "vmmt_internal.h"
bool vmmt_gemm_tc_eligible(const float*, int64_t, int, const float*, int64_t, int, const float*, int64_t,
                           int, int, int) { return false; }
int vmmt_gemm_tc(const float*, int64_t, int, const float*, int64_t, int, float*, int64_t, int, int, int,
                 const float*, int, int, cudaStream_t) { return VMMT_EINVAL; }
