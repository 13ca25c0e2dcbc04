// Internal (non-ABI) declarations shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vmmt.h"

int vmmt_gemm_simt(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb,
                   int b_kmajor, float* C, int64_t ldc, int M, int N, int K, const float* bias,
                   int act, int accumulate, cudaStream_t s);

// tcgen05 path; returns VMMT_EINVAL (without setting an error) when the problem does not meet its
// alignment / size constraints, in which case vmmt_gemm() uses the SIMT kernel.
int vmmt_gemm_tc(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb,
                 int b_kmajor, float* C, int64_t ldc, int M, int N, int K, const float* bias,
                 int act, int accumulate, cudaStream_t s);
// fused generator epilogues of the tensor-core GEMM (gemm_tc.cu): mode 1 = per-row log-sum-exp partials per 128-column
// tile instead of C (C may be null), mode 2 = C receives the softmax-NLL gradient of the logits
struct VmmtGenEpi {
  int mode;
  float* lse_part;          // mode 1: [ceil(N/128)][M][4] {max, sum exp(x - max), best logit, best column (int bits)}
  float* tgt_logit;         // mode 1: [M]
  const int64_t* target;    // [M]
  const float* row_lse;     // mode 2: [M]
  const float* gscale;      // mode 2: device scalar or null
  float scale;              // mode 2
  long long pad;            // mode 2
};
int vmmt_gemm_tc_ex(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor, float* C,
                    int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate,
                    const VmmtGenEpi* epi, cudaStream_t s);
bool vmmt_gemm_tc_eligible(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb,
                           int b_kmajor, const float* C, int64_t ldc, int M, int N, int K);
