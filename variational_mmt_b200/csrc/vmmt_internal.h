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
bool vmmt_gemm_tc_eligible(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb,
                           int b_kmajor, const float* C, int64_t ldc, int M, int N, int K);
