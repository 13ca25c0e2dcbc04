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
                 int act, int accumulate, int flags, cudaStream_t s);
// fused generator epilogues of the tensor-core GEMM (gemm_tc.cu): mode 1 = per-row log-sum-exp partials per 128-column
// tile instead of C (C may be null), mode 2 = C receives the softmax-NLL gradient of the logits, mode 3 = per-row, per-tile
// {max, sum exp} + the tile's top-`topk` logits with their columns (beam search: no [M,V] log-prob matrix at all)
constexpr int VMMT_TOPK_MAX = 8;
struct VmmtGenEpi {
  int mode;
  float* lse_part;          // mode 1: [ceil(N/128)][M][4] {max, sum exp(x - max), best logit, best column (int bits)}
  float* tgt_logit;         // mode 1: [M]
  const int64_t* target;    // [M]
  const float* row_lse;     // mode 2: [M]
  const float* gscale;      // mode 2: device scalar or null
  float scale;              // mode 2
  long long pad;            // mode 2
  int topk = 0;             // mode 3: candidates kept per (row, tile), <= VMMT_TOPK_MAX
  float* tile_lse = nullptr;  // mode 3: [ceil(N/128)][M][2] {max, sum exp(x - max)}
  float* tile_cand = nullptr; // mode 3: [ceil(N/128)][M][topk][2] {logit, column (int bits)}, best first
};
int vmmt_gemm_tc_ex(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor, float* C,
                    int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate,
                    const VmmtGenEpi* epi, int flags, cudaStream_t s);
// optional second operand pair of the tensor-core GEMM: C = act(A B^T + A2 B2^T + bias), all four operands K-major
// ([rows, K] row-major, the nn.Linear layout); the k-blocks of the second pair extend the same TMEM accumulation.
struct VmmtGemmSecond {
  const float* A2; int64_t lda2;      // [M, K2]
  const float* B2; int64_t ldb2;      // [N, K2]
  int K2;
};
int vmmt_gemm_tc_dual(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor, float* C,
                      int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate,
                      const VmmtGenEpi* epi, const VmmtGemmSecond* second, int flags, cudaStream_t s);
bool vmmt_gemm_tc_eligible(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb,
                           int b_kmajor, const float* C, int64_t ldc, int M, int N, int K, int flags = 0);
