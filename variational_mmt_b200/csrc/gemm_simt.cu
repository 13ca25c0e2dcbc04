// Exact-fp32 SIMT GEMM with fused bias/activation epilogue.
//
// C[M,N] (row-major, ldc) = act( op(A)[M,K] * op(B)[K,N] + bias[N] ) (+ C when `accumulate`)
//   a_kmajor=1: A stored [M,K] (row stride lda);  a_kmajor=0: A stored [K,M]
//   b_kmajor=1: B stored [N,K] (nn.Linear weight); b_kmajor=0: B stored [K,N]
// The three training layouts are forward (1,1), dgrad (1,0) and wgrad (0,0).
//
// This is the precision reference and the path for skinny / oddly-aligned problems; large
// aligned problems are routed to the tcgen05 kernel (gemm_tc.cu) by vmmt_gemm().
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;   // 256 threads, 4x4 outputs each

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case VMMT_ACT_RELU: return fmaxf(v, 0.0f);
    case VMMT_ACT_TANH: return tanhf(v);
    case VMMT_ACT_SOFTPLUS: return softplusf_(v);
    case VMMT_ACT_SIGMOID: return sigmoidf_(v);
    default: return v;
  }
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                 float* __restrict__ C, int64_t ldc, int M, int N, int K,
                 const float* __restrict__ bias, int act, int accumulate, int ksplit_len) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * ksplit_len;
  const int kend = min(K, kbeg + ksplit_len);
  const int ty = tid / 16, tx = tid % 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile (BM x BK) and B tile (BN x BK) into k-major shared memory ----
#pragma unroll
    for (int r = 0; r < (BM * BK) / 256; ++r) {
      const int e = tid + r * 256;
      int m, k;
      if (A_KMAJOR) { m = e / BK; k = e % BK; } else { k = e / BM; m = e % BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.0f;
      if (gm < M && gk < kend) v = A_KMAJOR ? A[(int64_t)gm * lda + gk] : A[(int64_t)gk * lda + gm];
      As[k][m] = v;
    }
#pragma unroll
    for (int r = 0; r < (BN * BK) / 256; ++r) {
      const int e = tid + r * 256;
      int n, k;
      if (B_KMAJOR) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.0f;
      if (gn < N && gk < kend) v = B_KMAJOR ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float* c = C + (int64_t)gm * ldc + gn;
      if (split) {                       // split-K: C was pre-initialised by the host wrapper
        atomicAdd(c, acc[i][j]);
      } else {
        float v = acc[i][j] + (bias ? bias[gn] : 0.0f);
        if (accumulate == 2) v += *c;              // act(C + A*B + bias)
        v = apply_act(v, act);
        *c = (accumulate == 1) ? (*c + v) : v;
      }
    }
  }
}

// C = bias broadcast (or 0) -- pre-pass for split-K without accumulate
__global__ void init_bias_kernel(float* C, int64_t ldc, int M, int N, const float* bias) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  C[(int64_t)m * ldc + n] = bias ? bias[n] : 0.0f;
}

}  // namespace

int vmmt_gemm_simt(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb,
                   int b_kmajor, float* C, int64_t ldc, int M, int N, int K, const float* bias,
                   int act, int accumulate, cudaStream_t s) {
  if (M <= 0 || N <= 0) return VMMT_OK;
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), 1);
  // split-K for skinny problems (e.g. M = batch = 40 against a 2048x2048 weight): only when the
  // epilogue is linear (no activation), so partial sums can be atomically added.
  int splits = 1;
  const int tiles = grid.x * grid.y;
  if (act == VMMT_ACT_NONE && tiles < vmmt_num_sms() / 2 && K >= 512) {
    splits = min(ceil_div(vmmt_num_sms(), tiles), K / 128);
    if (splits < 1) splits = 1;
  }
  int klen = K;
  if (splits > 1) {
    klen = ceil_div(ceil_div(K, splits), BK) * BK;
    splits = ceil_div(K, klen);
    grid.z = splits;
    if (!accumulate) {
      const int64_t tot = (int64_t)M * N;
      init_bias_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(C, ldc, M, N, bias);
      int rc0 = vmmt_check_launch("gemm_init_bias");
      if (rc0) return rc0;
    } else if (bias) {
      vmmt_set_error("vmmt_gemm_simt: bias with accumulate in split-K is unsupported");
      return VMMT_EINVAL;
    }
  }
#define LAUNCH(AK, BKM)                                                                      \
  gemm_simt_kernel<AK, BKM><<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, act, \
                                                  accumulate, klen)
  if (a_kmajor && b_kmajor) LAUNCH(true, true);
  else if (a_kmajor && !b_kmajor) LAUNCH(true, false);
  else if (!a_kmajor && b_kmajor) LAUNCH(false, true);
  else LAUNCH(false, false);
#undef LAUNCH
  return vmmt_check_launch("gemm_simt");
}
