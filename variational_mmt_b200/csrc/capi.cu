// C-ABI plumbing: error reporting, device queries, GEMM dispatch.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "common.cuh"
#include "vmmt_internal.h"

namespace {
thread_local char g_err[512] = "";
std::mutex g_mu;
int g_sms[64];
bool g_sms_init = false;
}

void vmmt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
extern "C" unsigned long long vmmt_launch_count(void) { return g_launches; }

int vmmt_check_launch(const char* what) {
  ++g_launches;
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    vmmt_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return VMMT_OK;
}

int vmmt_num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_sms_init) { memset(g_sms, 0, sizeof(g_sms)); g_sms_init = true; }
  if (g_sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sms[dev] = n;
  }
  return g_sms[dev];
}

extern "C" const char* vmmt_last_error(void) { return g_err; }
extern "C" int vmmt_version(void) { return 100; }

extern "C" int vmmt_gemm(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb,
                         int b_kmajor, float* C, int64_t ldc, int M, int N, int K, const float* bias,
                         int act, int accumulate, int flags, void* stream) {
  VMMT_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative dims");
  VMMT_REQUIRE(A && B && C, "gemm: null operand");
  flags &= ~VMMT_F_BF16;                 // fp32 operands here: bf16 contractions go through vmmt_gemm_bf16 on cast operands
  cudaStream_t s = (cudaStream_t)stream;
  static const bool log_calls = getenv("VMMT_GEMM_LOG") != nullptr;      // debugging: which path each shape takes
  if (log_calls)
    fprintf(stderr, "[vmmt_gemm] M=%d N=%d K=%d lda=%lld(%s) ldb=%lld(%s) ldc=%lld bias=%d act=%d acc=%d -> %s\n", M, N, K,
            (long long)lda, a_kmajor ? "k" : "mn", (long long)ldb, b_kmajor ? "k" : "mn", (long long)ldc, bias != nullptr,
            act, accumulate,
            (!(flags & VMMT_F_EXACT) && vmmt_gemm_tc_eligible(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, flags))
                ? "tcgen05" : "simt");
  if (!(flags & VMMT_F_EXACT) &&
      vmmt_gemm_tc_eligible(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, flags)) {
    return vmmt_gemm_tc(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, bias, act, accumulate, flags, s);
  }
  return vmmt_gemm_simt(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, bias, act, accumulate, s);
}

// C = act(A1 B1^T + A2 B2^T + bias): two K-major operand pairs contracted into ONE accumulator (one launch, one
// epilogue).  Falls back to two vmmt_gemm calls when the tensor-core kernel does not apply.
extern "C" int vmmt_gemm_dual(const float* A1, int64_t lda1, const float* B1, int64_t ldb1, int K1,
                              const float* A2, int64_t lda2, const float* B2, int64_t ldb2, int K2, float* C,
                              int64_t ldc, int M, int N, const float* bias, int act, int flags, void* stream) {
  VMMT_REQUIRE(M >= 0 && N >= 0 && K1 >= 1 && K2 >= 1, "gemm_dual: bad dims");
  VMMT_REQUIRE(A1 && B1 && A2 && B2 && C, "gemm_dual: null operand");
  flags &= ~VMMT_F_BF16;
  cudaStream_t s = (cudaStream_t)stream;
  if (!(flags & VMMT_F_EXACT) && vmmt_gemm_tc_eligible(A1, lda1, 1, B1, ldb1, 1, C, ldc, M, N, K1, flags) &&
      vmmt_gemm_tc_eligible(A2, lda2, 1, B2, ldb2, 1, C, ldc, M, N, K2, flags)) {
    // each pair has its own tensor maps, so a ragged K tail of either is zero-filled by the TMA independently
    VmmtGemmSecond second{A2, lda2, B2, ldb2, K2};
    return vmmt_gemm_tc_dual(A1, lda1, 1, B1, ldb1, 1, C, ldc, M, N, K1, bias, act, 0, nullptr, &second, flags, s);
  }
  int rc = vmmt_gemm(A1, lda1, 1, B1, ldb1, 1, C, ldc, M, N, K1, bias, VMMT_ACT_NONE, 0, flags, stream);
  if (rc) return rc;
  return vmmt_gemm(A2, lda2, 1, B2, ldb2, 1, C, ldc, M, N, K2, nullptr, act, act == VMMT_ACT_NONE ? 1 : 2, flags, stream);
}
