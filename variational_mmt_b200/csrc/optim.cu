// Fused global-norm clip + Adam over the flat parameter / gradient buffers.
//
// Reference: onmt/Optim.py:69-70 (Adam, betas (0.9,0.999), eps 1e-9), :94-96 (clip_grad_norm with
// max_grad_norm 5, then optimizer.step()).  torch semantics restated:
//   total = ||g||_2 over all parameters;  coef = min(1, max_norm / (total + 1e-6));  g <- coef * g
//   m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2
//   p <- p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// All parameters live in one contiguous fp32 buffer (the Python side builds the model that way),
// so the whole update is two launches and the norm never visits the host.
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

__global__ void __launch_bounds__(256)
sqnorm_partial_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  const int64_t n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s = fmaf(g[i], g[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = (double)s;
}

__global__ void sqnorm_final_kernel(const double* __restrict__ partial, int nblk,
                                    float* __restrict__ out, int accumulate) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + (float)sh[0];
}

__global__ void __launch_bounds__(256)
adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, int64_t n, const float* __restrict__ sqnorm, float max_norm,
                 float grad_scale, float step_size, float b1, float b2, float inv_bc2_sqrt,
                 float eps) {
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float total = sqrtf(sqnorm[0]) * grad_scale;
    coef *= fminf(1.0f, max_norm / (total + 1e-6f));
  }
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = n / 4;
  if (i < n4) {
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x = ga[e] * coef;
      ma[e] = b1 * ma[e] + (1.f - b1) * x;
      va[e] = b2 * va[e] + (1.f - b2) * x * x;
      pa[e] -= step_size * ma[e] / (sqrtf(va[e]) * inv_bc2_sqrt + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  } else if (i == n4) {
    for (int64_t k = n4 * 4; k < n; ++k) {
      const float x = g[k] * coef;
      m[k] = b1 * m[k] + (1.f - b1) * x;
      v[k] = b2 * v[k] + (1.f - b2) * x * x;
      p[k] -= step_size * m[k] / (sqrtf(v[k]) * inv_bc2_sqrt + eps);
    }
  }
}

// Zero fill on a FEW resident blocks (grid-stride float4 stores).  A plain memset of the 171 MB gradient buffer floods
// every SM with thousands of short blocks: fine alone, but beside the encoder recurrences it keeps their thread-block
// clusters from being placed, and beside the generator GEMM it takes a third of the HBM bandwidth.  `max_blocks`
// light blocks (one per SM, 256 threads) co-reside with anything and spread the same bytes over a longer window.
__global__ void __launch_bounds__(256)
fill_zero_kernel(float4* __restrict__ p, int64_t n4, float* __restrict__ tail, int ntail) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) p[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}

}  // namespace

extern "C" int vmmt_fill_zero(float* p, int64_t n, int max_blocks, void* stream) {
  if (n <= 0) return VMMT_OK;
  VMMT_REQUIRE(((uintptr_t)p & 15) == 0, "fill_zero: buffer must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int nblk = max_blocks > 0 ? max_blocks : vmmt_num_sms();
  if ((int64_t)nblk * 256 > n4 && n4 > 0) nblk = ceil_div(n4, 256);
  if (nblk < 1) nblk = 1;
  fill_zero_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>((float4*)p, n4, p + n4 * 4, (int)(n - n4 * 4));
  return vmmt_check_launch("fill_zero");
}

namespace {
// up to 5 small device-to-device copies in ONE launch (the step's static input buffers): done by SM threads, not by a copy
// engine -- measured: a cudaMemcpyAsync issued while the previous update's all-gather was still arriving over NVLink sat
// for 125 us (8 GPUs) before it completed, a kernel launched beside the same traffic is not held up
struct CopyList { const char* src[5]; char* dst[5]; long long bytes[5]; int n; };
__global__ void copy_list_kernel(const CopyList L) {
  for (int j = 0; j < L.n; ++j) {
    const long long nb = L.bytes[j];
    const bool vec = ((reinterpret_cast<uintptr_t>(L.src[j]) | reinterpret_cast<uintptr_t>(L.dst[j])) & 15) == 0;
    const long long n16 = vec ? nb / 16 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
      reinterpret_cast<uint4*>(L.dst[j])[i] = reinterpret_cast<const uint4*>(L.src[j])[i];
    for (long long i = n16 * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += (long long)gridDim.x * blockDim.x)
      L.dst[j][i] = L.src[j][i];
  }
}
}  // namespace

extern "C" int vmmt_copy_list(const void* const* src, void* const* dst, const int64_t* bytes, int n, void* stream) {
  VMMT_REQUIRE(n >= 1 && n <= 5 && src && dst && bytes, "copy_list: 1..5 copies");
  CopyList L;
  long long mx = 0;
  for (int j = 0; j < n; ++j) {
    L.src[j] = (const char*)src[j]; L.dst[j] = (char*)dst[j]; L.bytes[j] = bytes[j];
    if (bytes[j] > mx) mx = bytes[j];
  }
  L.n = n;
  int nblk = ceil_div(mx / 16 + 1, 256);
  if (nblk > 64) nblk = 64;
  copy_list_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(L);
  return vmmt_check_launch("copy_list");
}

extern "C" size_t vmmt_sqnorm_workspace_bytes(void) { return 1024 * sizeof(double); }

extern "C" int vmmt_sqnorm(const float* g, int64_t n, float* out, int accumulate, void* workspace,
                           void* stream) {
  VMMT_REQUIRE(((uintptr_t)g & 15) == 0, "sqnorm: buffer must be 16-byte aligned");
  int nblk = ceil_div(ceil_div(n, 4), 256);
  if (nblk > 1024) nblk = 1024;
  if (nblk < 1) nblk = 1;
  sqnorm_partial_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(g, n, (double*)workspace);
  int rc = vmmt_check_launch("sqnorm_partial");
  if (rc) return rc;
  sqnorm_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const double*)workspace, nblk, out, accumulate);
  return vmmt_check_launch("sqnorm_final");
}

extern "C" int vmmt_adam_clip_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                   int64_t n, const float* sqnorm, float max_norm, float grad_scale,
                                   float lr, float beta1, float beta2, float eps, int64_t step,
                                   void* stream) {
  VMMT_REQUIRE(step >= 1, "adam_clip_step: step must be >= 1");
  VMMT_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
               "adam_clip_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const int64_t n4 = n / 4;
  adam_clip_kernel<<<ceil_div(n4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, sqnorm, max_norm, grad_scale, step_size, beta1, beta2,
      inv_bc2_sqrt, eps);
  return vmmt_check_launch("adam_clip");
}
