// Generator = Linear(H,V) + LogSoftmax fused with the NLL criterion and the accuracy statistics.
//
// Reference: onmt/ModelConstructor.py:582-585 (generator), onmt/VILoss.py:228,243 (scores, NLLLoss
// with weight[pad]=0, size_average=False), onmt/VILoss.py:515-531 (_stats: argmax accuracy over
// non-pad targets).
//
// forward : logits tile -> per-row log-sum-exp, target log-prob, argmax  => {nll_sum, n_words,
//           n_correct}; log-probabilities are never written out.
// backward: dlogits = (softmax - onehot) * [tgt != pad] * scale, then dX = dlogits W,
//           dW += dlogits^T X, db += colsum(dlogits).  The logits are recomputed from (x, W, lse),
//           nothing of size M x V is saved between forward and backward.
// decode  : vmmt_generator_logprobs materialises log-probs [M,V] (beam search needs them).
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

// one CTA per row of logits [M,V]
__global__ void __launch_bounds__(256)
row_lse_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t pad,
               float* __restrict__ lse, float* __restrict__ rowstat, int V) {
  __shared__ float smax[8];
  __shared__ int sarg[8];
  __shared__ float red[32];
  const int row = blockIdx.x;
  const float* l = logits + (size_t)row * V;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    const float v = l[j];
    if (v > mx) { mx = v; arg = j; }          // strided scan keeps the smallest index per thread
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { smax[w] = mx; sarg[w] = arg; }
  __syncthreads();
  mx = smax[0]; arg = sarg[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i)
    if (smax[i] > mx || (smax[i] == mx && sarg[i] < arg)) { mx = smax[i]; arg = sarg[i]; }
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += expf(l[j] - mx);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float L = mx + logf(s);
    lse[row] = L;
    if (rowstat) {
      const int64_t tg = target[row];
      const bool on = tg != pad;
      rowstat[row * 3 + 0] = on ? (L - l[tg]) : 0.f;      // -log p(target)
      rowstat[row * 3 + 1] = on ? 1.f : 0.f;
      rowstat[row * 3 + 2] = (on && arg == (int)tg) ? 1.f : 0.f;
    }
  }
}

// fused path: combine the per-tile partials {max, sumexp, best, besti} of a row -> lse, row statistics
__global__ void __launch_bounds__(128)
lse_combine_kernel(const float4* __restrict__ part, const float* __restrict__ tgt_logit,
                   const int64_t* __restrict__ target, int64_t pad, int M, int ntile,
                   float* __restrict__ lse, float* __restrict__ rowstat) {
  // one warp per row: lane l folds tiles l, l+32, ... (online rescaling), then a fixed-order shuffle tree
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float mx = -INFINITY, s = 0.f, bestv = -INFINITY;
  int besti = 0x7fffffff;
  for (int t = lane; t < ntile; t += 32) {
    const float4 p = __ldg(part + (size_t)t * M + row);
    const float nm = fmaxf(mx, p.x);
    s = s * __expf(mx - nm) + p.y * __expf(p.x - nm);       // mx = -inf, s = 0 on the first tile: 0 * exp(-inf) = 0
    mx = nm;
    const int bi = __float_as_int(p.w);
    if (p.z > bestv || (p.z == bestv && bi < besti)) { bestv = p.z; besti = bi; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float obv = __shfl_xor_sync(0xffffffffu, bestv, o);
    const int obi = __shfl_xor_sync(0xffffffffu, besti, o);
    const float nm = fmaxf(mx, omx);
    const float a = (mx == -INFINITY) ? 0.f : s * __expf(mx - nm);
    const float b = (omx == -INFINITY) ? 0.f : os * __expf(omx - nm);
    s = a + b;
    mx = nm;
    if (obv > bestv || (obv == bestv && obi < besti)) { bestv = obv; besti = obi; }
  }
  if (lane == 0) {
    const float Lr = mx + logf(s);
    lse[row] = Lr;
    const int64_t tg = target[row];
    const bool on = tg != pad;
    rowstat[row * 3 + 0] = on ? (Lr - tgt_logit[row]) : 0.f;
    rowstat[row * 3 + 1] = on ? 1.f : 0.f;
    rowstat[row * 3 + 2] = (on && besti == (int)tg) ? 1.f : 0.f;
  }
}

// deterministic single-block reduction of the per-row statistics
__global__ void reduce_rowstat_kernel(const float* __restrict__ rowstat, int M, float* __restrict__ out) {
  __shared__ float red[32];
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    a += rowstat[i * 3]; b += rowstat[i * 3 + 1]; c += rowstat[i * 3 + 2];
  }
  a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red);
  if (threadIdx.x == 0) { out[0] = a; out[1] = b; out[2] = c; }
}

__global__ void dlogits_kernel(float* __restrict__ logits, const float* __restrict__ lse,
                               const int64_t* __restrict__ target, int64_t pad,
                               const float* __restrict__ gscale, float scale, int V) {
  if (gscale) scale *= gscale[0];
  const int row = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= V) return;
  const int64_t tg = target[row];
  float* p = logits + (size_t)row * V + j;
  float g = 0.f;
  if (tg != pad) g = (expf(*p - lse[row]) - (j == (int)tg ? 1.f : 0.f)) * scale;
  *p = g;
}

__global__ void logprob_kernel(float* __restrict__ logits, const float* __restrict__ lse, int V) {
  const int row = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < V) logits[(size_t)row * V + j] -= lse[row];
}

}  // namespace

extern "C" int vmmt_gemm(const float*, int64_t, int, const float*, int64_t, int, float*, int64_t,
                         int, int, int, const float*, int, int, int, void*);

extern "C" int vmmt_generator_nll_wgrad(const float* x, const void* workspace, int M, int H, int V, float* dW,
                                        float* db, int flags, void* stream);

extern "C" int vmmt_cast_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int rows, int cols, void* stream);

// [M,V] fp32 logits / dlogits | [M,4] row statistics | bf16 copies of x [M,Hp] and W [V,Hp] (VMMT_F_BF16), Hp = H rounded to 8
static size_t gen_f32_floats(int M, int V) { return (((size_t)M * V + (size_t)M * 4) + 3) / 4 * 4; }
extern "C" size_t vmmt_generator_workspace_bytes(int M, int H, int V) {
  const size_t hp = (size_t)(H + 7) / 8 * 8;
  return gen_f32_floats(M, V) * sizeof(float) + ((size_t)M + (size_t)V) * hp * 2 + 32;
}
// bf16 variant: x and W are cast into the tail of the workspace; returns the operand pointers / pitch to contract
static int gen_bf16_operands(const float* x, const float* W, int M, int H, int V, void* workspace, const float** xo,
                             const float** wo, int64_t* ld, void* stream) {
  const int64_t hp = (int64_t)(H + 7) / 8 * 8;
  char* base = reinterpret_cast<char*>(workspace) + gen_f32_floats(M, V) * sizeof(float);
  base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(base) + 15) & ~(uintptr_t)15);
  void* xb = base;
  void* wb = base + (size_t)M * hp * 2;
  int rc = vmmt_cast_bf16(x, H, xb, hp, M, H, stream);
  if (rc) return rc;
  rc = vmmt_cast_bf16(W, H, wb, hp, V, H, stream);
  if (rc) return rc;
  *xo = reinterpret_cast<const float*>(xb); *wo = reinterpret_cast<const float*>(wb); *ld = hp;
  return VMMT_OK;
}

extern "C" int vmmt_generator_nll_fwd(const float* x, const float* W, const float* b,
                                      const int64_t* target, int64_t pad_idx, int M, int H, int V,
                                      float* lse, float* stats3, void* workspace,
                                      size_t workspace_bytes, int flags, void* stream) {
  if (workspace_bytes < vmmt_generator_workspace_bytes(M, H, V)) {
    vmmt_set_error("generator_nll_fwd: workspace too small");
    return VMMT_EWORKSPACE;
  }
  float* logits = (float*)workspace;
  float* rowstat = logits + (size_t)M * V;
  int rc;
  // tensor-core mode: log-sum-exp / target logit / argmax are folded into the GEMM epilogue; the logits of a tile
  // live in tensor memory and registers only.  Partials: [ceil(V/128)][M] float4 at the head of the workspace.
  if (!(flags & VMMT_F_EXACT) && !getenv("VMMT_GEN_UNFUSED") &&
      vmmt_gemm_tc_eligible(x, H, 1, W, H, 1, nullptr, V, M, V, H, flags)) {
    const int ntile = ceil_div(V, 128);
    float* part = logits;                                  // ntile * M * 4 floats  <<  M * V
    float* tgt_logit = part + (size_t)ntile * M * 4;
    VmmtGenEpi epi{1, part, tgt_logit, target, nullptr, nullptr, 1.0f, (long long)pad_idx};
    const float *xo = x, *wo = W;
    int64_t ldo = H;
    if (flags & VMMT_F_BF16) {                             // bf16 operands: cast into the tail of the workspace
      rc = gen_bf16_operands(x, W, M, H, V, workspace, &xo, &wo, &ldo, stream);
      if (rc) return rc;
    }
    rc = vmmt_gemm_tc_ex(xo, ldo, 1, wo, ldo, 1, nullptr, V, M, V, H, b, VMMT_ACT_NONE, 0, &epi, flags, (cudaStream_t)stream);
    if (rc) return rc;
    lse_combine_kernel<<<ceil_div(M, 4), 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(part), tgt_logit, target, pad_idx, M, ntile, lse, rowstat);
    rc = vmmt_check_launch("lse_combine");
    if (rc) return rc;
    reduce_rowstat_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rowstat, M, stats3);
    return vmmt_check_launch("reduce_rowstat");
  }
  rc = vmmt_gemm(x, H, 1, W, H, 1, logits, V, M, V, H, b, VMMT_ACT_NONE, 0, flags & ~VMMT_F_BF16, stream);
  if (rc) return rc;
  row_lse_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(logits, target, pad_idx, lse, rowstat, V);
  rc = vmmt_check_launch("row_lse");
  if (rc) return rc;
  reduce_rowstat_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rowstat, M, stats3);
  return vmmt_check_launch("reduce_rowstat");
}

extern "C" int vmmt_generator_nll_bwd(const float* x, const float* W, const float* b,
                                      const int64_t* target, int64_t pad_idx, const float* lse,
                                      const float* gscale, float scale, int M, int H, int V, float* dx, float* dW,
                                      float* db, void* workspace, size_t workspace_bytes, int flags,
                                      void* stream) {
  if (workspace_bytes < vmmt_generator_workspace_bytes(M, H, V)) {
    vmmt_set_error("generator_nll_bwd: workspace too small");
    return VMMT_EWORKSPACE;
  }
  float* dl = (float*)workspace;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = VMMT_EINVAL;
  // tensor-core mode: the recomputed logits are turned into dlogits in the GEMM epilogue (one pass over M x V)
  if (!(flags & VMMT_F_EXACT) && !getenv("VMMT_GEN_UNFUSED") && (V & 3) == 0 &&
      vmmt_gemm_tc_eligible(x, H, 1, W, H, 1, dl, V, M, V, H, flags)) {
    VmmtGenEpi epi{2, nullptr, nullptr, target, lse, gscale, scale, (long long)pad_idx};
    const float *xo = x, *wo = W;
    int64_t ldo = H;
    rc = VMMT_OK;
    if (flags & VMMT_F_BF16) rc = gen_bf16_operands(x, W, M, H, V, workspace, &xo, &wo, &ldo, stream);   // the recompute matches the forward
    if (rc == VMMT_OK) rc = vmmt_gemm_tc_ex(xo, ldo, 1, wo, ldo, 1, dl, V, M, V, H, b, VMMT_ACT_NONE, 0, &epi, flags, s);
  }
  // the two gradient products below contract the fp32 softmax gradient: they stay on the TF32 path in the bf16 variant
  flags &= ~VMMT_F_BF16;
  if (rc != VMMT_OK) {
    rc = vmmt_gemm(x, H, 1, W, H, 1, dl, V, M, V, H, b, VMMT_ACT_NONE, 0, flags, stream);
    if (rc) return rc;
    dlogits_kernel<<<dim3(ceil_div(V, 256), M), 256, 0, s>>>(dl, lse, target, pad_idx, gscale, scale, V);
    rc = vmmt_check_launch("dlogits");
    if (rc) return rc;
  }
  if (dx) {   // dX[M,H] = dl[M,V] W[V,H]
    rc = vmmt_gemm(dl, V, 1, W, H, 0, dx, H, M, H, V, nullptr, VMMT_ACT_NONE, 0, flags, stream);
    if (rc) return rc;
  }
  if (dW == nullptr && db == nullptr) return VMMT_OK;      // weight gradients taken later by vmmt_generator_nll_wgrad
  return vmmt_generator_nll_wgrad(x, workspace, M, H, V, dW, db, flags, stream);
}

// dW[V,H] += dl^T[V,M] x[M,H];  db += colsum(dl), with dl = the dlogits vmmt_generator_nll_bwd left in `workspace`
// (lets the caller put the weight gradient on another stream, off the critical path of the backward pass)
extern "C" int vmmt_generator_nll_wgrad(const float* x, const void* workspace, int M, int H, int V, float* dW,
                                        float* db, int flags, void* stream) {
  const float* dl = (const float*)workspace;
  if (dW) {
    int rc = vmmt_gemm(dl, V, 0, x, H, 0, dW, H, V, H, M, nullptr, VMMT_ACT_NONE, 1, flags, stream);
    if (rc) return rc;
  }
  if (db) return vmmt_colsum_acc(dl, V, M, V, db, nullptr, stream);
  return VMMT_OK;
}

extern "C" int vmmt_generator_logprobs(const float* x, const float* W, const float* b, int M, int H,
                                       int V, float* out, float* lse_ws, int flags, void* stream) {
  int rc = vmmt_gemm(x, H, 1, W, H, 1, out, V, M, V, H, b, VMMT_ACT_NONE, 0, flags, stream);
  if (rc) return rc;
  row_lse_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(out, nullptr, 0, lse_ws, nullptr, V);
  rc = vmmt_check_launch("row_lse");
  if (rc) return rc;
  logprob_kernel<<<dim3(ceil_div(V, 256), M), 256, 0, (cudaStream_t)stream>>>(out, lse_ws, V);
  return vmmt_check_launch("logprob");
}

// ---- beam-search generator: no [M,V] log-prob matrix.  Workspace = [ntile][M] float2 {max, sum exp} followed by
// [ntile][M][K] float2 {logit, column}; consumed by vmmt_beam_advance_topk (decode.cu).
extern "C" size_t vmmt_generator_topk_workspace_bytes(int M, int V, int K) {
  const size_t ntile = (size_t)ceil_div(V, 128);
  return ntile * (size_t)M * (size_t)(1 + K) * sizeof(float2);
}

extern "C" int vmmt_generator_topk_supported(const float* x, const float* W, int M, int H, int V, int flags) {
  return (!(flags & VMMT_F_EXACT) && vmmt_gemm_tc_eligible(x, H, 1, W, H, 1, nullptr, V, M, V, H, flags)) ? 1 : 0;
}

extern "C" int vmmt_generator_topk(const float* x, const float* W, const float* b, int M, int H, int V, int K,
                                   void* workspace, size_t workspace_bytes, int flags, void* stream) {
  VMMT_REQUIRE(K >= 1 && K <= VMMT_TOPK_MAX, "generator_topk: K = %d outside [1,%d]", K, VMMT_TOPK_MAX);
  if (workspace_bytes < vmmt_generator_topk_workspace_bytes(M, V, K)) {
    vmmt_set_error("generator_topk: workspace too small");
    return VMMT_EWORKSPACE;
  }
  VMMT_REQUIRE(vmmt_generator_topk_supported(x, W, M, H, V, flags),
               "generator_topk: needs the tensor-core GEMM (no VMMT_F_EXACT, 16-byte aligned operands, H %% 4 == 0, V,H >= 64)");
  const size_t ntile = (size_t)ceil_div(V, 128);
  float* tile_lse = (float*)workspace;
  float* tile_cand = tile_lse + ntile * (size_t)M * 2;
  VmmtGenEpi epi{3, nullptr, nullptr, nullptr, nullptr, nullptr, 1.0f, 0};
  epi.topk = K;
  epi.tile_lse = tile_lse;
  epi.tile_cand = tile_cand;
  return vmmt_gemm_tc_ex(x, H, 1, W, H, 1, nullptr, V, M, V, H, b, VMMT_ACT_NONE, 0, &epi, flags & ~VMMT_F_BF16,
                         (cudaStream_t)stream);
}
