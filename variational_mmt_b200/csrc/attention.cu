// Luong "general" global attention core: scores, length-masked softmax, context vector.
//
// Reference: onmt/modules/GlobalAttention.py:108-113 (score), 169-184 (mask, softmax, bmm).
// linear_in / linear_out are GEMMs issued by the host wrapper (vmmt_gemm); this file fuses what
// lies between them so that scores and attention weights stay in registers / shuffles:
//   s[t,b,j] = qp[t,b,:] . ctx[j,b,:]   (j >= len[b] -> -inf)
//   a = softmax_j(s)                     -> align [T,B,S]   (returned to the caller: attns["std"])
//   c[t,b,:] = sum_j a[j] ctx[j,b,:]     -> cvec  [T,B,H]
// Everything is time-major, exactly as the decoder holds it, so the reference's two transposing
// copies (GlobalAttention.py:204-205) disappear.  One warp per query (t,b); the context rows of
// one batch element stay L1/L2 resident across the queries of that element.
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

constexpr int ATT_WARPS = 8;
constexpr int SMAX = 128;   // max source length handled in registers (4 scores per lane)

__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_fwd_kernel(const float* __restrict__ qp, const float* __restrict__ ctx,
                const int64_t* __restrict__ lengths, float* __restrict__ align,
                float* __restrict__ cvec, int T, int B, int S, int H, int tsplit) {
  const int b = blockIdx.x / tsplit, part = blockIdx.x % tsplit;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = lengths ? min((int)lengths[b], S) : S;
  const bool vec = (H & 3) == 0;
  for (int t = part * ATT_WARPS + warp; t < T; t += tsplit * ATT_WARPS) {
    const float* q = qp + ((size_t)t * B + b) * H;
    float sc[SMAX / 32];
#pragma unroll
    for (int r = 0; r < SMAX / 32; ++r) sc[r] = -INFINITY;
    // scores: lanes stride over H, one warp reduction per source position
    for (int j = 0; j < len; ++j) {
      const float* c = ctx + ((size_t)j * B + b) * H;
      float p = 0.f;
      if (vec) {
        for (int k = lane * 4; k < H; k += 128) {
          const float4 a = *reinterpret_cast<const float4*>(q + k);
          const float4 d = *reinterpret_cast<const float4*>(c + k);
          p = fmaf(a.x, d.x, p); p = fmaf(a.y, d.y, p); p = fmaf(a.z, d.z, p); p = fmaf(a.w, d.w, p);
        }
      } else {
        for (int k = lane; k < H; k += 32) p = fmaf(q[k], c[k], p);
      }
      p = warp_sum(p);
      if ((j & 31) == lane) {
#pragma unroll
        for (int r = 0; r < SMAX / 32; ++r) if (r == (j >> 5)) sc[r] = p;
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < SMAX / 32; ++r) mx = fmaxf(mx, sc[r]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < SMAX / 32; ++r) { sc[r] = expf(sc[r] - mx); sum += sc[r]; }   // exp(-inf)=0
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float* arow = align + ((size_t)t * B + b) * S;
#pragma unroll
    for (int r = 0; r < SMAX / 32; ++r) {
      sc[r] *= inv;
      const int j = r * 32 + lane;
      if (j < S) arow[j] = sc[r];
    }
    // context: lanes stride over H, weights broadcast by shuffle
    float* crow = cvec + ((size_t)t * B + b) * H;
    for (int k0 = 0; k0 < H; k0 += 128) {
      const int k = k0 + lane * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float acc1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < len; ++j) {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < SMAX / 32; ++r) if (r == (j >> 5)) a = sc[r];
        a = __shfl_sync(0xffffffffu, a, j & 31);
        const float* c = ctx + ((size_t)j * B + b) * H;
        if (vec) {
          if (k < H) {
            const float4 d = *reinterpret_cast<const float4*>(c + k);
            acc.x = fmaf(a, d.x, acc.x); acc.y = fmaf(a, d.y, acc.y);
            acc.z = fmaf(a, d.z, acc.z); acc.w = fmaf(a, d.w, acc.w);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kk = k0 + e * 32 + lane;
            if (kk < H) acc1[e] = fmaf(a, c[kk], acc1[e]);
          }
        }
      }
      if (vec) {
        if (k < H) *reinterpret_cast<float4*>(crow + k) = acc;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kk = k0 + e * 32 + lane;
          if (kk < H) crow[kk] = acc1[e];
        }
      }
    }
  }
}

// Backward, stage 1 (one warp per query): from dc = dL/dcvec
//   da[j] = dc . ctx[j];  ds = a * (da - sum_i a_i da_i);  dqp = sum_j ds[j] ctx[j]
// ds is written to `dscore` [T,B,S] for stage 2.
__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_bwd_query_kernel(const float* __restrict__ dc, const float* __restrict__ ctx,
                      const float* __restrict__ align, const int64_t* __restrict__ lengths,
                      float* __restrict__ dscore, float* __restrict__ dqp, int T, int B, int S,
                      int H, int tsplit) {
  const int b = blockIdx.x / tsplit, part = blockIdx.x % tsplit;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = lengths ? min((int)lengths[b], S) : S;
  for (int t = part * ATT_WARPS + warp; t < T; t += tsplit * ATT_WARPS) {
    const size_t q = (size_t)t * B + b;
    const float* g = dc + q * H;
    float da[SMAX / 32], a[SMAX / 32];
#pragma unroll
    for (int r = 0; r < SMAX / 32; ++r) {
      da[r] = 0.f;
      const int j = r * 32 + lane;
      a[r] = (j < len) ? align[q * S + j] : 0.f;
    }
    for (int j = 0; j < len; ++j) {
      const float* c = ctx + ((size_t)j * B + b) * H;
      float p = 0.f;
      for (int k = lane; k < H; k += 32) p = fmaf(g[k], c[k], p);
      p = warp_sum(p);
      if ((j & 31) == lane) {
#pragma unroll
        for (int r = 0; r < SMAX / 32; ++r) if (r == (j >> 5)) da[r] = p;
      }
    }
    float dot = 0.f;
#pragma unroll
    for (int r = 0; r < SMAX / 32; ++r) dot = fmaf(a[r], da[r], dot);
    dot = warp_sum(dot);
    float ds[SMAX / 32];
#pragma unroll
    for (int r = 0; r < SMAX / 32; ++r) {
      ds[r] = a[r] * (da[r] - dot);
      const int j = r * 32 + lane;
      if (j < S) dscore[q * S + j] = ds[r];
    }
    float* out = dqp + q * H;
    for (int k0 = 0; k0 < H; k0 += 32) {          // all lanes stay in the loop: shuffles need the full warp
      const int k = k0 + lane;
      float acc = 0.f;
      for (int j = 0; j < len; ++j) {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < SMAX / 32; ++r) if (r == (j >> 5)) v = ds[r];
        v = __shfl_sync(0xffffffffu, v, j & 31);
        if (k < H) acc = fmaf(v, ctx[((size_t)j * B + b) * H + k], acc);
      }
      if (k < H) out[k] = acc;
    }
  }
}

// Backward, stage 2: dctx[j,b,:] (+)= sum_t ( a[t,b,j] dc[t,b,:] + ds[t,b,j] qp[t,b,:] )
__global__ void attn_bwd_ctx_kernel(const float* __restrict__ dc, const float* __restrict__ qp,
                                    const float* __restrict__ align, const float* __restrict__ dscore,
                                    float* __restrict__ dctx, int T, int B, int S, int H,
                                    int accumulate) {
  const int j = blockIdx.x / B, b = blockIdx.x % B;
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const size_t q = (size_t)t * B + b;
      acc = fmaf(align[q * S + j], dc[q * H + k], acc);
      acc = fmaf(dscore[q * S + j], qp[q * H + k], acc);
    }
    float* o = dctx + ((size_t)j * B + b) * H + k;
    *o = accumulate ? (*o + acc) : acc;
  }
}

}  // namespace

extern "C" int vmmt_attention_fwd(const float* qp, const float* ctx, const int64_t* lengths,
                                  float* align, float* cvec, int T, int B, int S, int H,
                                  void* stream) {
  VMMT_REQUIRE(S >= 1 && S <= SMAX, "attention_fwd: src_len %d outside [1,%d]", S, SMAX);
  VMMT_REQUIRE(T >= 1 && B >= 1 && H >= 1, "attention_fwd: bad dims");
  int tsplit = ceil_div(T, ATT_WARPS);
  const int want = ceil_div(2 * vmmt_num_sms(), B);
  if (tsplit > want) tsplit = want;
  if (tsplit < 1) tsplit = 1;
  attn_fwd_kernel<<<B * tsplit, ATT_WARPS * 32, 0, (cudaStream_t)stream>>>(
      qp, ctx, lengths, align, cvec, T, B, S, H, tsplit);
  return vmmt_check_launch("attn_fwd_kernel");
}

extern "C" int vmmt_attention_bwd(const float* dcvec, const float* qp, const float* ctx,
                                  const float* align, const int64_t* lengths, float* dscore_ws,
                                  float* dqp, float* dctx, int accumulate_dctx, int T, int B, int S,
                                  int H, void* stream) {
  VMMT_REQUIRE(S >= 1 && S <= SMAX, "attention_bwd: src_len %d outside [1,%d]", S, SMAX);
  int tsplit = ceil_div(T, ATT_WARPS);
  const int want = ceil_div(2 * vmmt_num_sms(), B);
  if (tsplit > want) tsplit = want;
  if (tsplit < 1) tsplit = 1;
  cudaStream_t s = (cudaStream_t)stream;
  attn_bwd_query_kernel<<<B * tsplit, ATT_WARPS * 32, 0, s>>>(dcvec, ctx, align, lengths, dscore_ws,
                                                              dqp, T, B, S, H, tsplit);
  int rc = vmmt_check_launch("attn_bwd_query_kernel");
  if (rc) return rc;
  attn_bwd_ctx_kernel<<<S * B, 128, 0, s>>>(dcvec, qp, align, dscore_ws, dctx, T, B, S, H,
                                            accumulate_dctx);
  return vmmt_check_launch("attn_bwd_ctx_kernel");
}
