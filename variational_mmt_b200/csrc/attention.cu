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
#include <stdlib.h>
#include <mutex>
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


// ---------------------------------------------------------------------------------------------------------------
// Tiled kernels (used whenever T > 1): one CTA per (batch element, tile of 16 queries).  The two contractions of the
// attention core are tiny batched GEMMs ([16,H] x [H,S] and [16,S] x [S,H]); the one-warp-per-query kernels above
// re-read the whole context of a batch element once per query and are latency bound.  Here the context chunk is
// staged through shared memory once per 16 queries, scores / weights never leave registers + shared memory, and
// every global access is a coalesced 128-byte row segment.
constexpr int TT = 16;             // queries per CTA
constexpr int KC = 32;             // contraction chunk staged in shared memory
constexpr int TILE_THREADS = 256;

// P[tt][j] = sum_k X[(t0+tt), b, k] * ctx[j, b, k];  thread (ty = warp, tx = lane) holds tt = 2 ty + {0,1}, j = tx + 32 jj.
// The next k-chunk is fetched into registers while the current one is being contracted (one global-latency
// exposure per kernel instead of one per chunk).
__device__ __forceinline__ void qk_tile(const float* __restrict__ X, const float* __restrict__ ctx, int t0, int T, int b,
                                        int B, int S, int H, float (*Xs)[KC + 1], float (*Cs)[KC + 1], float (&acc)[2][4]) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  constexpr int XP = TT * KC / TILE_THREADS;           // 2 staged X elements per thread
  constexpr int CP = SMAX * KC / TILE_THREADS;         // up to 16 staged context elements per thread
  const int srows = ((S + 31) / 32) * 32;
  const int cp_used = srows * KC / TILE_THREADS;       // 4 per 32 source positions
  float xr[XP], cr[CP];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < XP; ++i) {
      const int e = tid + i * TILE_THREADS, r = e / KC, k = k0 + (e % KC), t = t0 + r;
      xr[i] = (t < T && k < H) ? __ldg(X + ((size_t)t * B + b) * H + k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < CP; ++i) {
      if (i < cp_used) {
        const int e = tid + i * TILE_THREADS, j = e / KC, k = k0 + (e % KC);
        cr[i] = (j < S && k < H) ? __ldg(ctx + ((size_t)j * B + b) * H + k) : 0.f;
      }
    }
  };
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
  fetch(0);
  for (int k0 = 0; k0 < H; k0 += KC) {
    __syncthreads();                                   // previous chunk fully consumed
#pragma unroll
    for (int i = 0; i < XP; ++i) { const int e = tid + i * TILE_THREADS; Xs[e / KC][e % KC] = xr[i]; }
#pragma unroll
    for (int i = 0; i < CP; ++i)
      if (i < cp_used) { const int e = tid + i * TILE_THREADS; Cs[e / KC][e % KC] = cr[i]; }
    __syncthreads();
    if (k0 + KC < H) fetch(k0 + KC);                   // in flight during the contraction below
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float x0 = Xs[2 * ty][k], x1 = Xs[2 * ty + 1][k];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        if (32 * jj < S) {                             // uniform: skip the 32-position groups beyond the source length
          const float c = Cs[tx + 32 * jj][k];
          acc[0][jj] = fmaf(x0, c, acc[0][jj]);
          acc[1][jj] = fmaf(x1, c, acc[1][jj]);
        }
      }
    }
  }
}

// O[(t0+tt), b, k] = sum_{j<len} W[j][tt] * ctx[j, b, k];  W is the [SMAX][TT] shared-memory tile (tt fastest)
__device__ __forceinline__ void wv_tile(const float (*Ws)[TT], const float* __restrict__ ctx, float* __restrict__ O,
                                        int t0, int T, int b, int B, int len, int H) {
  for (int k = threadIdx.x; k < H; k += TILE_THREADS) {
    float acc[TT];
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[i] = 0.f;
    for (int j0 = 0; j0 < len; j0 += 8) {
      float c8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)                        // 8 independent loads in flight; weights beyond len are 0
        c8[u] = (j0 + u < len) ? __ldg(ctx + ((size_t)(j0 + u) * B + b) * H + k) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j0 + u < len) {
          const float c = c8[u];
          const float4* w4 = reinterpret_cast<const float4*>(Ws[j0 + u]);
#pragma unroll
          for (int q = 0; q < TT / 4; ++q) {
            const float4 w = w4[q];
            acc[4 * q] = fmaf(w.x, c, acc[4 * q]); acc[4 * q + 1] = fmaf(w.y, c, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w.z, c, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(w.w, c, acc[4 * q + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < TT; ++i)
      if (t0 + i < T) O[((size_t)(t0 + i) * B + b) * H + k] = acc[i];
  }
}

__global__ void __launch_bounds__(TILE_THREADS)
attn_fwd_tiled_kernel(const float* __restrict__ qp, const float* __restrict__ ctx, const int64_t* __restrict__ lengths,
                      float* __restrict__ align, float* __restrict__ cvec, int T, int B, int S, int H, int ntile) {
  // dynamic shared memory sized by the source length (rounded up to 32 positions): 8 KB at S <= 32 instead of the 27 KB
  // of the S = 128 worst case, so that these CTAs still fit beside a 208 KB GEMM CTA on the same SM
  extern __shared__ __align__(16) float att_smem[];
  const int srows = ((S + 31) / 32) * 32;
  float (*Xs)[KC + 1] = reinterpret_cast<float (*)[KC + 1]>(att_smem);
  float (*Cs)[KC + 1] = reinterpret_cast<float (*)[KC + 1]>(att_smem + TT * (KC + 1));
  float (*Ws)[TT] = reinterpret_cast<float (*)[TT]>(att_smem + TT * (KC + 1) + srows * (KC + 1));
  const int b = blockIdx.x / ntile, t0 = (blockIdx.x % ntile) * TT;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int len = lengths ? min((int)lengths[b], S) : S;
  float acc[2][4];
  qk_tile(qp, ctx, t0, T, b, B, S, H, Xs, Cs, acc);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int tt = 2 * ty + i, t = t0 + tt;
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      if (tx + 32 * jj >= len) acc[i][jj] = -INFINITY;
      mx = fmaxf(mx, acc[i][jj]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) { acc[i][jj] = expf(acc[i][jj] - mx); sum += acc[i][jj]; }   // exp(-inf) = 0
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = tx + 32 * jj;
      const float a = acc[i][jj] * inv;
      if (j < srows) Ws[j][tt] = a;
      if (t < T && j < S) align[((size_t)t * B + b) * S + j] = a;
    }
  }
  __syncthreads();
  wv_tile(Ws, ctx, cvec, t0, T, b, B, len, H);
}

// backward, stage 1 per (b, query tile): da = dc ctx^T;  ds = a (da - sum a da) -> dscore;  dqp = ds ctx
__global__ void __launch_bounds__(TILE_THREADS)
attn_bwd_query_tiled_kernel(const float* __restrict__ dc, const float* __restrict__ ctx, const float* __restrict__ align,
                            const int64_t* __restrict__ lengths, float* __restrict__ dscore, float* __restrict__ dqp,
                            int T, int B, int S, int H, int ntile) {
  // dynamic shared memory sized by the source length (rounded up to 32 positions): 8 KB at S <= 32 instead of the 27 KB
  // of the S = 128 worst case, so that these CTAs still fit beside a 208 KB GEMM CTA on the same SM
  extern __shared__ __align__(16) float att_smem[];
  const int srows = ((S + 31) / 32) * 32;
  float (*Xs)[KC + 1] = reinterpret_cast<float (*)[KC + 1]>(att_smem);
  float (*Cs)[KC + 1] = reinterpret_cast<float (*)[KC + 1]>(att_smem + TT * (KC + 1));
  float (*Ws)[TT] = reinterpret_cast<float (*)[TT]>(att_smem + TT * (KC + 1) + srows * (KC + 1));
  const int b = blockIdx.x / ntile, t0 = (blockIdx.x % ntile) * TT;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int len = lengths ? min((int)lengths[b], S) : S;
  float da[2][4];
  qk_tile(dc, ctx, t0, T, b, B, S, H, Xs, Cs, da);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int tt = 2 * ty + i, t = t0 + tt;
    float a[4], dot = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = tx + 32 * jj;
      a[jj] = (t < T && j < len) ? align[((size_t)t * B + b) * S + j] : 0.f;
      dot = fmaf(a[jj], da[i][jj], dot);
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = tx + 32 * jj;
      const float ds = a[jj] * (da[i][jj] - dot);
      if (j < srows) Ws[j][tt] = ds;
      if (t < T && j < S) dscore[((size_t)t * B + b) * S + j] = ds;
    }
  }
  __syncthreads();
  wv_tile(Ws, ctx, dqp, t0, T, b, B, len, H);
}

// backward, stage 2 per (b, 64 hidden units): dctx[j,b,k] (+)= sum_t ( a[t,b,j] dc[t,b,k] + ds[t,b,j] qp[t,b,k] )
// thread (k = k0 + tid % 64, g = tid / 64): the four 64-thread groups split the work as (position group of 32) x (share of
// the queries): S <= 32 -> one position group, each thread group takes every 4th query and the partial sums are added in a
// fixed order through shared memory; S <= 64 -> 2 x 2; longer sources -> 4 position groups x all queries.  (With the
// position groups alone, three quarters of the CTA idled at the S = 30 of a Multi30k batch.)
__global__ void __launch_bounds__(TILE_THREADS)
attn_bwd_ctx_tiled_kernel(const float* __restrict__ dc, const float* __restrict__ qp, const float* __restrict__ align,
                          const float* __restrict__ dscore, float* __restrict__ dctx, int T, int B, int S, int H,
                          int accumulate, int nk) {
  __shared__ __align__(16) float As[TT][SMAX];
  __shared__ __align__(16) float Ds[TT][SMAX];
  constexpr int JB = SMAX / 4;                       // 32 positions per position group (covers S <= 128)
  __shared__ float Red[3][JB][64];                   // partial sums of the query shares 1..3
  const int b = blockIdx.x / nk, kk = threadIdx.x & 63, k = (blockIdx.x % nk) * 64 + kk;
  const int g = threadIdx.x >> 6;
  const int pg = (S + JB - 1) / JB;                  // position groups in use: 1..4
  const int ts = (pg == 1) ? 4 : (pg == 2 ? 2 : 1);  // query shares
  const int jg = (ts == 1) ? g : g % pg, tsi = (ts == 1) ? 0 : g / pg;
  float acc[JB];
#pragma unroll
  for (int i = 0; i < JB; ++i) acc[i] = 0.f;
  const int jb_used = min(JB, max(0, S - jg * JB));  // positions of this group that exist
  for (int t0 = 0; t0 < T; t0 += TT) {
    __syncthreads();
    for (int e = threadIdx.x; e < TT * SMAX; e += TILE_THREADS) {
      const int tt = e / SMAX, j = e % SMAX, t = t0 + tt;
      const bool ok = t < T && j < S;
      As[tt][j] = ok ? align[((size_t)t * B + b) * S + j] : 0.f;
      Ds[tt][j] = ok ? dscore[((size_t)t * B + b) * S + j] : 0.f;
    }
    __syncthreads();
    const int tn = min(TT, T - t0);
    for (int tt = tsi; tt < tn; tt += ts) {
      const size_t row = ((size_t)(t0 + tt) * B + b) * H;
      const float dcv = k < H ? __ldg(dc + row + k) : 0.f, qv = k < H ? __ldg(qp + row + k) : 0.f;
      const float4* a4 = reinterpret_cast<const float4*>(&As[tt][jg * JB]);
      const float4* d4 = reinterpret_cast<const float4*>(&Ds[tt][jg * JB]);
#pragma unroll
      for (int q = 0; q < JB / 4; ++q) {
        if (4 * q < jb_used) {                       // warp-uniform
          const float4 a = a4[q], d = d4[q];
          acc[4 * q] = fmaf(a.x, dcv, fmaf(d.x, qv, acc[4 * q]));
          acc[4 * q + 1] = fmaf(a.y, dcv, fmaf(d.y, qv, acc[4 * q + 1]));
          acc[4 * q + 2] = fmaf(a.z, dcv, fmaf(d.z, qv, acc[4 * q + 2]));
          acc[4 * q + 3] = fmaf(a.w, dcv, fmaf(d.w, qv, acc[4 * q + 3]));
        }
      }
    }
  }
  if (ts > 1) {                                      // block-uniform
    if (tsi > 0) {
#pragma unroll
      for (int i = 0; i < JB; ++i) Red[(tsi - 1) * pg + jg][i][kk] = acc[i];
    }
    __syncthreads();
    if (tsi > 0) return;
    for (int x = 0; x < ts - 1; ++x) {
#pragma unroll
      for (int i = 0; i < JB; ++i) acc[i] += Red[x * pg + jg][i][kk];
    }
  }
  if (k < H) {
#pragma unroll
    for (int i = 0; i < JB; ++i) {
      const int j = jg * JB + i;
      if (j < S) {
        float* o = dctx + ((size_t)j * B + b) * H + k;
        *o = accumulate ? (*o + acc[i]) : acc[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// v2 of the (sentence, 16-query tile) kernels -- forward (MODE 0: P = qp ctx^T -> softmax -> c = a ctx) and the query side
// of the backward (MODE 1: P = dc ctx^T -> ds = a (P - sum a P) -> dqp = ds ctx).  The v1 kernels above staged 32-wide
// contraction chunks through registers and issued 3 LDS per 2 FMA: 12 K instructions per warp, issue-bound at 31 us for
// 0.07 GFLOP (ncu, cfg1 shape).  Here
//   * the whole 16 x KP query tile and 32 x KP context block (KP <= 512 floats of the contraction) are brought into shared
//     memory with ONE burst of 16-byte cp.async (one global-latency exposure per block instead of one per chunk),
//   * the contraction is split over the 8 warps ALONG K; each lane owns a 4 x 4 block of the 16 x 32 score tile and reads
//     4 + 4 float4 per 64 FMA (rows 4 banks apart: conflict-free LDS.128), the 8 partial tiles are summed in warp order
//     through shared memory (deterministic),
//   * the second contraction reads the context block it already holds in shared memory (when S <= 32 and H <= 512;
//     longer sources / wider layers loop over blocks and re-read the context from L2).
// Needs H % 4 == 0 and 16-byte aligned operands (else v1).  Arithmetic stays exact fp32.
constexpr int V2_KCH = 512;
constexpr int V2_PS = 40;          // partial-tile row stride: bank = 8 * (row % 4) + col % 8 over a warp's 32 lanes

__device__ __forceinline__ uint32_t att_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void att_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void att_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

size_t v2_smem_bytes(int KP) {
  return (size_t)(48 * (KP + 4) + 8 * 16 * V2_PS + 16 * (SMAX + 4) + SMAX * TT) * sizeof(float);
}

template <int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
attn_v2_kernel(const float* __restrict__ X, const float* __restrict__ ctx, const int64_t* __restrict__ lengths,
               float* __restrict__ align, float* __restrict__ dscore, float* __restrict__ out, int T, int B, int S, int H,
               int ntile, int KP) {
  extern __shared__ __align__(16) float att_smem[];
  const int ST = KP + 4;                             // row stride: 16-byte aligned, consecutive rows 4 banks apart
  float* Xs = att_smem;                              // [16][ST]
  float* Cs = Xs + 16 * ST;                          // [32][ST]
  float* Pp = Cs + 32 * ST;                          // [8 warps][16][V2_PS]
  float* Pf = Pp + 8 * 16 * V2_PS;                   // [16][SMAX + 4] scores of the whole source
  float* Ws = Pf + 16 * (SMAX + 4);                  // [SMAX][16] weights, query index fastest
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / ntile, t0 = (blockIdx.x % ntile) * TT;
  const int len = lengths ? min((int)lengths[b], S) : S;
  const int nsb = (S + 31) / 32, nkc = (H + KP - 1) / KP;
  const int rg = lane >> 3, cg = lane & 7;           // this lane's rows rg + 4 i, columns cg + 8 i
  const int KQ = KP / 4;                             // float4 per staged row
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sb = 0; sb < nsb; ++sb) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int i2 = 0; i2 < 4; ++i2) acc[i][i2] = 0.f;
    for (int kc = 0; kc < nkc; ++kc) {
      const int k0 = kc * KP, kl = min(KP, H - k0);  // valid floats of this chunk (multiple of 4)
      __syncthreads();                               // the previous block / chunk has been consumed
      if (nkc > 1 || sb == 0) {
        for (int e = tid; e < 16 * KQ; e += TILE_THREADS) {
          const int r = e / KQ, k = (e % KQ) * 4, t = t0 + r;
          float* dst = Xs + r * ST + k;
          if (t < T && k < kl) att_cp_async16(att_smem_u32(dst), X + ((size_t)t * B + b) * H + k0 + k);
          else *reinterpret_cast<float4*>(dst) = zero4;
        }
      }
      for (int e = tid; e < 32 * KQ; e += TILE_THREADS) {
        const int r = e / KQ, k = (e % KQ) * 4, j = sb * 32 + r;
        float* dst = Cs + r * ST + k;
        if (j < S && k < kl) att_cp_async16(att_smem_u32(dst), ctx + ((size_t)j * B + b) * H + k0 + k);
        else *reinterpret_cast<float4*>(dst) = zero4;
      }
      att_cp_async_wait_all();
      __syncthreads();
      const float* xb = Xs + rg * ST + warp * (KP / 8);
      const float* cb = Cs + cg * ST + warp * (KP / 8);
#pragma unroll 2
      for (int kk = 0; kk < KP / 8; kk += 4) {
        float4 xv[4], cv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xb + 4 * i * ST + kk);
#pragma unroll
        for (int i = 0; i < 4; ++i) cv[i] = *reinterpret_cast<const float4*>(cb + 8 * i * ST + kk);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            acc[i][i2] = fmaf(xv[i].x, cv[i2].x, acc[i][i2]);
            acc[i][i2] = fmaf(xv[i].y, cv[i2].y, acc[i][i2]);
            acc[i][i2] = fmaf(xv[i].z, cv[i2].z, acc[i][i2]);
            acc[i][i2] = fmaf(xv[i].w, cv[i2].w, acc[i][i2]);
          }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int i2 = 0; i2 < 4; ++i2) Pp[(warp * 16 + rg + 4 * i) * V2_PS + cg + 8 * i2] = acc[i][i2];
    __syncthreads();
    for (int e = tid; e < 16 * 32; e += TILE_THREADS) {
      const int r = e >> 5, c = e & 31;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += Pp[(w * 16 + r) * V2_PS + c];
      Pf[r * (SMAX + 4) + sb * 32 + c] = sum;
    }
    // (Pp is rewritten only after the next block's staging barriers)
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int tt = 2 * warp + i, t = t0 + tt;
    float v[4];
    if (MODE == 0) {
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = lane + 32 * jj;
        v[jj] = (j < len) ? Pf[tt * (SMAX + 4) + j] : -INFINITY;
        mx = fmaxf(mx, v[jj]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) { v[jj] = expf(v[jj] - mx); sum += v[jj]; }   // exp(-inf) = 0
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) v[jj] *= inv;
    } else {
      float a[4], dot = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = lane + 32 * jj;
        a[jj] = (t < T && j < len) ? align[((size_t)t * B + b) * S + j] : 0.f;
        v[jj] = (j < len) ? Pf[tt * (SMAX + 4) + j] : 0.f;
        dot = fmaf(a[jj], v[jj], dot);
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) v[jj] = a[jj] * (v[jj] - dot);
    }
    float* gout = (MODE == 0) ? align : dscore;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      if (j < nsb * 32) Ws[j * TT + tt] = v[jj];
      if (t < T && j < S) gout[((size_t)t * B + b) * S + j] = v[jj];
    }
  }
  __syncthreads();
  if (nsb == 1 && nkc == 1) {                        // the context block is still resident
    for (int k = tid; k < H; k += TILE_THREADS) {
      float o[TT];
#pragma unroll
      for (int i = 0; i < TT; ++i) o[i] = 0.f;
      for (int j = 0; j < len; ++j) {
        const float c = Cs[j * ST + k];
        const float4* w4 = reinterpret_cast<const float4*>(Ws + j * TT);
#pragma unroll
        for (int q = 0; q < TT / 4; ++q) {
          const float4 w = w4[q];
          o[4 * q] = fmaf(w.x, c, o[4 * q]); o[4 * q + 1] = fmaf(w.y, c, o[4 * q + 1]);
          o[4 * q + 2] = fmaf(w.z, c, o[4 * q + 2]); o[4 * q + 3] = fmaf(w.w, c, o[4 * q + 3]);
        }
      }
#pragma unroll
      for (int i = 0; i < TT; ++i)
        if (t0 + i < T) out[((size_t)(t0 + i) * B + b) * H + k] = o[i];
    }
  } else {
    wv_tile(reinterpret_cast<const float (*)[TT]>(Ws), ctx, out, t0, T, b, B, len, H);
  }
}

// Used where it wins: H <= 512 (one contraction chunk) and S <= 64 (at most two source blocks).  Measured at the cfg5
// shape (T = S = 80, H = 1024, B = 512: 3 blocks x 2 chunks, each with its own staging barriers and one CTA per SM) v2 is
// 18 % slower than the register-staged v1 kernels (1.19 vs 1.01 ms forward); at cfg1 (S = 30, H = 500) 19 vs 34 us.
bool v2_ok(const void* a, const void* b, int T, int S, int H) {
  return T > 1 && H <= V2_KCH && S <= 64 && (H & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0 &&
         !getenv("VMMT_ATTN_V1");
}

template <int MODE>
int launch_v2(const float* X, const float* ctx, const int64_t* lengths, float* align, float* dscore, float* out, int T,
              int B, int S, int H, cudaStream_t st) {
  const int KP = min(V2_KCH, ((H + 31) / 32) * 32);
  const size_t smem = v2_smem_bytes(KP);
  {
    static std::mutex mu;                                    // function attributes are sticky: once per device
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && !done[dev]) {
      VMMT_CUDA(cudaFuncSetAttribute(attn_v2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)v2_smem_bytes(V2_KCH)));
      done[dev] = true;
    }
  }
  const int ntile = ceil_div(T, TT);
  attn_v2_kernel<MODE><<<B * ntile, TILE_THREADS, smem, st>>>(X, ctx, lengths, align, dscore, out, T, B, S, H, ntile, KP);
  return vmmt_check_launch(MODE == 0 ? "attn_v2_kernel<fwd>" : "attn_v2_kernel<bwd>");
}

size_t tiled_smem_bytes(int S) {
  const int srows = ((S + 31) / 32) * 32;
  return (size_t)(TT * (KC + 1) + srows * (KC + 1) + srows * TT) * sizeof(float);
}

}  // namespace

extern "C" int vmmt_attention_fwd(const float* qp, const float* ctx, const int64_t* lengths,
                                  float* align, float* cvec, int T, int B, int S, int H,
                                  void* stream) {
  VMMT_REQUIRE(S >= 1 && S <= SMAX, "attention_fwd: src_len %d outside [1,%d]", S, SMAX);
  VMMT_REQUIRE(T >= 1 && B >= 1 && H >= 1, "attention_fwd: bad dims");
  if (v2_ok(qp, ctx, T, S, H) && !getenv("VMMT_ATTN_WARP"))
    return launch_v2<0>(qp, ctx, lengths, align, nullptr, cvec, T, B, S, H, (cudaStream_t)stream);
  if (T > 1 && !getenv("VMMT_ATTN_WARP")) {          // sequence mode: tiled kernel (one CTA per 16 queries of one row)
    const int ntile = ceil_div(T, TT);
    attn_fwd_tiled_kernel<<<B * ntile, TILE_THREADS, tiled_smem_bytes(S), (cudaStream_t)stream>>>(
        qp, ctx, lengths, align, cvec, T, B, S, H, ntile);
    return vmmt_check_launch("attn_fwd_tiled_kernel");
  }
  int tsplit = ceil_div(T, ATT_WARPS);
  const int want = ceil_div(2 * vmmt_num_sms(), B);
  if (tsplit > want) tsplit = want;
  if (tsplit < 1) tsplit = 1;
  attn_fwd_kernel<<<B * tsplit, ATT_WARPS * 32, 0, (cudaStream_t)stream>>>(
      qp, ctx, lengths, align, cvec, T, B, S, H, tsplit);
  return vmmt_check_launch("attn_fwd_kernel");
}

// Backward in two launches: the QUERY side (d(scores) and dqp: what the decoder's backward chain waits for) and the CONTEXT
// side (dctx: only the encoders' backward needs it).  vmmt_attention_bwd issues both on one stream; the two-entry form lets
// the caller put the context side on another stream so that it does not sit between the query side and its consumer.
extern "C" int vmmt_attention_bwd_query(const float* dcvec, const float* ctx, const float* align, const int64_t* lengths,
                                        float* dscore_ws, float* dqp, int T, int B, int S, int H, void* stream) {
  VMMT_REQUIRE(S >= 1 && S <= SMAX, "attention_bwd: src_len %d outside [1,%d]", S, SMAX);
  cudaStream_t st = (cudaStream_t)stream;
  if (!getenv("VMMT_ATTN_WARP")) {
    if (v2_ok(dcvec, ctx, T, S, H) && (reinterpret_cast<uintptr_t>(dqp) & 15) == 0)
      return launch_v2<1>(dcvec, ctx, lengths, const_cast<float*>(align), dscore_ws, dqp, T, B, S, H, st);
    const int ntile = ceil_div(T, TT);
    attn_bwd_query_tiled_kernel<<<B * ntile, TILE_THREADS, tiled_smem_bytes(S), st>>>(dcvec, ctx, align, lengths,
                                                                                      dscore_ws, dqp, T, B, S, H, ntile);
    return vmmt_check_launch("attn_bwd_query_tiled_kernel");
  }
  int tsplit = ceil_div(T, ATT_WARPS);
  const int want = ceil_div(2 * vmmt_num_sms(), B);
  if (tsplit > want) tsplit = want;
  if (tsplit < 1) tsplit = 1;
  attn_bwd_query_kernel<<<B * tsplit, ATT_WARPS * 32, 0, st>>>(dcvec, ctx, align, lengths, dscore_ws, dqp, T, B, S, H, tsplit);
  return vmmt_check_launch("attn_bwd_query_kernel");
}

extern "C" int vmmt_attention_bwd_ctx(const float* dcvec, const float* qp, const float* align, const float* dscore_ws,
                                      float* dctx, int accumulate_dctx, int T, int B, int S, int H, void* stream) {
  VMMT_REQUIRE(S >= 1 && S <= SMAX, "attention_bwd: src_len %d outside [1,%d]", S, SMAX);
  cudaStream_t st = (cudaStream_t)stream;
  if (!getenv("VMMT_ATTN_WARP")) {
    const int nk = ceil_div(H, 64);
    attn_bwd_ctx_tiled_kernel<<<B * nk, TILE_THREADS, 0, st>>>(dcvec, qp, align, dscore_ws, dctx, T, B, S, H,
                                                               accumulate_dctx, nk);
    return vmmt_check_launch("attn_bwd_ctx_tiled_kernel");
  }
  attn_bwd_ctx_kernel<<<S * B, 128, 0, st>>>(dcvec, qp, align, dscore_ws, dctx, T, B, S, H, accumulate_dctx);
  return vmmt_check_launch("attn_bwd_ctx_kernel");
}

extern "C" int vmmt_attention_bwd(const float* dcvec, const float* qp, const float* ctx,
                                  const float* align, const int64_t* lengths, float* dscore_ws,
                                  float* dqp, float* dctx, int accumulate_dctx, int T, int B, int S,
                                  int H, void* stream) {
  int rc = vmmt_attention_bwd_query(dcvec, ctx, align, lengths, dscore_ws, dqp, T, B, S, H, stream);
  if (rc) return rc;
  return vmmt_attention_bwd_ctx(dcvec, qp, align, dscore_ws, dctx, accumulate_dctx, T, B, S, H, stream);
}
