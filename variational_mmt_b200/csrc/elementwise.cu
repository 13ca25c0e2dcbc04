// Memory-bound kernels of the VI-model-1 step: embedding gather / scatter, masked mean,
// activation derivatives, bias-gradient column sums, Philox dropout, the latent block
// (sample + analytic KL) and the image-feature head's gate and loss.
#include <cooperative_groups.h>
#include <stdio.h>
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

// ---------------------------------------------------------------- embeddings (Embeddings.py:169-188)
__global__ void embedding_fwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ table,
                                     float* __restrict__ out, int64_t n, int E, int64_t rows) {
  const int64_t row = blockIdx.x;
  const int64_t id = idx[row];
  if (id < 0 || id >= rows) {                        // nn.Embedding / numpy fancy indexing raise; here: a device-side trap
    if (threadIdx.x == 0) printf("vmmt_embedding_fwd: index %lld out of range [0, %lld) at position %lld\n", (long long)id, (long long)rows, (long long)row);
    __trap();
  }
  const float* src = table + id * (int64_t)E;
  float* dst = out + row * E;
  for (int k = threadIdx.x; k < E; k += blockDim.x) dst[k] = src[k];
}
// dense scatter-add; the padding row receives no gradient (nn.Embedding padding_idx)
__global__ void embedding_bwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dout,
                                     float* __restrict__ dtable, int64_t n, int E, int64_t pad, int64_t rows) {
  const int64_t row = blockIdx.x;
  const int64_t id = idx[row];
  if (id == pad) return;
  if (id < 0 || id >= rows) {
    if (threadIdx.x == 0) printf("vmmt_embedding_bwd: index %lld out of range [0, %lld) at position %lld\n", (long long)id, (long long)rows, (long long)row);
    __trap();
  }
  const float* src = dout + row * E;
  float* dst = dtable + id * (int64_t)E;
  for (int k = threadIdx.x; k < E; k += blockDim.x) atomicAdd(dst + k, src[k]);
}

// ---------------------------------------------------------------- masked mean (NormalVariationalEncoder.py:65-84)
// x[t, b, k] at x + t * st + b * sb + k (time-major [T,B,H]: st = B*H, sb = H; the transposed view of a [B,T,H] tensor: st = H,
// sb = T*H -- the target encoder's output needs no transposing copy).  One CTA per (b, 128 hidden units); the t-loop keeps
// 8 independent loads in flight and adds them in t order (the sum is the same sequential fp32 sum for every launch shape).
__global__ void __launch_bounds__(128)
masked_mean_fwd_kernel(const float* __restrict__ x, int64_t st, int64_t sb, const int64_t* __restrict__ len,
                       float* __restrict__ out, int64_t out_ld, int T, int H) {
  const int b = blockIdx.x, k = blockIdx.y * 128 + threadIdx.x;
  if (k >= H) return;
  const int L = min((int)len[b], T);
  const float* p = x + (size_t)b * sb + k;
  float s = 0.f;
  int t = 0;
  for (; t + 8 <= L; t += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (size_t)(t + u) * st);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; t < L; ++t) s += __ldg(p + (size_t)t * st);
  out[(size_t)b * out_ld + k] = s / (float)len[b];
}
__global__ void masked_mean_bwd_kernel(const float* __restrict__ dout, int64_t dout_ld,
                                       const int64_t* __restrict__ len, float* __restrict__ dx, int64_t st, int64_t sb,
                                       int T, int B, int H, int accumulate) {
  const int t = blockIdx.x / B, b = blockIdx.x % B;
  const bool on = t < (int)len[b];
  const float inv = 1.0f / (float)len[b];
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    const float g = on ? dout[(size_t)b * dout_ld + k] * inv : 0.f;
    float* o = dx + (size_t)t * st + (size_t)b * sb + k;
    *o = accumulate ? (*o + g) : g;
  }
}

// ---------------------------------------------------------------- activation derivatives (from outputs)
__device__ __forceinline__ float act_grad(float g, float v, int act) {
  switch (act) {
    case VMMT_ACT_RELU: return v > 0.f ? g : 0.f;
    case VMMT_ACT_TANH: return g * (1.f - v * v);
    case VMMT_ACT_SOFTPLUS: return g * (1.f - expf(-v));     // sigmoid(pre) = 1 - exp(-softplus)
    case VMMT_ACT_SIGMOID: return g * v * (1.f - v);
    default: return g;
  }
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                               float* __restrict__ dx, int64_t n, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dx[i] = act_grad(dy[i], y[i], act);
}
// 16-byte aligned buffers: four elements per thread
__global__ void __launch_bounds__(256)
act_bwd4_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, float4* __restrict__ dx,
                int64_t n4, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 g = dy[i], v = y[i];
  dx[i] = make_float4(act_grad(g.x, v.x, act), act_grad(g.y, v.y, act), act_grad(g.z, v.z, act),
                      act_grad(g.w, v.w, act));
}

// column sums of a row-major [M,N] matrix (bias gradients), accumulated into out[N]
__global__ void colsum_kernel(const float* __restrict__ a, int64_t lda, int M, int N,
                              float* __restrict__ out, float* __restrict__ out2, int rows_per_block) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += a[(size_t)m * lda + n];
  if (gridDim.y == 1) { out[n] += s; if (out2) out2[n] += s; }
  else { atomicAdd(out + n, s); if (out2) atomicAdd(out2 + n, s); }
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float alpha, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(alpha, x[i], y[i]);
}

// ---------------------------------------------------------------- dropout (Philox, mask regenerated in bwd)
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p,
                               uint64_t seed, uint64_t offset, const uint64_t* __restrict__ base) {
  if (base) offset += *base;                                  // device-resident step counter (CUDA-graph replays)
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  const uint4 r = Philox(seed)((uint64_t)i4, offset);
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
  const float scale = 1.0f / (1.0f - p);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t i = i4 * 4 + e;
    if (i < n) y[i] = (u32_to_unit(rr[e]) > p) ? x[i] * scale : 0.f;
  }
}

// ---------------------------------------------------------------- latent block
// z = mu + sd * eps (Dists.py:21-26; eps injected or drawn from Philox via Box-Muller)
__global__ void sample_kernel(const float* __restrict__ mu, const float* __restrict__ sd,
                              const float* __restrict__ eps, float* __restrict__ z, int64_t n,
                              uint64_t seed, uint64_t offset, const uint64_t* __restrict__ base) {
  if (base) offset += *base;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e;
  if (eps) {
    e = eps[i];
  } else {
    const uint4 r = Philox(seed)((uint64_t)i, offset);
    e = sqrtf(-2.0f * logf(u32_to_unit(r.x))) * cospif(2.0f * u32_to_unit(r.y));
  }
  z[i] = fmaf(sd[i], e, mu[i]);
}

// KL[N(mu_q,sd_q) || N(mu_p,sd_p)] summed over Z, mean over B (VILoss.py:439-460).
// One block; out[0] = KL.  mu_p / sd_p may be null (standard normal prior, Models.py:936-939).
// One thread-block cluster of KL_CTAS CTAs: each reduces a contiguous slice, the block sums meet in CTA 0's shared
// memory through DSMEM and are added in rank order (deterministic; no global scratch, no atomics).
constexpr int KL_CTAS = 8;
__global__ void __cluster_dims__(KL_CTAS, 1, 1) __launch_bounds__(512)
kl_fwd_kernel(const float* __restrict__ mq, const float* __restrict__ sq,
              const float* __restrict__ mp, const float* __restrict__ sp,
              float* __restrict__ out, int B, int Z) {
  __shared__ float red[32];
  __shared__ float part[KL_CTAS];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int n = B * Z;
  const int per = (n + KL_CTAS - 1) / KL_CTAS;
  const int lo = rank * per, hi = min(n, lo + per);
  float s = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float m2 = mp ? mp[i] : 0.f, s2 = sp ? sp[i] : 1.f;
    const float d = mq[i] - m2, v1 = sq[i] * sq[i], v2 = s2 * s2;
    s += 0.5f / v2 * (d * d + v1 - v2) + logf(s2) - logf(sq[i]);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) *cluster.map_shared_rank(&part[rank], 0) = s;
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0) {
    float t = 0.f;
    for (int r = 0; r < KL_CTAS; ++r) t += part[r];
    out[0] = t / (float)B;
  }
}
// gradients scaled by *gscale (device scalar) * kl_weight / B
__global__ void kl_bwd_kernel(const float* __restrict__ mq, const float* __restrict__ sq,
                              const float* __restrict__ mp, const float* __restrict__ sp,
                              float* __restrict__ dmq, float* __restrict__ dsq,
                              float* __restrict__ dmp, float* __restrict__ dsp,
                              const float* __restrict__ gscale, float scale, int B, int Z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Z) return;
  if (gscale) scale *= gscale[0];
  const float m2 = mp ? mp[i] : 0.f, s2 = sp ? sp[i] : 1.f;
  const float d = mq[i] - m2, s1 = sq[i], v2 = s2 * s2;
  const float w = scale / (float)B;
  dmq[i] = w * d / v2;
  dsq[i] = w * (s1 / v2 - 1.f / s1);
  if (dmp) dmp[i] = -w * d / v2;
  if (dsp) dsp[i] = w * (-(d * d + s1 * s1) / (v2 * s2) + 1.f / s2);
}

// ---------------------------------------------------------------- image head
// g = sigmoid(z.w + b) per row; gated = z * g  (NormalVariationalEncoder.py:286-299)
__global__ void gate_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ gate,
                                float* __restrict__ gated, int Z) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int k = threadIdx.x; k < Z; k += blockDim.x) s = fmaf(z[(size_t)b * Z + k], w[k], s);
  s = block_sum(s, red);
  const float g = sigmoidf_(s + bias[0]);
  if (threadIdx.x == 0) gate[b] = g;
  for (int k = threadIdx.x; k < Z; k += blockDim.x) gated[(size_t)b * Z + k] = z[(size_t)b * Z + k] * g;
}
// z is a constant in backward (hazard H2): only the gate parameters receive gradient.
// dpre[b] = (sum_k dgated[b,k] z[b,k]) g (1-g);  dw += sum_b dpre[b] z[b,:];  db += sum_b dpre[b]
__global__ void gate_bwd_row_kernel(const float* __restrict__ dgated, const float* __restrict__ z,
                                    const float* __restrict__ gate, float* __restrict__ dpre, int Z) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int k = threadIdx.x; k < Z; k += blockDim.x)
    s = fmaf(dgated[(size_t)b * Z + k], z[(size_t)b * Z + k], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) dpre[b] = s * gate[b] * (1.f - gate[b]);
}
__global__ void gate_bwd_param_kernel(const float* __restrict__ dpre, const float* __restrict__ z,
                                      float* __restrict__ dw, float* __restrict__ db, int B, int Z) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < Z) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dpre[b], z[(size_t)b * Z + k], s);
    dw[k] += s;
  }
  if (k == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dpre[b];
    db[0] += s;
  }
}

// Image loss (VILoss.py:22-56,289-296,317-332; hazard H3).  Per row: L2-normalise loc and v,
// cos_b = p^.v^, sq_b = sum_d (p^-v^)^2.  rowstats[b] = {cos_b, sq_b, |loc_b|, |v_b|}.
__global__ void image_loss_row_kernel(const float* __restrict__ loc, const float* __restrict__ v,
                                      float* __restrict__ rowstats, int D) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* p = loc + (size_t)b * D;
  const float* q = v + (size_t)b * D;
  float pp = 0.f, qq = 0.f, pq = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    pp = fmaf(p[k], p[k], pp); qq = fmaf(q[k], q[k], qq); pq = fmaf(p[k], q[k], pq);
  }
  pp = block_sum(pp, red); qq = block_sum(qq, red); pq = block_sum(pq, red);
  const float np = sqrtf(pp), nq = sqrtf(qq);
  float sq = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const float d = p[k] / np - q[k] / nq;
    sq = fmaf(d, d, sq);
  }
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) {
    rowstats[b * 4 + 0] = pq / (np * nq);
    rowstats[b * 4 + 1] = sq;
    rowstats[b * 4 + 2] = np;
    rowstats[b * 4 + 3] = nq;
  }
}
// out[0] = image log-prob (sum over B, mean over D), out[1] = mean cosine
__global__ void image_loss_final_kernel(const float* __restrict__ rowstats, float* __restrict__ out,
                                        int B, int D) {
  __shared__ float red[32];
  float c = 0.f, s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) { c += rowstats[b * 4]; s += rowstats[b * 4 + 1]; }
  c = block_sum(c, red); s = block_sum(s, red);
  if (threadIdx.x == 0) {
    out[0] = -0.5f * s / (float)D - (float)B * 0.91893853320467274178f;   // 0.5*log(2*pi)
    out[1] = c / (float)B;
  }
}
// d(-IMG)/dloc * scale.  legacy: (p^-v^)/D passed straight through the normalisation (H3);
// otherwise the true Jacobian J = (I - p^ p^T)/|loc| is applied.
__global__ void image_loss_bwd_kernel(const float* __restrict__ loc, const float* __restrict__ v,
                                      const float* __restrict__ rowstats, float* __restrict__ dloc,
                                      const float* __restrict__ gscale, float scale, int D, int legacy) {
  const int b = blockIdx.x;
  if (gscale) scale *= gscale[0];
  const float np = rowstats[b * 4 + 2], nq = rowstats[b * 4 + 3], cosb = rowstats[b * 4];
  const float w = scale / (float)D;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const float ph = loc[(size_t)b * D + k] / np, vh = v[(size_t)b * D + k] / nq;
    float g = ph - vh;
    if (!legacy) g = (g - ph * (1.f - cosb)) / np;     // (I - p^p^T)(p^ - v^) = p^-v^ - p^(1-cos)
    dloc[(size_t)b * D + k] = w * g;
  }
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" int vmmt_embedding_fwd(const int64_t* idx, int64_t n, const float* table, int64_t rows, int E,
                                  float* out, void* stream) {
  if (n <= 0) return VMMT_OK;
  embedding_fwd_kernel<<<(unsigned)n, 128, 0, ST(stream)>>>(idx, table, out, n, E, rows);
  return vmmt_check_launch("embedding_fwd");
}
extern "C" int vmmt_embedding_bwd(const int64_t* idx, int64_t n, const float* dout, int E,
                                  int64_t pad_idx, float* dtable, int64_t rows, void* stream) {
  if (n <= 0) return VMMT_OK;
  embedding_bwd_kernel<<<(unsigned)n, 128, 0, ST(stream)>>>(idx, dout, dtable, n, E, pad_idx, rows);
  return vmmt_check_launch("embedding_bwd");
}
extern "C" int vmmt_masked_mean_fwd(const float* x, int64_t stride_t, int64_t stride_b, const int64_t* lengths, float* out,
                                    int64_t out_ld, int T, int B, int H, void* stream) {
  if (B <= 0 || H <= 0) return VMMT_OK;
  masked_mean_fwd_kernel<<<dim3(B, ceil_div(H, 128)), 128, 0, ST(stream)>>>(x, stride_t, stride_b, lengths, out, out_ld, T, H);
  return vmmt_check_launch("masked_mean_fwd");
}
extern "C" int vmmt_masked_mean_bwd(const float* dout, int64_t dout_ld, const int64_t* lengths,
                                    float* dx, int64_t stride_t, int64_t stride_b, int accumulate, int T, int B, int H,
                                    void* stream) {
  if (T <= 0 || B <= 0) return VMMT_OK;
  masked_mean_bwd_kernel<<<T * B, 128, 0, ST(stream)>>>(dout, dout_ld, lengths, dx, stride_t, stride_b, T, B, H, accumulate);
  return vmmt_check_launch("masked_mean_bwd");
}
extern "C" int vmmt_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act,
                            void* stream) {
  if (n <= 0) return VMMT_OK;
  if ((n & 3) == 0 && (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dx) & 15) == 0) {
    act_bwd4_kernel<<<ceil_div(n / 4, 256), 256, 0, ST(stream)>>>((const float4*)dy, (const float4*)y, (float4*)dx,
                                                                 n / 4, act);
    return vmmt_check_launch("act_bwd4");
  }
  act_bwd_kernel<<<ceil_div(n, 256), 256, 0, ST(stream)>>>(dy, y, dx, n, act);
  return vmmt_check_launch("act_bwd");
}
extern "C" int vmmt_colsum_acc(const float* a, int64_t lda, int M, int N, float* out, float* out2, void* stream) {
  if (M <= 0 || N <= 0) return VMMT_OK;
  int rows = 64;
  dim3 grid(ceil_div(N, 128), ceil_div(M, rows));
  if (grid.y == 1) rows = M;
  colsum_kernel<<<grid, 128, 0, ST(stream)>>>(a, lda, M, N, out, out2, rows);
  return vmmt_check_launch("colsum");
}
extern "C" int vmmt_axpy(float* y, const float* x, float alpha, int64_t n, void* stream) {
  if (n <= 0) return VMMT_OK;
  axpy_kernel<<<ceil_div(n, 256), 256, 0, ST(stream)>>>(y, x, alpha, n);
  return vmmt_check_launch("axpy");
}
__global__ void counter_add_kernel(uint64_t* ctr, uint64_t inc) { *ctr += inc; }

// stats = {nll, n_words, n_correct, kl, img_logprob, img_cos, kl_after, elbo}: objective and the two derived entries
// in one launch (onmt/VILoss.py:462-496: loss = nll - img_logprob + kl_weight * kl)
__global__ void loss_finalize_kernel(float* __restrict__ stats, float kl_weight, float* __restrict__ loss) {
  const float l = stats[0] - stats[4] + kl_weight * stats[3];
  stats[6] = stats[3] * kl_weight;
  stats[7] = l;
  loss[0] = l;
}
extern "C" int vmmt_loss_finalize(float* stats8, float kl_weight, float* loss1, void* stream) {
  loss_finalize_kernel<<<1, 1, 0, ST(stream)>>>(stats8, kl_weight, loss1);
  return vmmt_check_launch("loss_finalize");
}

__global__ void stamp_kernel(unsigned long long* buf, int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  buf[slot] = t;
}
extern "C" int vmmt_stamp(uint64_t* buf, int slot, void* stream) {
  stamp_kernel<<<1, 1, 0, ST(stream)>>>(reinterpret_cast<unsigned long long*>(buf), slot);
  return vmmt_check_launch("stamp");
}
extern "C" int vmmt_counter_add(uint64_t* ctr, uint64_t inc, void* stream) {
  counter_add_kernel<<<1, 1, 0, ST(stream)>>>(ctr, inc);
  return vmmt_check_launch("counter_add");
}

extern "C" int vmmt_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed,
                            uint64_t offset, const uint64_t* base, void* stream) {
  if (n <= 0) return VMMT_OK;
  VMMT_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f outside [0,1)", (double)p);
  dropout_kernel<<<ceil_div(ceil_div(n, 4), 256), 256, 0, ST(stream)>>>(x, y, n, p, seed, offset, base);
  return vmmt_check_launch("dropout");
}
extern "C" int vmmt_normal_sample(const float* mu, const float* sd, const float* eps, float* z,
                                  int64_t n, uint64_t seed, uint64_t offset, const uint64_t* base, void* stream) {
  sample_kernel<<<ceil_div(n, 256), 256, 0, ST(stream)>>>(mu, sd, eps, z, n, seed, offset, base);
  return vmmt_check_launch("normal_sample");
}
extern "C" int vmmt_kl_fwd(const float* mu_q, const float* sd_q, const float* mu_p,
                           const float* sd_p, float* out, int B, int Z, void* stream) {
  kl_fwd_kernel<<<KL_CTAS, 512, 0, ST(stream)>>>(mu_q, sd_q, mu_p, sd_p, out, B, Z);
  return vmmt_check_launch("kl_fwd");
}
extern "C" int vmmt_kl_bwd(const float* mu_q, const float* sd_q, const float* mu_p,
                           const float* sd_p, float* dmu_q, float* dsd_q, float* dmu_p,
                           float* dsd_p, const float* gscale, float scale, int B, int Z, void* stream) {
  kl_bwd_kernel<<<ceil_div((int64_t)B * Z, 256), 256, 0, ST(stream)>>>(mu_q, sd_q, mu_p, sd_p, dmu_q, dsd_q,
                                                              dmu_p, dsd_p, gscale, scale, B, Z);
  return vmmt_check_launch("kl_bwd");
}
extern "C" int vmmt_gate_fwd(const float* z, const float* w, const float* bias, float* gate,
                             float* gated, int B, int Z, void* stream) {
  gate_fwd_kernel<<<B, 128, 0, ST(stream)>>>(z, w, bias, gate, gated, Z);
  return vmmt_check_launch("gate_fwd");
}
extern "C" int vmmt_gate_bwd(const float* dgated, const float* z, const float* gate, float* dpre_ws,
                             float* dw, float* db, int B, int Z, void* stream) {
  gate_bwd_row_kernel<<<B, 128, 0, ST(stream)>>>(dgated, z, gate, dpre_ws, Z);
  int rc = vmmt_check_launch("gate_bwd_row");
  if (rc) return rc;
  gate_bwd_param_kernel<<<ceil_div(Z, 128), 128, 0, ST(stream)>>>(dpre_ws, z, dw, db, B, Z);
  return vmmt_check_launch("gate_bwd_param");
}
extern "C" int vmmt_image_loss_fwd(const float* loc, const float* v, float* rowstats, float* out2,
                                   int B, int D, void* stream) {
  image_loss_row_kernel<<<B, 256, 0, ST(stream)>>>(loc, v, rowstats, D);
  int rc = vmmt_check_launch("image_loss_row");
  if (rc) return rc;
  image_loss_final_kernel<<<1, 256, 0, ST(stream)>>>(rowstats, out2, B, D);
  return vmmt_check_launch("image_loss_final");
}
extern "C" int vmmt_image_loss_bwd(const float* loc, const float* v, const float* rowstats,
                                   float* dloc, const float* gscale, float scale,
                                   int legacy_passthrough, int B, int D, void* stream) {
  image_loss_bwd_kernel<<<B, 256, 0, ST(stream)>>>(loc, v, rowstats, dloc, gscale, scale, D, legacy_passthrough);
  return vmmt_check_launch("image_loss_bwd");
}
