// Batched prior-only beam search step kernels (device-side beam bookkeeping) and the single-step
// LSTM cell used when the batch (sentences x beam) is too large for the persistent kernel.
//
// Reference: onmt/translate/Beam.py:64-123 (advance), onmt/Models.py:589-594 (beam_update),
// onmt/translate/TranslatorMultimodalVI.py:163-218 (the step loop).  Rows are beam-major
// (row = k*B + b).  alpha = beta = 0 (opts.py:425-429): hypothesis score = summed
// log-prob.  A sentence is frozen once Beam.done() holds for it (the reference decodes one
// sentence per call, so nothing is ever advanced past that point).
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

constexpr int KMAX = 8;

struct Cand { float v; int i; };
__device__ __forceinline__ bool better(float v, int i, float w, int j) {
  return v > w || (v == w && i < j);
}

// Shared tail of the two advance kernels: block-wide top-K over the threads' local lists, then Beam.advance's
// bookkeeping for sentence b (scores, back pointers, tokens, finished hypotheses, done flag).
__device__ __forceinline__ void beam_select_and_update(
    Cand (&loc)[KMAX], Cand* cand, Cand* best, float* rv, int* ri, int* rp, int b, int B, int K, int V, int step,
    int64_t* __restrict__ tok_cur, int32_t* __restrict__ prev_cur, int64_t eos, float* __restrict__ scores,
    int64_t* __restrict__ next_ys, int32_t* __restrict__ prev_ks, float* __restrict__ fin_score,
    int32_t* __restrict__ fin_t, int32_t* __restrict__ fin_k, int32_t* __restrict__ n_fin, int32_t* __restrict__ done,
    int32_t* __restrict__ n_active) {
  const int tid = threadIdx.x;
  for (int q = 0; q < K; ++q) cand[tid * KMAX + q] = loc[q];
  __syncthreads();
  // K rounds of block-wide arg-max over the 256*K local winners
  for (int r = 0; r < K; ++r) {
    float bv = -INFINITY; int bi = 0x7fffffff, bp = -1;
    for (int q = 0; q < K; ++q) {
      const Cand c = cand[tid * KMAX + q];
      if (c.i != 0x7fffffff && better(c.v, c.i, bv, bi)) { bv = c.v; bi = c.i; bp = tid * KMAX + q; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; bp = op; }
    }
    if ((tid & 31) == 0) { rv[tid >> 5] = bv; ri[tid >> 5] = bi; rp[tid >> 5] = bp; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (better(rv[w], ri[w], rv[0], ri[0])) { rv[0] = rv[w]; ri[0] = ri[w]; rp[0] = rp[w]; }
      best[r].v = rv[0]; best[r].i = ri[0];
      if (rp[0] >= 0) cand[rp[0]].i = 0x7fffffff;        // remove the winner
    }
    __syncthreads();
  }
  if (tid == 0) {
    int nf = n_fin[b];
    for (int k = 0; k < K; ++k) {
      const int pk = best[k].i / V, tok = best[k].i - pk * V;
      scores[b * K + k] = best[k].v;
      prev_ks[((size_t)step * K + k) * B + b] = pk;
      next_ys[((size_t)(step + 1) * K + k) * B + b] = tok;
      if (tok_cur) tok_cur[(size_t)k * B + b] = tok;
      if (prev_cur) prev_cur[(size_t)k * B + b] = pk;
      if (tok == (int)eos) {                              // Beam.py:112-118
        if (nf == 0 || best[k].v > fin_score[b]) { fin_score[b] = best[k].v; fin_t[b] = step + 1; fin_k[b] = k; }
        ++nf;
      }
    }
    n_fin[b] = nf;
    const int pk0 = best[0].i / V;
    if (best[0].i - pk0 * V == (int)eos && nf >= 1) {     // eos_top and n_best finished -> done()
      done[b] = 1;
      atomicSub(n_active, 1);
    }
  }
}

// One CTA per sentence.  logp [K*B, V] -> top-K of the flattened [K*V] candidate scores.
__global__ void __launch_bounds__(256)
beam_advance_kernel(const float* __restrict__ logp, int B, int K, int V, int step_host,
                    const int64_t* __restrict__ step_dev,   // device-resident step index (CUDA-graph replays) or null
                    int64_t* __restrict__ tok_cur,          // [K,B] newest tokens (fixed address) or null
                    int32_t* __restrict__ prev_cur,         // [K,B] newest back pointers (fixed address) or null
                    int64_t eos,
                    float* __restrict__ scores,        // [B,K] running hypothesis scores (in/out)
                    int64_t* __restrict__ next_ys,     // [Lmax+1, K, B] tokens; slice `step` is the input
                    int32_t* __restrict__ prev_ks,     // [Lmax, K, B] back pointers
                    float* __restrict__ fin_score,     // [B] best finished score
                    int32_t* __restrict__ fin_t, int32_t* __restrict__ fin_k,   // [B]
                    int32_t* __restrict__ n_fin, int32_t* __restrict__ done,    // [B]
                    int32_t* __restrict__ n_active) {
  const int b = blockIdx.x;
  const int step = step_dev ? (int)*step_dev : step_host;
  if (done[b]) {
    if (threadIdx.x < K && prev_cur) prev_cur[threadIdx.x * B + b] = threadIdx.x;   // frozen sentence: identity reorder
    return;
  }
  __shared__ Cand cand[256 * KMAX];
  __shared__ Cand best[KMAX];
  __shared__ float rv[8];
  __shared__ int ri[8], rp[8];
  const int tid = threadIdx.x;
  Cand loc[KMAX];
#pragma unroll
  for (int q = 0; q < KMAX; ++q) { loc[q].v = -INFINITY; loc[q].i = 0x7fffffff; }
  const int nbeam = (step == 0) ? 1 : K;               // Beam.py:93-94: first step uses beam row 0 only
  for (int k = 0; k < nbeam; ++k) {
    const float base = (step == 0) ? 0.f : scores[b * K + k];
    const bool dead = step > 0 && next_ys[((size_t)step * K + k) * B + b] == eos;   // Beam.py:89-92
    const float* row = logp + ((size_t)k * B + b) * V;
    for (int j = tid; j < V; j += blockDim.x) {
      const float v = dead ? -1e20f : row[j] + base;
      const int id = k * V + j;
      if (better(v, id, loc[K - 1].v, loc[K - 1].i)) {
        loc[K - 1].v = v; loc[K - 1].i = id;
#pragma unroll
        for (int q = KMAX - 1; q > 0; --q)
          if (q < K && better(loc[q].v, loc[q].i, loc[q - 1].v, loc[q - 1].i)) {
            const Cand tmp = loc[q]; loc[q] = loc[q - 1]; loc[q - 1] = tmp;
          }
      }
    }
  }
  beam_select_and_update(loc, cand, best, rv, ri, rp, b, B, K, V, step, tok_cur, prev_cur, eos, scores, next_ys,
                         prev_ks, fin_score, fin_t, fin_k, n_fin, done, n_active);
}

// Fused-generator variant: the generator GEMM's epilogue (gemm_tc.cu, mode 3) left, per row and 128-column tile,
// {max, sum exp} and the tile's best logits.  One CTA per sentence: warp k combines the tile partials of beam row k
// into its log-sum-exp, then the K * ntile * kc candidates (logit - lse + hypothesis score) go through the same
// selection as above.  The global top-K of a [K, V] score matrix is contained in the union of the per-tile top-K
// lists, so the result is the one beam_advance_kernel finds on the materialised log-probs.
__global__ void __launch_bounds__(256)
beam_advance_topk_kernel(const float2* __restrict__ tile_lse, const float2* __restrict__ tile_cand, int ntile, int kc,
                         int B, int K, int V, int step_host, const int64_t* __restrict__ step_dev,
                         int64_t* __restrict__ tok_cur, int32_t* __restrict__ prev_cur, int64_t eos,
                         float* __restrict__ scores, int64_t* __restrict__ next_ys, int32_t* __restrict__ prev_ks,
                         float* __restrict__ fin_score, int32_t* __restrict__ fin_t, int32_t* __restrict__ fin_k,
                         int32_t* __restrict__ n_fin, int32_t* __restrict__ done, int32_t* __restrict__ n_active) {
  const int b = blockIdx.x;
  const int step = step_dev ? (int)*step_dev : step_host;
  if (done[b]) {
    if (threadIdx.x < K && prev_cur) prev_cur[threadIdx.x * B + b] = threadIdx.x;
    return;
  }
  __shared__ Cand cand[256 * KMAX];
  __shared__ Cand best[KMAX];
  __shared__ float rv[8];
  __shared__ int ri[8], rp[8];
  __shared__ float lse_s[KMAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = K * B;
  const int nbeam = (step == 0) ? 1 : K;
  if (warp < nbeam) {                                    // K <= 8 = warps per CTA
    const int row = warp * B + b;
    float mx = -INFINITY, sm = 0.f;
    for (int t = lane; t < ntile; t += 32) {
      const float2 p = tile_lse[(size_t)t * M + row];
      const float nm = fmaxf(mx, p.x);
      sm = sm * __expf(mx - nm) + p.y * __expf(p.x - nm);
      mx = nm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sm, o);
      const float nm = fmaxf(mx, omx);
      const float a = (mx == -INFINITY) ? 0.f : sm * __expf(mx - nm);
      const float c = (omx == -INFINITY) ? 0.f : os * __expf(omx - nm);
      sm = a + c;
      mx = nm;
    }
    if (lane == 0) lse_s[warp] = mx + logf(sm);
  }
  __syncthreads();
  Cand loc[KMAX];
#pragma unroll
  for (int q = 0; q < KMAX; ++q) { loc[q].v = -INFINITY; loc[q].i = 0x7fffffff; }
  auto push = [&](float v, int id) {
    if (better(v, id, loc[K - 1].v, loc[K - 1].i)) {
      loc[K - 1].v = v; loc[K - 1].i = id;
#pragma unroll
      for (int q = KMAX - 1; q > 0; --q)
        if (q < K && better(loc[q].v, loc[q].i, loc[q - 1].v, loc[q - 1].i)) {
          const Cand tmp = loc[q]; loc[q] = loc[q - 1]; loc[q - 1] = tmp;
        }
    }
  };
  const int per_row = ntile * kc;
  for (int k = 0; k < nbeam; ++k) {
    const bool dead = step > 0 && next_ys[((size_t)step * K + k) * B + b] == eos;     // Beam.py:89-92
    if (dead) {                                           // every column scores -1e20: the K lowest columns win
      if (tid < K && tid < V) push(-1e20f, k * V + tid);
      continue;
    }
    const float base = (step == 0) ? 0.f : scores[b * K + k];
    const float lse = lse_s[k];
    const int row = k * B + b;
    for (int i = tid; i < per_row; i += blockDim.x) {
      const int t = i / kc, q = i - t * kc;
      const float2 c = tile_cand[((size_t)t * M + row) * kc + q];
      const int col = __float_as_int(c.y);
      if (col != 0x7fffffff) push((c.x - lse) + base, k * V + col);
    }
  }
  beam_select_and_update(loc, cand, best, rv, ri, rp, b, B, K, V, step, tok_cur, prev_cur, eos, scores, next_ys,
                         prev_ks, fin_score, fin_t, fin_k, n_fin, done, n_active);
}

// new[l, k*B+b, :] = old[l, prev_k[k,b]*B + b, :]   (DecoderState.beam_update)
__global__ void beam_reorder_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                    const int32_t* __restrict__ prev_k, const int32_t* __restrict__ done,
                                    int L, int K, int B, int H) {
  const int r = blockIdx.x;           // destination row in [0, K*B)
  const int l = blockIdx.y;
  const int k = r / B, b = r % B;
  const int pk = done[b] ? k : prev_k[(size_t)k * B + b];
  const float* s = src + ((size_t)l * K * B + (size_t)pk * B + b) * H;
  float* d = dst + ((size_t)l * K * B + r) * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) d[i] = s[i];
}

// single-step cell: G = gpre + biases (+rowbias); c' = f c + i g; h' = o tanh(c')
__global__ void lstm_cell_kernel(const float* __restrict__ gpre, const float* __restrict__ b_ih,
                                 const float* __restrict__ b_hh, const float* __restrict__ rowbias,
                                 const float* __restrict__ c_prev, float* __restrict__ h_out,
                                 float* __restrict__ c_out, int N, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H) return;
  const int n = i / H, u = i % H;
  float G[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const size_t j = (size_t)g * H + u;
    G[g] = gpre[(size_t)n * 4 * H + j] + (b_ih ? b_ih[j] : 0.f) + (b_hh ? b_hh[j] : 0.f) +
           (rowbias ? rowbias[(size_t)n * 4 * H + j] : 0.f);
  }
  const float ig = sigmoidf_(G[0]), fg = sigmoidf_(G[1]), gg = tanhf(G[2]), og = sigmoidf_(G[3]);
  const float c = fg * c_prev[i] + ig * gg;
  c_out[i] = c;
  h_out[i] = og * tanhf(c);
}

}  // namespace

extern "C" int vmmt_beam_advance(const float* logp, int B, int K, int V, int step, const int64_t* step_dev,
                                 int64_t* tok_cur, int32_t* prev_cur, int64_t eos,
                                 float* scores, int64_t* next_ys, int32_t* prev_ks, float* fin_score,
                                 int32_t* fin_t, int32_t* fin_k, int32_t* n_fin, int32_t* done,
                                 int32_t* n_active, void* stream) {
  VMMT_REQUIRE(K >= 1 && K <= KMAX, "beam_advance: beam size %d outside [1,%d]", K, KMAX);
  beam_advance_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(logp, B, K, V, step, step_dev, tok_cur, prev_cur, eos,
                                                          scores, next_ys, prev_ks, fin_score, fin_t, fin_k,
                                                          n_fin, done, n_active);
  return vmmt_check_launch("beam_advance");
}

extern "C" int vmmt_beam_advance_topk(const void* gen_workspace, int B, int K, int V, int step,
                                      const int64_t* step_dev, int64_t* tok_cur, int32_t* prev_cur, int64_t eos,
                                      float* scores, int64_t* next_ys, int32_t* prev_ks, float* fin_score,
                                      int32_t* fin_t, int32_t* fin_k, int32_t* n_fin, int32_t* done,
                                      int32_t* n_active, void* stream) {
  VMMT_REQUIRE(K >= 1 && K <= KMAX, "beam_advance_topk: beam size %d outside [1,%d]", K, KMAX);
  const int ntile = ceil_div(V, 128);
  const size_t M = (size_t)K * B;
  const float2* tile_lse = (const float2*)gen_workspace;           // layout of vmmt_generator_topk
  const float2* tile_cand = tile_lse + (size_t)ntile * M;
  beam_advance_topk_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(tile_lse, tile_cand, ntile, K, B, K, V, step, step_dev,
                                                               tok_cur, prev_cur, eos, scores, next_ys, prev_ks,
                                                               fin_score, fin_t, fin_k, n_fin, done, n_active);
  return vmmt_check_launch("beam_advance_topk");
}

// hist[step][r][:] = cur[r][:]  (attention rows of the current step into the per-step history)
__global__ void beam_record_kernel(const float* __restrict__ cur, float* __restrict__ hist,
                                   const int64_t* __restrict__ step_dev, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) hist[(size_t)*step_dev * n + i] = cur[i];
}

extern "C" int vmmt_beam_record(const float* cur, float* hist, const int64_t* step_dev, int64_t n, void* stream) {
  beam_record_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(cur, hist, step_dev, n);
  return vmmt_check_launch("beam_record");
}

extern "C" int vmmt_beam_reorder(const float* src, float* dst, const int32_t* prev_k_step,
                                 const int32_t* done, int L, int K, int B, int H, void* stream) {
  beam_reorder_kernel<<<dim3(K * B, L), 128, 0, (cudaStream_t)stream>>>(src, dst, prev_k_step, done, L,
                                                                        K, B, H);
  return vmmt_check_launch("beam_reorder");
}

extern "C" int vmmt_lstm_cell_fwd(const float* gates_pre, const float* b_ih, const float* b_hh,
                                  const float* rowbias, const float* c_prev, float* h_out,
                                  float* c_out, int N, int H, void* stream) {
  lstm_cell_kernel<<<ceil_div((int64_t)N * H, 256), 256, 0, (cudaStream_t)stream>>>(
      gates_pre, b_ih, b_hh, rowbias, c_prev, h_out, c_out, N, H);
  return vmmt_check_launch("lstm_cell");
}
