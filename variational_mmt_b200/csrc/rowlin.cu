// Row-block linear layers in EXACT fp32 for the batch-row networks of VI model 1: the prior / posterior location and
// scale MLPs (onmt/modules/NormalVariationalEncoder.py:12-43, 93-110, 164-228) and the image-feature head
// (NormalVariationalEncoder.py:286-304).  These layers have M = batch rows (40) against K up to 3048 and N up to 2048:
// they are weight-bandwidth bound, a 128-row tensor-core tile is 70 % padding, and the TF32 operand rounding of the
// K = 3048 posterior layer moved z by 5e-3 (measured: attention max-rel 1.27e-3 > north_star's 1e-3, tools/parity_probe.py).
// Here every product is an fp32 FMA, the contraction is split over a thread-block CLUSTER (split-K through distributed
// shared memory, summed in rank order: deterministic, no atomics, no init / finish launches) and each output row's
// arithmetic is independent of how many rows there are (a sentence alone = the same sentence in a batch).
//
//   forward form  (w_transposed = 0): out_p[M,N] = act_p( x_p[M,K] W_p[N,K]^T + b_p ),  W rows = output columns
//   gradient form (w_transposed = 1): out_p[M,N] =        x_p[M,K] W_p[K,N],            W rows = contraction index,
//                                     x_p := dy_p * act'(y_p) applied while the tile is loaded (and optionally written
//                                     out: it is d(pre-activation), what the weight / bias gradients consume)
// Up to two problems per launch: either independent (own x, own output: grid.z) or SUMMED into one output
// (the two heads of an MLP pair sharing their input: dx = dpre_loc W_loc + dpre_scale W_scale).
// x may be the column-wise concatenation of up to three matrices ([mean(x) ; mean(y) ; v], Models.py:911) so that the
// concatenation is never materialised.
#include <stdlib.h>
#include <mutex>
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

constexpr int BM = 40;        // rows per CTA (8 row groups x 5)
constexpr int BN = 32;        // output columns per CTA (16 column groups x 2)
constexpr int KC = 32;        // contraction chunk per pipeline stage
constexpr int TS = 36;        // padded tile row stride (floats): 16-byte aligned rows, conflict-free LDS.128 / LDS.64
constexpr int NT = 256;       // threads: two groups of 128, each contracting one half of every 32-wide chunk

struct Params {
  VmmtRowLin pr[2];
  int nprob, sum, M, N, K, KS;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(uint32_t addr, uint32_t rank) {
  uint32_t r; float v;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r));
  return v;
}

// up to 4 consecutive floats p[0..3], `valid` of them in range (the rest 0); vector load when possible
__device__ __forceinline__ float4 ld4(const float* p, int valid) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid >= 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  if (valid > 0) v.x = __ldg(p);
  if (valid > 1) v.y = __ldg(p + 1);
  if (valid > 2) v.z = __ldg(p + 2);
  if (valid > 3) v.w = __ldg(p + 3);
  return v;
}
__device__ __forceinline__ float dact(float dy, float y, int act) {
  switch (act) {
    case 1: return y > 0.f ? dy : 0.f;                       // relu
    case 2: return dy * (1.f - y * y);                       // tanh
    case 3: return dy * (1.f - expf(-y));                    // softplus: sigma(u) = 1 - exp(-softplus(u))
    case 4: return dy * y * (1.f - y);                       // sigmoid
    default: return dy;
  }
}
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return tanhf(v);
    case 3: return softplusf_(v);
    case 4: return sigmoidf_(v);
    default: return v;
  }
}

// x tile element fetch: 4 consecutive k of row m from the segmented x of problem Q (transformed by act'(y) when y is set)
__device__ __forceinline__ float4 fetch_x(const VmmtRowLin& Q, int m, int k, int M, int K) {
  if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
  int sidx = 0;
#pragma unroll
  for (int i = 1; i < 3; ++i)
    if (i < Q.nseg && k >= Q.seg[i].k0) sidx = i;
  const VmmtRowLinSeg& S = Q.seg[sidx];
  const int kend = (sidx + 1 < Q.nseg) ? Q.seg[sidx + 1].k0 : K;
  float4 v;
  if (k + 4 <= kend) v = ld4(S.p + (size_t)m * S.ld + (k - S.k0), 4);
  else {                                                     // the four values straddle a segment boundary / the end
    float t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = k + i;
      float e = 0.f;
      if (kk < K) {
        int sj = 0;
#pragma unroll
        for (int j = 1; j < 3; ++j)
          if (j < Q.nseg && kk >= Q.seg[j].k0) sj = j;
        e = __ldg(Q.seg[sj].p + (size_t)m * Q.seg[sj].ld + (kk - Q.seg[sj].k0));
      }
      t[i] = e;
    }
    v = make_float4(t[0], t[1], t[2], t[3]);
  }
  if (Q.y) {
    const float4 y = ld4(Q.y + (size_t)m * Q.ldy + k, min(4, K - k));
    v.x = dact(v.x, y.x, Q.yact); v.y = dact(v.y, y.y, Q.yact); v.z = dact(v.z, y.z, Q.yact); v.w = dact(v.w, y.w, Q.yact);
  }
  return v;
}

// Aligned form (every base 16-byte aligned, every pitch / K / N / segment start a multiple of 4 floats): a float4 is entirely
// inside one segment and entirely valid or entirely out of range, so a tile is filled by predicated cp.async copies with no
// control flow around them and NST chunks are in flight per CTA.  (First version: guarded register loads, one chunk in
// flight, the two paths of each guarded load merged through the loaded registers: 2.9 us per 32-wide chunk, measured.)
// 16-byte global -> shared asynchronous copy; `ok` false: the destination is zero-filled (src-size 0)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int NST = 4;                                       // pipeline stages (chunks in flight per CTA)
constexpr int XT = BM * TS, WTILE = 32 * TS;                 // floats per x / y tile and per weight tile
constexpr int STAGE_FLOATS = 2 * XT + WTILE;                 // x | y | w
constexpr size_t SMEM_BYTES = (size_t)(NST * STAGE_FLOATS + 2 * BM * BN) * sizeof(float);

template <bool WT, bool AL>
__global__ void __launch_bounds__(NT) rowlin_kernel(const Params P) {
  extern __shared__ __align__(16) float smem[];
  float* red = smem + NST * STAGE_FLOATS;                    // [2][BM][BN] partial tiles (k-half 0 holds the CTA's sum)
  const int tid = threadIdx.x;
  const int KS = P.KS;
  const int rank = KS > 1 ? (int)cluster_rank() : 0;
  const int nb = blockIdx.x / KS;                            // output column block
  const int n0 = nb * BN, m0 = blockIdx.y * BM;
  const int M = P.M, N = P.N, K = P.K;
  const int kh = tid >> 7;                                   // k-half of each chunk this thread contracts
  const int rg = (tid & 127) >> 4, cg = tid & 15;
  // contraction chunks of this rank
  const int nch = (K + KC - 1) / KC;
  const int per = (nch + KS - 1) / KS;
  const int ch0 = rank * per, ch1 = min(nch, ch0 + per);
  const int pfirst = P.sum ? 0 : blockIdx.z, plast = P.sum ? P.nprob : blockIdx.z + 1;
  const uint32_t smem_base = smem_u32(smem);

  float acc[5][2];
#pragma unroll
  for (int i = 0; i < 5; ++i) acc[i][0] = acc[i][1] = 0.f;

  for (int p = pfirst; p < plast; ++p) {
    const VmmtRowLin& Q = P.pr[p];
    const bool write_xt = Q.xt_out != nullptr && nb == 0;
    const bool has_y = Q.y != nullptr;
    // issue the loads of chunk `ch` into stage `st`.  Aligned form: asynchronous 16-byte copies (zero-filled out of
    // range), NST chunks in flight; generic form (ragged dims / unaligned views): synchronous guarded loads.
    auto issue = [&](int ch, int st) {
      const int k0 = ch * KC;
      float* xs = smem + st * STAGE_FLOATS;
      float* ys = xs + XT;
      float* wsm = xs + 2 * XT;
      const uint32_t xa = smem_base + (uint32_t)(st * STAGE_FLOATS) * 4u;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = tid + j * NT;
        const int row = idx >> 3, kq = (idx & 7) * 4;
        if (idx < BM * 8) {
          const int m = m0 + row, k = k0 + kq;
          const uint32_t off = (uint32_t)(row * TS + kq) * 4u;
          if (AL) {
            const bool ok = m < M && k < K;
            const bool s1 = Q.nseg > 1 && k >= Q.seg[1].k0, s2 = Q.nseg > 2 && k >= Q.seg[2].k0;
            const float* base = s2 ? Q.seg[2].p : (s1 ? Q.seg[1].p : Q.seg[0].p);
            const long long ld = s2 ? Q.seg[2].ld : (s1 ? Q.seg[1].ld : Q.seg[0].ld);
            const int sk0 = s2 ? Q.seg[2].k0 : (s1 ? Q.seg[1].k0 : 0);
            cp_async16(xa + off, base + (size_t)(ok ? m : 0) * ld + (ok ? k - sk0 : 0), ok);
            if (has_y) cp_async16(xa + (uint32_t)XT * 4u + off, Q.y + (size_t)(ok ? m : 0) * Q.ldy + (ok ? k : 0), ok);
          } else {
            const float4 v = fetch_x(Q, m, k, M, K);         // already transformed by act'(y)
            *reinterpret_cast<float4*>(xs + row * TS + kq) = v;
            if (write_xt && m < M) {
              float* o = Q.xt_out + (size_t)m * Q.ld_xt + k;
              if (k < K) o[0] = v.x;
              if (k + 1 < K) o[1] = v.y;
              if (k + 2 < K) o[2] = v.z;
              if (k + 3 < K) o[3] = v.w;
            }
          }
        }
      }
      (void)ys;
      {
        const int idx = tid;                                 // 256 float4 of the 32 x 32 weight tile
        const int r = idx >> 3, c4 = (idx & 7) * 4;
        const uint32_t off = (uint32_t)(2 * XT + r * TS + c4) * 4u;
        // !WT: tile row = output column n0 + r, 4 consecutive k;  WT: tile row = contraction index k0 + r, 4 consecutive columns
        const int wr_ = WT ? k0 + r : n0 + r, wc = WT ? n0 + c4 : k0 + c4;
        const int rlim = WT ? K : N, clim = WT ? N : K;
        const bool ok = wr_ < rlim && wc < clim;
        if (AL) cp_async16(xa + off, Q.w + (size_t)(ok ? wr_ : 0) * Q.ldw + (ok ? wc : 0), ok);
        else *reinterpret_cast<float4*>(wsm + r * TS + c4) =
                 ok ? ld4(Q.w + (size_t)wr_ * Q.ldw + wc, min(4, clim - wc)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
      if (ch0 + s < ch1) issue(ch0 + s, s);
      cp_async_commit();
    }
    for (int ch = ch0; ch < ch1; ++ch) {
      const int st = (ch - ch0) % NST;
      if (ch + NST - 1 < ch1) issue(ch + NST - 1, (ch - ch0 + NST - 1) % NST);   // the stage computed in the previous iteration
      cp_async_commit();
      cp_async_wait<NST - 1>();                              // this thread's copies of chunk ch have landed
      __syncthreads();                                       // ... and everybody else's
      float* xs = smem + st * STAGE_FLOATS;
      if (AL && (has_y || write_xt)) {
        // gradient form: x := dy * act'(y) in place; the transformed tile (d pre-activation) also leaves for xt_out
        const float* ys = has_y ? xs + XT : xs;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int idx = tid + j * NT;
          if (idx < BM * 8) {
            const int row = idx >> 3, kq = (idx & 7) * 4;
            float4 v = *reinterpret_cast<const float4*>(xs + row * TS + kq);
            if (has_y) {
              const float4 y = *reinterpret_cast<const float4*>(ys + row * TS + kq);
              v.x = dact(v.x, y.x, Q.yact); v.y = dact(v.y, y.y, Q.yact); v.z = dact(v.z, y.z, Q.yact); v.w = dact(v.w, y.w, Q.yact);
              *reinterpret_cast<float4*>(xs + row * TS + kq) = v;
            }
            if (write_xt && m0 + row < M && ch * KC + kq < K)
              *reinterpret_cast<float4*>(Q.xt_out + (size_t)(m0 + row) * Q.ld_xt + ch * KC + kq) = v;
          }
        }
        __syncthreads();
      }
      const float* xb = xs + rg * 5 * TS;
      const float* wb = xs + 2 * XT;
#pragma unroll
      for (int kq = 0; kq < KC / 2; kq += 4) {
        const int kk = kh * (KC / 2) + kq;
        float4 xv[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) xv[i] = *reinterpret_cast<const float4*>(xb + i * TS + kk);
        if (!WT) {
          const float4 w0 = *reinterpret_cast<const float4*>(wb + cg * TS + kk);
          const float4 w1 = *reinterpret_cast<const float4*>(wb + (cg + 16) * TS + kk);
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            acc[i][0] = fmaf(xv[i].x, w0.x, acc[i][0]); acc[i][0] = fmaf(xv[i].y, w0.y, acc[i][0]);
            acc[i][0] = fmaf(xv[i].z, w0.z, acc[i][0]); acc[i][0] = fmaf(xv[i].w, w0.w, acc[i][0]);
            acc[i][1] = fmaf(xv[i].x, w1.x, acc[i][1]); acc[i][1] = fmaf(xv[i].y, w1.y, acc[i][1]);
            acc[i][1] = fmaf(xv[i].z, w1.z, acc[i][1]); acc[i][1] = fmaf(xv[i].w, w1.w, acc[i][1]);
          }
        } else {
          const float2 a0 = *reinterpret_cast<const float2*>(wb + (kk + 0) * TS + 2 * cg);
          const float2 a1 = *reinterpret_cast<const float2*>(wb + (kk + 1) * TS + 2 * cg);
          const float2 a2 = *reinterpret_cast<const float2*>(wb + (kk + 2) * TS + 2 * cg);
          const float2 a3 = *reinterpret_cast<const float2*>(wb + (kk + 3) * TS + 2 * cg);
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            acc[i][0] = fmaf(xv[i].x, a0.x, acc[i][0]); acc[i][0] = fmaf(xv[i].y, a1.x, acc[i][0]);
            acc[i][0] = fmaf(xv[i].z, a2.x, acc[i][0]); acc[i][0] = fmaf(xv[i].w, a3.x, acc[i][0]);
            acc[i][1] = fmaf(xv[i].x, a0.y, acc[i][1]); acc[i][1] = fmaf(xv[i].y, a1.y, acc[i][1]);
            acc[i][1] = fmaf(xv[i].z, a2.y, acc[i][1]); acc[i][1] = fmaf(xv[i].w, a3.y, acc[i][1]);
          }
        }
      }
      __syncthreads();                                       // the stage is overwritten by the next iteration's issue
    }
    cp_async_wait<0>();
  }

  // this rank's partial tile -> shared memory; rank r then sums rows [r rpr, (r+1) rpr) over all ranks in rank order
  const int c0 = WT ? 2 * cg : cg, c1 = WT ? 2 * cg + 1 : cg + 16;
  if (kh == 1) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      red[BM * BN + (rg * 5 + i) * BN + c0] = acc[i][0];
      red[BM * BN + (rg * 5 + i) * BN + c1] = acc[i][1];
    }
  }
  __syncthreads();
  if (kh == 0) {                                             // k-half 0 + k-half 1, always in this order
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      red[(rg * 5 + i) * BN + c0] = acc[i][0] + red[BM * BN + (rg * 5 + i) * BN + c0];
      red[(rg * 5 + i) * BN + c1] = acc[i][1] + red[BM * BN + (rg * 5 + i) * BN + c1];
    }
  }
  if (KS > 1) cluster_sync_all(); else __syncthreads();
  const int rpr = (BM + KS - 1) / KS;
  const int r0 = rank * rpr, r1 = min(BM, r0 + rpr);
  const VmmtRowLin& O = P.pr[P.sum ? 0 : blockIdx.z];
  const uint32_t red_addr = smem_u32(red);
  for (int e = tid; e < (r1 - r0) * BN; e += NT) {
    const int row = r0 + e / BN, col = e % BN;
    float v;
    if (KS > 1) {
      float part[8];                                         // all ranks' loads in flight together, summed in rank order
#pragma unroll
      for (int q = 0; q < 8; ++q) part[q] = q < KS ? ld_dsmem(red_addr + (uint32_t)((row * BN + col) * 4), (uint32_t)q) : 0.f;
      v = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) v += part[q];
    } else {
      v = red[row * BN + col];
    }
    const int m = m0 + row, n = n0 + col;
    if (m < M && n < N) {
      if (O.bias) v += __ldg(O.bias + n);
      O.out[(size_t)m * O.ldo + n] = apply_act(v, O.act);
    }
  }
  if (KS > 1) cluster_sync_all();                            // no CTA exits while a peer may still read its partial tile
}

}  // namespace

extern "C" int vmmt_rowlin(const VmmtRowLin* probs, int nprob, int sum_outputs, int w_transposed, int M, int N, int K,
                           void* stream) {
  VMMT_REQUIRE(nprob == 1 || nprob == 2, "rowlin: nprob must be 1 or 2 (got %d)", nprob);
  VMMT_REQUIRE(M > 0 && N > 0 && K > 0, "rowlin: bad dims M=%d N=%d K=%d", M, N, K);
  Params P;
  for (int p = 0; p < nprob; ++p) {
    P.pr[p] = probs[p];
    VMMT_REQUIRE(P.pr[p].nseg >= 1 && P.pr[p].nseg <= 3 && P.pr[p].seg[0].k0 == 0, "rowlin: bad x segments");
    VMMT_REQUIRE(P.pr[p].w && P.pr[p].out, "rowlin: null operand");
  }
  if (nprob == 1) P.pr[1] = probs[0];
  P.nprob = nprob; P.sum = (sum_outputs && nprob == 2) ? 1 : 0;
  P.M = M; P.N = N; P.K = K;
  // split the contraction over a cluster so that every CTA streams <= ~8 chunks of weights (<= 8 ranks: portable size)
  const int nch = ceil_div(K, KC);
  int KS = 1;
  while (KS < 8 && nch > 4 * KS) KS *= 2;
  if (const char* e = getenv("VMMT_ROWLIN_KS")) KS = atoi(e);      // tuning / debugging
  P.KS = KS;
  dim3 grid(ceil_div(N, BN) * KS, ceil_div(M, BM), P.sum ? 1 : nprob);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = (cudaStream_t)stream;
  {
    static std::mutex mu;                                    // function attributes are sticky: once per device
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && !done[dev]) {
      VMMT_CUDA(cudaFuncSetAttribute(rowlin_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      VMMT_CUDA(cudaFuncSetAttribute(rowlin_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      VMMT_CUDA(cudaFuncSetAttribute(rowlin_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      VMMT_CUDA(cudaFuncSetAttribute(rowlin_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      done[dev] = true;
    }
  }
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = KS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = KS > 1 ? 1 : 0;
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  bool al = (K & 3) == 0 && (N & 3) == 0;
  for (int p = 0; p < nprob && al; ++p) {
    const VmmtRowLin& Q = P.pr[p];
    for (int i = 0; i < Q.nseg; ++i) al = al && a16(Q.seg[i].p) && (Q.seg[i].ld & 3) == 0 && (Q.seg[i].k0 & 3) == 0;
    al = al && a16(Q.w) && (Q.ldw & 3) == 0;
    if (Q.y) al = al && a16(Q.y) && (Q.ldy & 3) == 0;
    if (Q.xt_out) al = al && a16(Q.xt_out) && (Q.ld_xt & 3) == 0;
  }
  if (w_transposed) {
    if (al) VMMT_CUDA(cudaLaunchKernelEx(&cfg, rowlin_kernel<true, true>, P));
    else VMMT_CUDA(cudaLaunchKernelEx(&cfg, rowlin_kernel<true, false>, P));
  } else {
    if (al) VMMT_CUDA(cudaLaunchKernelEx(&cfg, rowlin_kernel<false, true>, P));
    else VMMT_CUDA(cudaLaunchKernelEx(&cfg, rowlin_kernel<false, false>, P));
  }
  return vmmt_check_launch("rowlin_kernel");
}
