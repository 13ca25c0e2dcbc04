// Persistent LSTM recurrence (forward and BPTT) for small-batch sequences.
//
// Reference semantics: torch nn.LSTM as used by onmt/Models.py:124-149 (encoder, packed = length
// masked), onmt/Models.py:892-893 (bidirectional target encoder, no masking) and
// onmt/VI_Model1.py:106 (decoder, initial state from the encoder); gate order i,f,g,o.
//
// Layout: the input projections gx = x W_ih^T for all timesteps are produced beforehand by one
// batched tensor-core GEMM (vmmt_gemm).  These kernels run the whole time loop in ONE launch with
// every CTA resident for the entire sequence:
//   * the batch rows are independent in the recurrence, so they are split into G groups; a group is
//     served by C CTAs, CTA c owning U hidden units (all four gates) -- its 4U rows of W_hh stay in
//     shared memory for the whole sequence (forward), or its U columns of W_hh (backward);
//   * c / h (forward) and dc / dh (backward) of a cell live in the registers of one owner thread;
//   * per step a CTA only exchanges state with the C-1 other CTAs of ITS group: the new h slice (or
//     the dgate slice) goes through an L2-resident buffer, followed by a release/acquire counter
//     barrier among those C CTAs -- no grid-wide barrier, and each CTA re-reads Ng x H floats, not N x H.
// Up to two directions (forward / reverse of a bidirectional layer) run side by side in one launch.
#include <stdlib.h>
#include "common.cuh"
#include "vmmt_internal.h"
#include "lstm_tc.h"

namespace {

constexpr int LSTM_THREADS = 256;

struct FwdParams {
  VmmtLstmDir d[2];
  float* hbuf[2];            // per direction: [2][N][H] exchange buffer
  unsigned* ctr;             // [ndir * G] arrival counters (zeroed by the host wrapper)
  const int64_t* lengths;    // [N] or null
  int T, N, H, U, C, G, Ng, KS, HS;      // HS: padded smem row stride (floats, multiple of 4)
  int dbg;
};

struct BwdParams {
  VmmtLstmDirBwd d[2];
  unsigned* ctr;
  const int64_t* lengths;
  int T, N, H, U, C, G, Ng, KS, JS;      // JS: padded row stride over the 4H gate axis
};

__device__ __forceinline__ float4 ldcg4(const float* p) {
  return __ldcg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Barrier among the C CTAs of one batch group; `target` = C * (number of barriers so far).
// Called by all threads; writes made by any thread of the CTA before the call are visible to every
// thread of the group's CTAs after it.
__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire(ctr) < target) { }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LSTM_THREADS, 1) lstm_fwd_kernel(const FwdParams P) {
  extern __shared__ __align__(16) float smem[];
  const int per_dir = P.G * P.C;
  const int dir = blockIdx.x / per_dir;
  const int rem = blockIdx.x % per_dir;
  const int grp = rem / P.C, cta = rem % P.C;
  const int u0 = cta * P.U;
  const int n0 = grp * P.Ng;
  const VmmtLstmDir& D = P.d[dir];
  const int T = P.T, N = P.N, H = P.H, U = P.U, HS = P.HS, KS = P.KS;
  const int nn = min(P.Ng, N - n0);                  // batch rows of this group (>= 1)
  const int HS4 = HS / 4;
  float* ws = smem;                                  // [4U][HS]
  float* hs = ws + (size_t)4 * U * HS;               // [Ng][HS]
  float4* part = reinterpret_cast<float4*>(hs + (size_t)P.Ng * HS);   // [KS][Ng][U]
  const int tid = threadIdx.x, nthr = blockDim.x;
  unsigned* ctr = P.ctr + dir * P.G + grp;

  // resident recurrent weights: row j = g*U + u  <-  W_hh[g*H + u0 + u, :]
  for (int e = tid; e < 4 * U * HS; e += nthr) {
    const int j = e / HS, k = e % HS;
    const int g = j / U, unit = u0 + (j % U);
    ws[e] = (unit < H && k < H) ? D.w_hh[((size_t)g * H + unit) * H + k] : 0.0f;
  }
  for (int e = tid; e < P.Ng * HS; e += nthr) {
    const int n = e / HS, k = e % HS;
    hs[e] = (n < nn && k < H && D.h0) ? D.h0[(size_t)(n0 + n) * H + k] : 0.0f;
  }
  // owner thread of (n, u): holds c and h of that cell in registers for the whole sequence
  const bool owner = tid < nn * U;
  const int on = owner ? tid / U : 0, ou = owner ? tid % U : 0, unit = u0 + ou;
  const int gn = n0 + on;                            // global batch row
  const bool live = owner && unit < H;
  float c = 0.f, h = 0.f, bias[4] = {0.f, 0.f, 0.f, 0.f};
  int len = T;
  if (live) {
    if (D.c0) c = D.c0[(size_t)gn * H + unit];
    if (D.h0) h = D.h0[(size_t)gn * H + unit];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const size_t j = (size_t)g * H + unit;
      bias[g] = (D.b_ih ? D.b_ih[j] : 0.f) + (D.b_hh ? D.b_hh[j] : 0.f) +
                (D.rowbias ? D.rowbias[(size_t)gn * 4 * H + j] : 0.f);
    }
    if (P.lengths) len = (int)P.lengths[gn];
  }
  __syncthreads();

  const int ntiles = (nn + 3) / 4;
  const int items = ntiles * U * KS;
  const float4* hs4 = reinterpret_cast<const float4*>(hs);
  const float4* ws4 = reinterpret_cast<const float4*>(ws);
  float* hbuf = P.hbuf[dir];

  for (int s = 0; s < T; ++s) {
    const int t = D.reverse ? T - 1 - s : s;
    float gx[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float* g = D.gx + ((size_t)t * N + gn) * 4 * H + unit;
#pragma unroll
      for (int q = 0; q < 4; ++q) gx[q] = __ldg(g + (size_t)q * H);
    }
    // ---- h_{t-1} W_hh^T for this CTA's 4U gate rows: 4(n) x 4(gates) register tiles, K split KS ways
    for (int w = tid; w < ((P.dbg & 1) ? 0 : items); w += nthr) {
      const int ks = w % KS, r = w / KS, u = r % U, nt = r / U;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[i][g] = 0.f;
      int nrow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) nrow[i] = min(nt * 4 + i, nn - 1);
      for (int k4 = ks; k4 < HS4; k4 += KS) {
        float4 hv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hv[i] = hs4[(size_t)nrow[i] * HS4 + k4];
#pragma unroll
        for (int g = 0; g < 4; ++g) wv[g] = ws4[(size_t)(g * U + u) * HS4 + k4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            acc[i][g] = fmaf(hv[i].x, wv[g].x, acc[i][g]);
            acc[i][g] = fmaf(hv[i].y, wv[g].y, acc[i][g]);
            acc[i][g] = fmaf(hv[i].z, wv[g].z, acc[i][g]);
            acc[i][g] = fmaf(hv[i].w, wv[g].w, acc[i][g]);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = nt * 4 + i;
        if (n < nn) part[((size_t)ks * P.Ng + n) * U + u] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
    __syncthreads();
    // ---- fused cell update (sigmoid/tanh epilogue) by the owner threads
    if (live) {
      float G[4] = {gx[0] + bias[0], gx[1] + bias[1], gx[2] + bias[2], gx[3] + bias[3]};
      for (int ks = 0; ks < KS; ++ks) {
        const float4 p = part[((size_t)ks * P.Ng + on) * U + ou];
        G[0] += p.x; G[1] += p.y; G[2] += p.z; G[3] += p.w;
      }
      const float ig = sigmoidf_(G[0]), fg = sigmoidf_(G[1]), gg = tanhf(G[2]), og = sigmoidf_(G[3]);
      const float cn = fg * c + ig * gg;
      const float hn = og * tanhf(cn);
      const bool m = t < len;
      if (m) { c = cn; h = hn; }
      const size_t row = (size_t)t * N + gn;
      D.out[row * D.out_ld + unit] = m ? hn : 0.f;
      if (D.gates) {
        float* gp = D.gates + row * 4 * H + unit;
        gp[0] = ig; gp[(size_t)H] = fg; gp[(size_t)2 * H] = gg; gp[(size_t)3 * H] = og;
      }
      if (D.cs) D.cs[row * H + unit] = c;
      hbuf[((size_t)(s & 1) * N + gn) * H + unit] = h;
    }
    if (s + 1 < T) {
      if (!(P.dbg & 2)) group_barrier(ctr, (unsigned)(s + 1) * (unsigned)P.C); else __syncthreads();
      const float* hb = hbuf + ((size_t)(s & 1) * N + n0) * H;
      if ((H & 3) == 0) {
        const int H4 = H / 4;
        float4* hsw = reinterpret_cast<float4*>(hs);
        for (int e = tid; e < nn * H4; e += nthr) {
          const int n = e / H4, k4 = e % H4;
          hsw[(size_t)n * HS4 + k4] = ldcg4(hb + (size_t)n * H + 4 * k4);
        }
      } else {
        for (int e = tid; e < nn * H; e += nthr) {
          const int n = e / H, k = e % H;
          hs[(size_t)n * HS + k] = __ldcg(hb + (size_t)n * H + k);
        }
      }
      __syncthreads();
    }
  }
  if (live) {
    if (D.hT) D.hT[(size_t)gn * H + unit] = h;
    if (D.cT) D.cT[(size_t)gn * H + unit] = c;
  }
}

__global__ void __launch_bounds__(LSTM_THREADS, 1) lstm_bwd_kernel(const BwdParams P) {
  extern __shared__ __align__(16) float smem[];
  const int per_dir = P.G * P.C;
  const int dir = blockIdx.x / per_dir;
  const int rem = blockIdx.x % per_dir;
  const int grp = rem / P.C, cta = rem % P.C;
  const int u0 = cta * P.U;
  const int n0 = grp * P.Ng;
  const VmmtLstmDirBwd& D = P.d[dir];
  const int T = P.T, N = P.N, H = P.H, U = P.U, JS = P.JS, KS = P.KS;
  const int nn = min(P.Ng, N - n0);
  const int J = 4 * H, J4 = J / 4, JS4 = JS / 4, UT = U / 4;
  float* wt = smem;                                        // [U][JS]: wt[u][j] = W_hh[j][u0+u]
  float* dgs = wt + (size_t)U * JS;                        // [Ng][JS]: this group's dgate rows of the current step
  float4* part = reinterpret_cast<float4*>(dgs + (size_t)P.Ng * JS);   // [KS][Ng][UT]
  const int tid = threadIdx.x, nthr = blockDim.x;
  unsigned* ctr = P.ctr + dir * P.G + grp;
  for (int e = tid; e < U * JS; e += nthr) {
    const int u = e / JS, j = e % JS;
    wt[e] = (u0 + u < H && j < J) ? D.w_hh[(size_t)j * H + u0 + u] : 0.0f;
  }
  for (int e = tid; e < P.Ng * JS; e += nthr) dgs[e] = 0.0f;
  const bool owner = tid < nn * U;
  const int on = owner ? tid / U : 0, ou = owner ? tid % U : 0, unit = u0 + ou;
  const int gn = n0 + on;
  const bool live = owner && unit < H;
  float dc = 0.f, dhr = 0.f;
  int len = T;
  if (live) {
    if (D.dcT) dc = D.dcT[(size_t)gn * H + unit];
    if (D.dhT) dhr = D.dhT[(size_t)gn * H + unit];
    if (P.lengths) len = (int)P.lengths[gn];
  }
  __syncthreads();
  const int ntiles = (nn + 3) / 4;
  const int items = ntiles * UT * KS;
  const float4* wt4 = reinterpret_cast<const float4*>(wt);
  const float4* dg4 = reinterpret_cast<const float4*>(dgs);

  for (int s = 0; s < T; ++s) {
    const int t = D.reverse ? s : T - 1 - s;            // opposite to the forward order
    float pass = 0.f;
    if (live) {
      const size_t row = (size_t)t * N + gn;
      float dG[4] = {0.f, 0.f, 0.f, 0.f};
      if (t < len) {
        const float* gp = D.gates + row * 4 * H + unit;
        const float ig = gp[0], fg = gp[(size_t)H], gg = gp[(size_t)2 * H], og = gp[(size_t)3 * H];
        const float ct = D.cs[row * H + unit];
        const int tp = D.reverse ? t + 1 : t - 1;
        float cp;
        if (tp >= 0 && tp < T) cp = D.cs[((size_t)tp * N + gn) * H + unit];
        else cp = D.c0 ? D.c0[(size_t)gn * H + unit] : 0.f;
        const float dh = dhr + (D.dout ? D.dout[row * D.dout_ld + unit] : 0.f);
        const float tc = tanhf(ct);
        const float dct = dc + dh * og * (1.f - tc * tc);
        dG[0] = dct * gg * ig * (1.f - ig);
        dG[1] = dct * cp * fg * (1.f - fg);
        dG[2] = dct * ig * (1.f - gg * gg);
        dG[3] = dh * tc * og * (1.f - og);
        dc = dct * fg;
      } else {
        pass = dhr;                                        // frozen state: gradients pass through
      }
      float* dg = D.dgates + row * 4 * H + unit;
      dg[0] = dG[0]; dg[(size_t)H] = dG[1]; dg[(size_t)2 * H] = dG[2]; dg[(size_t)3 * H] = dG[3];
    }
    if (s + 1 == T && D.dh0 == nullptr) break;            // the last dh_prev is only needed for dh0
    group_barrier(ctr, (unsigned)(s + 1) * (unsigned)P.C);
    // ---- stage this group's dgate rows (written by the C CTAs of the group) into shared memory
    {
      const float* dgt = D.dgates + ((size_t)t * N + n0) * J;
      float4* dsw = reinterpret_cast<float4*>(dgs);
      for (int e = tid; e < nn * J4; e += nthr) {
        const int n = e / J4, j4 = e % J4;
        dsw[(size_t)n * JS4 + j4] = ldcg4(dgt + (size_t)n * J + 4 * j4);
      }
    }
    __syncthreads();
    // ---- dh_{prev}[n, u] = sum_j dG_t[n, j] W_hh[j, u]: 4(n) x 4(u) tiles, j split KS ways
    for (int w = tid; w < items; w += nthr) {
      const int ks = w % KS, r = w / KS, ut = r % UT, nt = r / UT;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
      int nrow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) nrow[i] = min(nt * 4 + i, nn - 1);
      for (int j4 = ks; j4 < J4; j4 += KS) {
        float4 gv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) gv[i] = dg4[(size_t)nrow[i] * JS4 + j4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wv[q] = wt4[(size_t)(ut * 4 + q) * JS4 + j4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[i][q] = fmaf(gv[i].x, wv[q].x, acc[i][q]);
            acc[i][q] = fmaf(gv[i].y, wv[q].y, acc[i][q]);
            acc[i][q] = fmaf(gv[i].z, wv[q].z, acc[i][q]);
            acc[i][q] = fmaf(gv[i].w, wv[q].w, acc[i][q]);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = nt * 4 + i;
        if (n < nn) part[((size_t)ks * P.Ng + n) * UT + ut] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
    __syncthreads();
    if (live) {
      float sum = pass;
      const float* pf = reinterpret_cast<const float*>(part);
      for (int ks = 0; ks < KS; ++ks) sum += pf[(((size_t)ks * P.Ng + on) * UT) * 4 + ou];
      dhr = sum;
    }
  }
  if (live) {
    if (D.dh0) D.dh0[(size_t)gn * H + unit] = dhr;
    if (D.dc0) D.dc0[(size_t)gn * H + unit] = dc;
  }
}

struct Plan {
  int G, C, U, Ng, KS, stride;
  size_t smem;
};

// Largest number of batch groups whose per-CTA working set fits shared memory: fewer CTAs per group
// means fewer barrier participants and less exchange traffic per step.
int make_plan(bool bwd, int ndir, int N, int H, Plan* best) {
  const int budget = vmmt_num_sms() / ndir;
  const size_t cap = 227 * 1024;
  bool found = false;
  const int gforce = getenv("VMMT_LSTM_G") ? atoi(getenv("VMMT_LSTM_G")) : 0;
  for (int G = 1; G <= 8 && G <= N; ++G) {
    if (!bwd && gforce && G != gforce) continue;
    Plan p;
    p.Ng = ceil_div(N, G);
    p.G = ceil_div(N, p.Ng);
    if (p.G != G) continue;
    const int cmax = budget / G;
    if (cmax < 1) break;
    int U = ceil_div(H, cmax);
    if (bwd) U = ((U + 3) / 4) * 4;
    p.U = U;
    p.C = ceil_div(H, U);
    if (p.Ng * U > LSTM_THREADS) continue;            // one owner thread per (row, unit)
    const int ntiles = (p.Ng + 3) / 4;
    if (!bwd) {
      p.stride = ((H + 3) / 4) * 4 + 4;
      p.KS = max(1, min(16, LSTM_THREADS / (ntiles * U)));
      p.smem = ((size_t)(4 * U + p.Ng) * p.stride) * 4 + (size_t)p.KS * p.Ng * U * 16;
    } else {
      p.stride = 4 * H + 4;
      p.KS = max(1, min(32, LSTM_THREADS / (ntiles * (U / 4))));
      p.smem = ((size_t)(U + p.Ng) * p.stride) * 4 + (size_t)p.KS * p.Ng * (U / 4) * 16;
    }
    if (p.smem > cap) continue;
    *best = p;                                        // keep the largest feasible G
    found = true;
  }
  return found ? VMMT_OK : VMMT_EINVAL;
}

template <typename K>
int resident_launch(K kernel, int grid, size_t smem, void* params, cudaStream_t s, const char* what) {
  // attributes / occupancy are checked once per (kernel, device, configuration), not per launch
  static size_t smem_set[64] = {0};
  static int ok_grid[64] = {0};
  static size_t ok_smem[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem > smem_set[dev]) {
    VMMT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[dev] = smem;
  }
  if (!(grid <= ok_grid[dev] && smem <= ok_smem[dev])) {
    int per_sm = 0;
    VMMT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, LSTM_THREADS, smem));
    if (per_sm * vmmt_num_sms() < grid) {
      vmmt_set_error("%s: grid of %d CTAs is not co-resident (%d/SM)", what, grid, per_sm);
      return VMMT_ELAUNCH;
    }
    ok_grid[dev] = grid; ok_smem[dev] = smem;
  }
  // cooperative launch = the driver guarantees co-residency of the whole grid (the inter-CTA
  // barriers above spin); it is stream-ordered and graph-capturable like any other launch
  void* args[] = {params};
  VMMT_CUDA(cudaLaunchCooperativeKernel((void*)kernel, dim3(grid), dim3(LSTM_THREADS), args, smem, s));
  return vmmt_check_launch(what);
}

constexpr size_t CTR_BYTES = 256;    // counters live at the head of the workspace

}  // namespace

extern "C" size_t vmmt_lstm_workspace_bytes(int ndir, int N, int H) {
  return CTR_BYTES + vmmt_lstm_step_workspace_floats(ndir, N, H) * sizeof(float);
}

extern "C" int vmmt_lstm_seq_supported(int ndir, int N, int H) {
  (void)N; (void)H;                       // any shape: cluster / persistent kernels, else the step-wise GEMM path
  return (ndir >= 1 && ndir <= 2) ? 1 : 0;
}

extern "C" int vmmt_lstm_seq_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T,
                                 int N, int H, int flags, int cluster_budget, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  VMMT_REQUIRE(ndir == 1 || ndir == 2, "lstm_seq_fwd: ndir must be 1 or 2 (got %d)", ndir);
  VMMT_REQUIRE(T > 0 && N > 0 && H > 0, "lstm_seq_fwd: bad dims T=%d N=%d H=%d", T, N, H);
  // tensor-core cluster path (default); VMMT_F_EXACT selects the exact-fp32 SIMT kernels
  if (!(flags & VMMT_F_EXACT) && !getenv("VMMT_LSTM_SIMT") && !getenv("VMMT_LSTM_STEPWISE") &&
      vmmt_lstm_tc_supported(ndir, N, H))
    return vmmt_lstm_tc_fwd(dirs, ndir, lengths, T, N, H, cluster_budget, (cudaStream_t)stream);
  if (workspace_bytes < vmmt_lstm_workspace_bytes(ndir, N, H)) {
    vmmt_set_error("lstm_seq_fwd: workspace too small");
    return VMMT_EWORKSPACE;
  }
  Plan p;
  // large batch / hidden size: one GEMM + one cell kernel per step (lstm_step.cu)
  if (getenv("VMMT_LSTM_STEPWISE") || make_plan(false, ndir, N, H, &p) != VMMT_OK)
    return vmmt_lstm_step_fwd(dirs, ndir, lengths, T, N, H, flags,
                              reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + CTR_BYTES), (cudaStream_t)stream);
  cudaStream_t s = (cudaStream_t)stream;
  VMMT_CUDA(cudaMemsetAsync(workspace, 0, CTR_BYTES, s));
  FwdParams P;
  float* hb = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + CTR_BYTES);
  for (int d = 0; d < ndir; ++d) {
    P.d[d] = dirs[d];
    P.hbuf[d] = hb + (size_t)d * 2 * N * H;
  }
  if (ndir == 1) { P.d[1] = dirs[0]; P.hbuf[1] = P.hbuf[0]; }
  P.ctr = reinterpret_cast<unsigned*>(workspace);
  P.lengths = lengths;
  P.T = T; P.N = N; P.H = H; P.U = p.U; P.C = p.C; P.G = p.G; P.Ng = p.Ng; P.KS = p.KS; P.HS = p.stride;
  P.dbg = getenv("VMMT_LSTM_DBG") ? atoi(getenv("VMMT_LSTM_DBG")) : 0;
  return resident_launch(lstm_fwd_kernel, ndir * p.G * p.C, p.smem, &P, s, "lstm_fwd_kernel");
}

static bool bwd_takes_tc_path(int ndir, int N, int H, int flags) {
  return !(flags & VMMT_F_EXACT) && !getenv("VMMT_LSTM_SIMT") && !getenv("VMMT_LSTM_SIMT_BWD") &&
         !getenv("VMMT_LSTM_STEPWISE") && vmmt_lstm_tc_supported(ndir, N, H);
}

extern "C" int vmmt_lstm_seq_bwd_fuses_bias(int ndir, int N, int H, int flags) {
  return (bwd_takes_tc_path(ndir, N, H, flags) && !getenv("VMMT_LSTM_NO_BIAS_FUSE")) ? 1 : 0;
}

extern "C" int vmmt_lstm_seq_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths,
                                 int T, int N, int H, int flags, int cluster_budget, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  VMMT_REQUIRE(ndir == 1 || ndir == 2, "lstm_seq_bwd: ndir must be 1 or 2 (got %d)", ndir);
  if (bwd_takes_tc_path(ndir, N, H, flags))
    return vmmt_lstm_tc_bwd(dirs, ndir, lengths, T, N, H, cluster_budget, (cudaStream_t)stream);
  if (workspace_bytes < vmmt_lstm_workspace_bytes(ndir, N, H)) {
    vmmt_set_error("lstm_seq_bwd: workspace too small");
    return VMMT_EWORKSPACE;
  }
  Plan p;
  if (getenv("VMMT_LSTM_STEPWISE") || make_plan(true, ndir, N, H, &p) != VMMT_OK)
    return vmmt_lstm_step_bwd(dirs, ndir, lengths, T, N, H, flags,
                              reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + CTR_BYTES), (cudaStream_t)stream);
  cudaStream_t s = (cudaStream_t)stream;
  VMMT_CUDA(cudaMemsetAsync(workspace, 0, CTR_BYTES, s));
  BwdParams P;
  for (int d = 0; d < ndir; ++d) P.d[d] = dirs[d];
  if (ndir == 1) P.d[1] = dirs[0];
  P.ctr = reinterpret_cast<unsigned*>(workspace);
  P.lengths = lengths;
  P.T = T; P.N = N; P.H = H; P.U = p.U; P.C = p.C; P.G = p.G; P.Ng = p.Ng; P.KS = p.KS; P.JS = p.stride;
  return resident_launch(lstm_bwd_kernel, ndir * p.G * p.C, p.smem, &P, s, "lstm_bwd_kernel");
}
