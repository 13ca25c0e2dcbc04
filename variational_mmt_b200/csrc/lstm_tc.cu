// Cluster-resident LSTM recurrence on tcgen05 tensor cores (forward and BPTT), H <= 512.
//
// Reference semantics: torch nn.LSTM (onmt/Models.py:124-149, 892-893; onmt/VI_Model1.py:106), gate
// order i,f,g,o.  gx = x W_ih^T for all timesteps comes from one batched tensor-core GEMM (vmmt_gemm).
//
// The batch rows of a recurrence are independent, so they are split into G groups of <= 16 rows.  One
// thread-block CLUSTER of C = ceil(H/32) CTAs (<= 16, one per SM) serves one (direction, group) for the whole
// sequence with no global-memory round trip per step:
//   * CTA c owns 32 hidden units = 128 gate rows [i(32) f(32) g(32) o(32)] of W_hh.  Its 128 x Kp slice stays
//     resident in shared memory as fp16 (K-major, 128B swizzle).  fp16 carries the same 11 significand bits as
//     TF32, which is what a tensor-core "fp32" GEMM rounds its operands to; accumulation is fp32 in TMEM;
//     c / h state stays fp32 in registers, h is rounded to fp16 only as next step's MMA operand.
//   * forward step : D[128 gate rows, 16 batch cols] = W_slice[128,Kp] * h_{t-1}^T  (32 tcgen05.mma M128 N16 K16)
//                    -> tcgen05.ld -> gates regrouped through shared memory -> sigmoid/tanh cell update fused in
//                    the epilogue -> the CTA's new h slice is broadcast as 16-byte DSMEM stores straight into
//                    the B-operand tile of all C CTAs (double buffered) -> one cluster barrier.
//   * backward step: the SAME resident slice read MN-major gives A = W_slice^T: each CTA contracts its own 128
//                    gate rows, partial dh[512 units, batch] = W_slice^T * dG_t^T (dG enters as fp16 hi + lo
//                    columns of dG * 2^12: 22 significand bits, no gradient magnitude is lost), the partials are reduce-scattered through
//                    DSMEM to the unit owners and summed in a fixed order (deterministic) after one cluster
//                    barrier.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <stdio.h>
#include "common.cuh"
#include "vmmt_internal.h"
#include "lstm_tc.h"

namespace {

constexpr int UC = 32;                 // hidden units per CTA
constexpr int NP = 16;                 // batch columns per cluster (MMA N)
constexpr int THREADS = 160;           // warps 0-3: epilogue / cell owners, warp 4: MMA issuer
constexpr int A_BLOCK = 128 * 128;     // one k-block of the resident slice: 128 rows x 64 fp16 = 16 KB
constexpr int B_BLOCK = NP * 128;      // one k-block of the h tile: 16 rows x 64 fp16 = 2 KB
constexpr int NCH = 4;                 // independent accumulator chains (a dependent tcgen05.mma chain is latency bound at N = 16)
constexpr float DG_SCALE = 4096.0f;    // backward: dG enters the MMA as fp16 hi + lo of dG * 2^12

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// fast transcendental forms for the tensor-core path (their ~1e-6 error is far below the operand rounding)
__device__ __forceinline__ float fsigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// shared-memory matrix descriptor, 128-byte swizzle, sm_100 version bits
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// byte offset of the 16-byte chunk holding elements [8*chunk, 8*chunk+8) of row `row` in a [rows][64 x 2B]
// 128B-swizzled block
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

struct TcFwdParams {
  VmmtLstmDir d[2];
  const int64_t* lengths;
  int T, N, H, C, G, Ng, Kp;
  long long* trace;
};
struct TcBwdParams {
  VmmtLstmDirBwd d[2];
  const int64_t* lengths;
  int T, N, H, C, G, Ng, Kp;
};

// resident slice: row r = gate*32 + unit  <-  W_hh[gate*H + u0 + unit, 0:H] as fp16, zero padded
__device__ __forceinline__ void load_w_slice(uint8_t* A, const float* __restrict__ w_hh, int H, int Kp, int u0) {
  const int chunks_per_row = Kp / 8;
  for (int e = threadIdx.x; e < 128 * chunks_per_row; e += blockDim.x) {
    const int r = e / chunks_per_row, ci = e % chunks_per_row;
    const int gate = r >> 5, unit = u0 + (r & 31);
    __align__(16) __half hv[8];
    const int k0 = ci * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i;
      const float w = (unit < H && k < H) ? __ldg(w_hh + ((size_t)gate * H + unit) * H + k) : 0.0f;
      hv[i] = __float2half_rn(w);
    }
    *reinterpret_cast<uint4*>(A + (size_t)(ci >> 3) * A_BLOCK + sw_off(r, ci & 7)) = *reinterpret_cast<uint4*>(hv);
  }
}

// TS mode: the slice lives in TMEM as the MMA's A operand (lane = row, one 32-bit column = two consecutive k).
// forward: row r = gate*32 + unit (warp = gate, lane = unit), columns k/2 for k in [0, Kp)
__device__ __forceinline__ void load_w_tmem_fwd(uint32_t tmem_a, const float* __restrict__ w_hh, int H, int Kp, int u0,
                                                int warp, int lane) {
  const int unit = u0 + lane;
  const float* row = w_hh + ((size_t)warp * H + min(unit, H - 1)) * H;
  const bool vec = (H & 3) == 0;
  for (int c = 0; c < Kp / 32; ++c) {
    float f[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = c * 32 + 4 * i;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (unit < H) {
        if (vec && k + 3 < H) q = __ldg(reinterpret_cast<const float4*>(row + k));
        else {
          if (k < H) q.x = __ldg(row + k);
          if (k + 1 < H) q.y = __ldg(row + k + 1);
          if (k + 2 < H) q.z = __ldg(row + k + 2);
          if (k + 3 < H) q.w = __ldg(row + k + 3);
        }
      }
      f[4 * i] = q.x; f[4 * i + 1] = q.y; f[4 * i + 2] = q.z; f[4 * i + 3] = q.w;
    }
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack_h2(f[2 * i], f[2 * i + 1]);
    tmem_st_x16(tmem_a + ((uint32_t)(32 * warp) << 16) + (uint32_t)(c * 16), pk);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// backward: A = W_slice^T: tile mt, row = unit 128 mt + 32 warp + lane, columns j/2 for this CTA's 128 gate rows j
__device__ __forceinline__ void load_w_tmem_bwd(uint32_t tmem_a, const float* __restrict__ w_hh, int H, int u0, int n_mt,
                                                int warp, int lane) {
  for (int mt = 0; mt < n_mt; ++mt) {
    const int u = 128 * mt + 32 * warp + lane;
    for (int gate = 0; gate < 4; ++gate) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int ul0 = u0 + 2 * i, ul1 = ul0 + 1;
        const float a = (u < H && ul0 < H) ? __ldg(w_hh + ((size_t)gate * H + ul0) * H + u) : 0.f;
        const float b = (u < H && ul1 < H) ? __ldg(w_hh + ((size_t)gate * H + ul1) * H + u) : 0.f;
        pk[i] = pack_h2(a, b);
      }
      tmem_st_x16(tmem_a + ((uint32_t)(32 * warp) << 16) + (uint32_t)(mt * 64 + gate * 16), pk);
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ forward
template <bool TS>
__global__ void __launch_bounds__(THREADS, 1) lstm_tc_fwd_kernel(const TcFwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int Kp = P.Kp, nkb = Kp / 64;
  const uint32_t a_bytes = TS ? 0u : (uint32_t)nkb * A_BLOCK;   // SS mode: resident slice in shared memory
  uint8_t* A = sm;                                          // [nkb][128][128 B]
  uint8_t* Bt = A + a_bytes;                                // [2][nkb][NP][128 B]
  float* gsm = reinterpret_cast<float*>(Bt + (size_t)2 * nkb * B_BLOCK);   // [4][NP][32]
  __half* hstage = reinterpret_cast<__half*>(gsm + 4 * NP * 32);          // [NP][32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(hstage + NP * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int C = P.C;
  const int cl = blockIdx.x / C;                            // cluster index = (dir, group)
  const int crank = (int)cluster_rank();
  const int dir = cl / P.G, grp = cl % P.G;
  const VmmtLstmDir& D = P.d[dir];
  const int T = P.T, N = P.N, H = P.H;
  const int n0 = grp * P.Ng;
  const int nn = min(P.Ng, N - n0);
  const int u0 = crank * UC;

  if (P.trace && blockIdx.x == 0 && tid == 0) { P.trace[8] = clock64(); unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); P.trace[11] = (long long)g; }
  if (!TS) load_w_slice(A, D.w_hh, H, Kp, u0);
  // h_{-1}: every CTA fills its own operand tile (buffer 0) from h0; rows >= nn and k >= H are zero
  for (int e = tid; e < 2 * nkb * NP * 8; e += blockDim.x) {
    const int bufi = e / (nkb * NP * 8), rem = e % (nkb * NP * 8);
    const int kb = rem / (NP * 8), row = (rem / 8) % NP, ch = rem % 8;
    __align__(16) __half hv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kb * 64 + ch * 8 + i;
      const float v = (bufi == 0 && D.h0 && row < nn && k < H) ? D.h0[(size_t)(n0 + row) * H + k] : 0.0f;
      hv[i] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(Bt + (size_t)bufi * nkb * B_BLOCK + (size_t)kb * B_BLOCK + sw_off(row, ch)) =
        *reinterpret_cast<uint4*>(hv);
  }
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    if (TS) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy smem writes -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base + (TS ? 256u : 0u);     // TS: columns [0,256) hold the weight slice
  if (TS && warp < 4) {
    load_w_tmem_fwd(tmem_base, D.w_hh, H, Kp, u0, warp, lane);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  // cell ownership: thread (warp q < 4, lane) owns cells (n = q + 4 r, unit = lane), r = 0..3
  float c[4] = {0.f, 0.f, 0.f, 0.f}, h[4] = {0.f, 0.f, 0.f, 0.f}, bias[4][4];
  int len[4];
  const int unit = u0 + lane;
  const bool ulive = warp < 4 && unit < H;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = warp + 4 * r;
    len[r] = T;
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[r][g] = 0.f;
    if (ulive && n < nn) {
      const int gn = n0 + n;
      if (D.c0) c[r] = D.c0[(size_t)gn * H + unit];
      if (D.h0) h[r] = D.h0[(size_t)gn * H + unit];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const size_t j = (size_t)g * H + unit;
        bias[r][g] = (D.b_ih ? D.b_ih[j] : 0.f) + (D.b_hh ? D.b_hh[j] : 0.f) +
                     (D.rowbias ? D.rowbias[(size_t)gn * 4 * H + j] : 0.f);
      }
      if (P.lengths) len[r] = (int)P.lengths[gn];
    }
  }
  // instruction descriptor: D f32, A/B fp16 K-major, N = 16, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  cluster_sync_all();          // every CTA of the cluster is initialised before any remote store lands
  if (P.trace && blockIdx.x == 0 && tid == 0) P.trace[9] = clock64();

  for (int s = 0; s < T; ++s) {
    const int t = D.reverse ? T - 1 - s : s;
    const int buf = s & 1;
    const bool tr = P.trace != nullptr && blockIdx.x == 0 && s == 5;
    if (tr && tid == 0) P.trace[0] = clock64();
    if (warp == 4) {
      if (lane == 0) {
        if (tr) P.trace[13] = clock64();
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tr) P.trace[14] = clock64();
        const uint32_t a0 = base, b0 = base + a_bytes + buf * nkb * B_BLOCK;
        for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {                     // k-step j of block kb accumulates into chain j
            const uint64_t bd = make_desc(b0 + kb * B_BLOCK + j * 32, 16, 1024);
            if (TS) {
              umma_f16_ts(tmem_d + j * NP, tmem_base + (uint32_t)(kb * 32 + j * 8), bd, idesc, kb > 0 ? 1u : 0u);
            } else {
              const uint64_t ad = make_desc(a0 + kb * A_BLOCK + j * 32, 16, 1024);
              umma_f16(tmem_d + j * NP, ad, bd, idesc, kb > 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(smem_u32(bar));
        if (tr) P.trace[1] = clock64();
      }
      __syncwarp();
    } else {
      // input-projection terms of this step (independent of the recurrence: issued before the wait)
      float gx[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = warp + 4 * r;
#pragma unroll
        for (int g = 0; g < 4; ++g) gx[r][g] = 0.f;
        if (ulive && n < nn) {
          const float* gp = D.gx + ((size_t)t * N + n0 + n) * 4 * H + unit;
#pragma unroll
          for (int g = 0; g < 4; ++g) gx[r][g] = __ldg(gp + (size_t)g * H);
        }
      }
      if (tr && tid == 0) P.trace[2] = clock64();
      mbar_wait(smem_u32(bar), (uint32_t)(s & 1));
      if (tr && tid == 0) P.trace[3] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float acc[NP];
      {
        uint32_t v[32];                                     // chains 0,1 then 2,3: 4 x 16 columns
        tmem_ld_x32(tmem_d + ((uint32_t)(32 * warp) << 16), v);
#pragma unroll
        for (int n = 0; n < NP; ++n) acc[n] = __uint_as_float(v[n]) + __uint_as_float(v[NP + n]);
        tmem_ld_x32(tmem_d + ((uint32_t)(32 * warp) << 16) + 2 * NP, v);
#pragma unroll
        for (int n = 0; n < NP; ++n) acc[n] += __uint_as_float(v[n]) + __uint_as_float(v[NP + n]);
      }
#pragma unroll
      for (int n = 0; n < NP; ++n) gsm[(warp * NP + n) * 32 + lane] = acc[n];   // gate `warp`, unit `lane`
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tr && tid == 0) P.trace[4] = clock64();
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = warp + 4 * r;
        float hn16 = 0.f;
        if (ulive && n < nn) {
          const float Gi = gsm[(0 * NP + n) * 32 + lane] + gx[r][0] + bias[r][0];
          const float Gf = gsm[(1 * NP + n) * 32 + lane] + gx[r][1] + bias[r][1];
          const float Gg = gsm[(2 * NP + n) * 32 + lane] + gx[r][2] + bias[r][2];
          const float Go = gsm[(3 * NP + n) * 32 + lane] + gx[r][3] + bias[r][3];
          const float ig = fsigmoid(Gi), fg = fsigmoid(Gf), gg = ftanh(Gg), og = fsigmoid(Go);
          const float cn = fg * c[r] + ig * gg;
          const float hn = og * ftanh(cn);
          const bool m = t < len[r];
          if (m) { c[r] = cn; h[r] = hn; }
          const size_t row = (size_t)t * N + n0 + n;
          D.out[row * D.out_ld + unit] = m ? hn : 0.f;
          if (D.gates) {
            float* gp = D.gates + row * 4 * H + unit;
            gp[0] = ig; gp[(size_t)H] = fg; gp[(size_t)2 * H] = gg; gp[(size_t)3 * H] = og;
          }
          if (D.cs) D.cs[row * H + unit] = c[r];
          hn16 = h[r];
        }
        hstage[n * 32 + lane] = __float2half_rn(hn16);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tr && tid == 0) P.trace[5] = clock64();
      if (s + 1 < T) {
        // broadcast this CTA's h slice (nn rows x 32 units = 4 chunks of 16 B per row) into the operand tile
        // (other buffer) of every CTA of the cluster, own CTA included
        const int kb = u0 >> 6, ch0 = (u0 & 63) >> 3;
        const uint32_t dst_local = base + a_bytes + (buf ^ 1) * nkb * B_BLOCK + kb * B_BLOCK;
        // thread -> (row n, chunk cj) = tid & 63, destinations (tid >> 6) + 2 i
        const int n = (tid & 63) >> 2, cj = tid & 3;
        if (n < nn) {
          const uint4 val = *reinterpret_cast<const uint4*>(hstage + n * 32 + cj * 8);
          const uint32_t off = dst_local + sw_off(n, ch0 + cj);
          for (int dstc = tid >> 6; dstc < C; dstc += 2) st_cluster_v4(mapa(off, (uint32_t)dstc), val);
        }
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (tr && tid == 0) P.trace[6] = clock64();
    }
    if (s + 1 < T) cluster_sync_all();
    if (tr && tid == 0) P.trace[7] = clock64();
  }
  if (P.trace && blockIdx.x == 0 && tid == 0) { P.trace[10] = clock64(); unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); P.trace[12] = (long long)g; }
  if (warp < 4) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = warp + 4 * r;
      if (ulive && n < nn) {
        if (D.hT) D.hT[(size_t)(n0 + n) * H + unit] = h[r];
        if (D.cT) D.cT[(size_t)(n0 + n) * H + unit] = c[r];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    if (TS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64));
  }
  cluster_sync_all();          // no CTA exits while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------ backward
template <bool TS>
__global__ void __launch_bounds__(THREADS, 1) lstm_tc_bwd_kernel(const TcBwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int Kp = P.Kp, nkb = Kp / 64;
  const uint32_t a_bytes = TS ? 0u : (uint32_t)nkb * A_BLOCK;
  uint8_t* A = sm;                                          // [nkb][128][128 B]  (same resident slice as forward)
  uint8_t* Bt = A + a_bytes;                                // [2 k-blocks][2*NP rows][128 B]: dG hi rows 0..15, lo rows 16..31
  float* recv = reinterpret_cast<float*>(Bt + 2 * (2 * NP) * 128);       // [2][16 src][NP][32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(recv + 2 * 16 * NP * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int C = P.C;
  const int cl = blockIdx.x / C;
  const int crank = (int)cluster_rank();
  const int dir = cl / P.G, grp = cl % P.G;
  const VmmtLstmDirBwd& D = P.d[dir];
  const int T = P.T, N = P.N, H = P.H;
  const int n0 = grp * P.Ng;
  const int nn = min(P.Ng, N - n0);
  const int u0 = crank * UC;

  if (!TS) load_w_slice(A, D.w_hh, H, Kp, u0);
  for (int e = tid; e < 2 * (2 * NP) * 128 / 16; e += blockDim.x) reinterpret_cast<uint4*>(Bt)[e] = make_uint4(0, 0, 0, 0);
  for (int e = tid; e < 2 * 16 * NP * 32; e += blockDim.x) recv[e] = 0.f;
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    if (TS) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base + (TS ? 256u : 0u);
  const int n_mt = (P.C * UC + 127) / 128;                  // 128-unit output tiles that hold real units
  if (TS && warp < 4) {
    load_w_tmem_bwd(tmem_base, D.w_hh, H, u0, n_mt, warp, lane);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  float dc[4] = {0.f, 0.f, 0.f, 0.f}, dhr[4] = {0.f, 0.f, 0.f, 0.f};
  int len[4];
  const int unit = u0 + lane;
  const bool ulive = warp < 4 && unit < H;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = warp + 4 * r;
    len[r] = T;
    if (ulive && n < nn) {
      const int gn = n0 + n;
      if (D.dcT) dc[r] = D.dcT[(size_t)gn * H + unit];
      if (D.dhT) dhr[r] = D.dhT[(size_t)gn * H + unit];
      if (P.lengths) len[r] = (int)P.lengths[gn];
    }
  }
  // D f32, A fp16 MN-major (W_slice^T), B fp16 K-major, N = 32 (hi | lo), M = 128
  const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | ((uint32_t)((2 * NP) >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_ts = idesc & ~(1u << 15);            // A from TMEM has no major bit
  cluster_sync_all();

  for (int s = 0; s < T; ++s) {
    const int t = D.reverse ? s : T - 1 - s;
    const int buf = s & 1;
    const bool last = (s + 1 == T);
    if (warp < 4) {
      // ---- elementwise BPTT of this CTA's cells -> dG_t (global, and bf16 hi/lo operand rows n / 16+n)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = warp + 4 * r;
        float dG[4] = {0.f, 0.f, 0.f, 0.f};
        if (ulive && n < nn) {
          const int gn = n0 + n;
          const size_t row = (size_t)t * N + gn;
          if (t < len[r]) {
            const float* gp = D.gates + row * 4 * H + unit;
            const float ig = gp[0], fg = gp[(size_t)H], gg = gp[(size_t)2 * H], og = gp[(size_t)3 * H];
            const float ct = D.cs[row * H + unit];
            const int tp = D.reverse ? t + 1 : t - 1;
            float cp;
            if (tp >= 0 && tp < T) cp = D.cs[((size_t)tp * N + gn) * H + unit];
            else cp = D.c0 ? D.c0[(size_t)gn * H + unit] : 0.f;
            const float dh = dhr[r] + (D.dout ? D.dout[row * D.dout_ld + unit] : 0.f);
            const float tc = ftanh(ct);
            const float dct = dc[r] + dh * og * (1.f - tc * tc);
            dG[0] = dct * gg * ig * (1.f - ig);
            dG[1] = dct * cp * fg * (1.f - fg);
            dG[2] = dct * ig * (1.f - gg * gg);
            dG[3] = dh * tc * og * (1.f - og);
            dc[r] = dct * fg;
            dhr[r] = 0.f;                                   // consumed; the next value comes from the reduce below
          }                                                 // else: frozen state, dhr passes through unchanged
          float* dg = D.dgates + row * 4 * H + unit;
          dg[0] = dG[0]; dg[(size_t)H] = dG[1]; dg[(size_t)2 * H] = dG[2]; dg[(size_t)3 * H] = dG[3];
        }
        // operand tile: k = gate*32 + lane (this CTA's 128 gate rows), row n (hi) and NP + n (lo)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          // fp16 hi + lo of dG * 2^12: 22 significand bits, magnitudes from 1.5e-11 to 16 (clamped beyond)
          const float xs = fminf(fmaxf(dG[g] * DG_SCALE, -60000.f), 60000.f);
          const __half hi = __float2half_rn(xs);
          const __half lo = __float2half_rn(xs - __half2float(hi));
          const int k = g * 32 + lane, kb = k >> 6, ch = (k & 63) >> 3, el = k & 7;
          *reinterpret_cast<__half*>(Bt + kb * (2 * NP * 128) + sw_off(n, ch) + el * 2) = hi;
          *reinterpret_cast<__half*>(Bt + kb * (2 * NP * 128) + sw_off(NP + n, ch) + el * 2) = lo;
        }
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (last && D.dh0 == nullptr) break;                    // the last dh_prev only feeds dh0
    __syncthreads();
    if (warp == 4) {
      if (lane == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b0 = base + a_bytes;
        for (int mt = 0; mt < n_mt; ++mt) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                  // K = 128 gate rows = 8 steps of 16
            const uint64_t bd = make_desc(b0 + (ks >> 2) * (2 * NP * 128) + (ks & 3) * 32, 16, 1024);
            if (TS) {
              umma_f16_ts(tmem_d + mt * 32, tmem_base + (uint32_t)(mt * 64 + ks * 8), bd, idesc_ts, ks > 0 ? 1u : 0u);
            } else {
              // A^T: M = units [128 mt, +128) = unit blocks 2 mt, 2 mt + 1 (LBO apart), K rows 16 ks.. (SBO = 8 rows)
              const uint64_t ad = make_desc(base + (2 * mt) * A_BLOCK + ks * 2048, A_BLOCK, 1024);
              umma_f16(tmem_d + mt * 32, ad, bd, idesc, ks > 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(smem_u32(bar));
      }
      __syncwarp();
    } else {
      mbar_wait(smem_u32(bar), (uint32_t)(s & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // partial dh[unit = 128 mt + 32 warp + lane, n] -> owner CTA 4 mt + warp, slot recv[buf][src = crank][n][lane]
      for (int mt = 0; mt < n_mt; ++mt) {
        uint32_t v[32];
        tmem_ld_x32(tmem_d + ((uint32_t)(32 * warp) << 16) + (uint32_t)(mt * 32), v);
        const int owner = 4 * mt + warp;
        if (owner < C) {
          const uint32_t dst = mapa(smem_u32(recv) + (uint32_t)(((buf * 16 + crank) * NP) * 32 + lane) * 4u, (uint32_t)owner);
#pragma unroll
          for (int n = 0; n < NP; ++n)
            if (n < nn) st_cluster_f32(dst + (uint32_t)n * 128u, (__uint_as_float(v[n]) + __uint_as_float(v[NP + n])) * (1.0f / DG_SCALE));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    cluster_sync_all();
    if (warp < 4) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = warp + 4 * r;
        if (ulive && n < nn) {
          float sum = dhr[r];                               // non-zero only for frozen (masked) cells
          for (int src = 0; src < C; ++src) sum += recv[((buf * 16 + src) * NP + n) * 32 + lane];
          dhr[r] = sum;
        }
      }
    }
  }
  if (warp < 4) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = warp + 4 * r;
      if (ulive && n < nn) {
        if (D.dh0) D.dh0[(size_t)(n0 + n) * H + unit] = dhr[r];
        if (D.dc0) D.dc0[(size_t)(n0 + n) * H + unit] = dc[r];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    if (TS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
  }
  cluster_sync_all();
}

struct TcPlan { int C, G, Ng, Kp; size_t smem_f, smem_b; };

// weight slice in TMEM (default) or in shared memory (VMMT_LSTM_SS=1)
bool tc_ts_mode() { static const bool ts = getenv("VMMT_LSTM_SS") == nullptr; return ts; }

bool tc_plan(int ndir, int N, int H, TcPlan* p) {
  if (H < 32 || H > 512 || N < 1) return false;
  p->C = ceil_div(H, UC);
  if (p->C > 16) return false;
  p->Kp = ceil_div(H, 64) * 64;
  // as many groups as there are cluster slots (one CTA per SM), at most 16 rows per group
  const int slots = max(1, vmmt_num_sms() / (p->C * ndir));
  int G = min(min(slots, 8), N);
  int Ng = ceil_div(N, G);
  if (Ng > NP) { Ng = NP; }
  G = ceil_div(N, Ng);
  p->G = G; p->Ng = Ng;
  const int nkb = p->Kp / 64;
  const size_t a_bytes = tc_ts_mode() ? 0 : (size_t)nkb * A_BLOCK;
  p->smem_f = 1024 + a_bytes + (size_t)2 * nkb * B_BLOCK + 4 * NP * 32 * 4 + NP * 32 * 2 + 64;
  p->smem_b = 1024 + a_bytes + 2 * (2 * NP) * 128 + (size_t)2 * 16 * NP * 32 * 4 + 64;
  return p->smem_f <= 227 * 1024 && p->smem_b <= 227 * 1024;
}

template <typename K, typename PT>
int cluster_launch(K kernel, const PT& params, int grid, int C, size_t smem, cudaStream_t s, const char* what) {
  // function attributes are sticky: set them once per (kernel, device) and only raise the smem limit when needed
  static size_t smem_set[64] = {0};
  static bool np_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem > smem_set[dev]) {
    VMMT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[dev] = smem;
  }
  if (C > 8 && !np_set[dev]) {
    VMMT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    np_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  VMMT_CUDA(cudaLaunchKernelEx(&cfg, kernel, params));
  return vmmt_check_launch(what);
}

}  // namespace

bool vmmt_lstm_tc_supported(int ndir, int N, int H) {
  TcPlan p;
  return (ndir == 1 || ndir == 2) && tc_plan(ndir, N, H, &p);
}

int vmmt_lstm_tc_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, cudaStream_t s) {
  TcPlan p;
  if (!tc_plan(ndir, N, H, &p)) return VMMT_EINVAL;
  TcFwdParams P;
  for (int d = 0; d < ndir; ++d) P.d[d] = dirs[d];
  if (ndir == 1) P.d[1] = dirs[0];
  P.lengths = lengths;
  P.T = T; P.N = N; P.H = H; P.C = p.C; P.G = p.G; P.Ng = p.Ng; P.Kp = p.Kp;
  P.trace = nullptr;
  static long long* tbuf = nullptr;
  if (getenv("VMMT_LSTM_TRACE")) {
    if (!tbuf) cudaMalloc(&tbuf, 128);
    P.trace = tbuf;
    int rc = tc_ts_mode() ? cluster_launch(lstm_tc_fwd_kernel<true>, P, ndir * p.G * p.C, p.C, p.smem_f, s, "lstm_tc_fwd_kernel")
                          : cluster_launch(lstm_tc_fwd_kernel<false>, P, ndir * p.G * p.C, p.C, p.smem_f, s, "lstm_tc_fwd_kernel");
    cudaStreamSynchronize(s);
    long long h[16];
    cudaMemcpy(h, tbuf, 128, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[lstm trace] mma warp: start %lld after_fence %lld issued %lld\n", h[13] - h[0], h[14] - h[0], h[1] - h[0]);
    fprintf(stderr, "[lstm trace] T=%d setup %lld cycles, loop %lld cycles (%lld per step), wall %lld ns\n", T, h[9] - h[8], h[10] - h[9], (h[10] - h[9]) / T, h[12] - h[11]);
    fprintf(stderr, "[lstm trace] C=%d G=%d Ng=%d step5 cycles: mma_issued %lld | epi: wait_start %lld mma_done %lld gates_xchg %lld cells %lld dsmem %lld cluster_bar %lld\n",
            p.C, p.G, p.Ng, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0]);
    return rc;
  }
  return tc_ts_mode() ? cluster_launch(lstm_tc_fwd_kernel<true>, P, ndir * p.G * p.C, p.C, p.smem_f, s, "lstm_tc_fwd_kernel")
                      : cluster_launch(lstm_tc_fwd_kernel<false>, P, ndir * p.G * p.C, p.C, p.smem_f, s, "lstm_tc_fwd_kernel");
}

int vmmt_lstm_tc_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, cudaStream_t s) {
  TcPlan p;
  if (!tc_plan(ndir, N, H, &p)) return VMMT_EINVAL;
  TcBwdParams P;
  for (int d = 0; d < ndir; ++d) P.d[d] = dirs[d];
  if (ndir == 1) P.d[1] = dirs[0];
  P.lengths = lengths;
  P.T = T; P.N = N; P.H = H; P.C = p.C; P.G = p.G; P.Ng = p.Ng; P.Kp = p.Kp;
  return tc_ts_mode() ? cluster_launch(lstm_tc_bwd_kernel<true>, P, ndir * p.G * p.C, p.C, p.smem_b, s, "lstm_tc_bwd_kernel")
                      : cluster_launch(lstm_tc_bwd_kernel<false>, P, ndir * p.G * p.C, p.C, p.smem_b, s, "lstm_tc_bwd_kernel");
}
