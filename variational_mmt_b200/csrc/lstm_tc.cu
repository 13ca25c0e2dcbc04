// Cluster-resident LSTM recurrence on tcgen05 tensor cores (forward and BPTT), H <= 512.
//
// Reference semantics: torch nn.LSTM (onmt/Models.py:124-149, 892-893; onmt/VI_Model1.py:106), gate
// order i,f,g,o.  gx = x W_ih^T for all timesteps comes from one batched tensor-core GEMM (vmmt_gemm).
//
// The batch rows of a recurrence are independent, so they are split into G groups of <= 16 rows.  One
// thread-block CLUSTER of C = ceil(H/32) CTAs (<= 16, one per SM) serves one (direction, group) for the whole
// sequence with no global-memory round trip and NO cluster-wide barrier per step:
//   * CTA c owns 32 hidden units = 128 gate rows [i(32) f(32) g(32) o(32)] of W_hh.  Its 128 x Kp slice stays
//     resident in TENSOR MEMORY as fp16 for the whole sequence (A operand of tcgen05.mma from TMEM).  fp16 carries
//     the same 11 significand bits as TF32, which is what a tensor-core "fp32" GEMM rounds its operands to;
//     accumulation is fp32 in TMEM; c / h state stays fp32 in registers, h is rounded to fp16 only as the next
//     step's MMA operand.
//   * forward step : D[128 gate rows, 8 or 16 batch cols] = W_slice[128,Kp] * h_{t-1}^T  (Kp/16 tcgen05.mma M128 K16 with the
//                    A operand in TMEM, independent accumulator chains per row group) -> 8 epilogue warps drain the
//                    accumulator (tcgen05.ld) -> gates regrouped through a double-buffered shared-memory tile ->
//                    sigmoid/tanh cell update fused in the epilogue (cell (n = warp + 8 r, unit = lane); pinned
//                    __fmaf_rn / __fadd_rn arithmetic so that the 8-row and 16-row instantiations round alike) -> the
//                    new fp16 h values are packed 8 lanes -> 16 bytes by shuffles and sent with st.async (16 B,
//                    mbarrier::complete_tx) straight from registers into the K-major operand tile of EVERY CTA of the
//                    cluster; the next step's gx[t+1] is prefetched meanwhile.  The MMA warp of each CTA waits on the
//                    tile's mbarrier only (expect_tx = C * rows * 64 bytes): producer -> consumer data flow replaces a
//                    cluster barrier; the consumer issues fence.proxy.async before the MMA reads the tile.
//   * backward step: the SAME slice transposed (A = W_slice^T, resident in TMEM): each CTA contracts its own 128
//                    gate rows, partial dh[512 units, batch] = W_slice^T * dG_t^T (dG enters as fp16 hi + lo columns
//                    of dG * 2^12: 22 significand bits, no gradient magnitude is lost); the partials are
//                    reduce-scattered to the unit owners with one bulk copy per PEER owner (cp.async.bulk
//                    shared::cta -> shared::cluster, completing on the owner's mbarrier; the CTA's own block is read
//                    from its staging buffer) and summed in a fixed order (deterministic).
// Double buffering + the data dependence of the recurrence itself make every buffer reuse safe (see the comments
// at the copies); compute-sanitizer memcheck / racecheck / synccheck are clean (profiles/sanitizer_*_r2.txt).
// Measured per step at H = 500, 8 rows per cluster (ncu launch list, profiles/launches_r2_summary.txt): 1.8 us forward,
// 2.4 us backward; VMMT_LSTM_FLOOR=1 runs the hand-off chain without the cell arithmetic (the measured floor bench.py reports).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <stdio.h>
#include <mutex>
#include "common.cuh"
#include "vmmt_internal.h"
#include "lstm_tc.h"

namespace {

constexpr int UC = 32;                 // hidden units per CTA
constexpr int NP = 16;                 // batch columns per cluster (MMA N)
constexpr int EW = 8;                  // epilogue / cell-owner warps: warp w drains TMEM lane quadrant w & 3
constexpr int THREADS = 32 * (EW + 1); // + warp EW: MMA issuer
constexpr float DG_SCALE = 4096.0f;    // backward: dG enters the MMA as fp16 hi + lo of dG * 2^12
constexpr uint32_t TMEM_COLS = 512;    // whole tensor memory: the allocation then starts at column 0
constexpr uint32_t TMEM_D = 256;       // accumulators; columns [0,256) hold the weight slice

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// shared::cta -> shared::cluster bulk copy, completion (bytes) signalled on an mbarrier of the destination CTA
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
// 16-byte store into another CTA's shared memory; its completion counts 16 tx bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(dst_cluster), "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
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

// issue-only TMEM loads (the caller waits once for all of them)
__device__ __forceinline__ void tmem_ld_x4_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x8_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x16_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.mma kind::f16, A from TMEM, B through a shared-memory descriptor given as its two 32-bit halves (the issue
// loop only ever advances the low half: start address >> 4)
__device__ __forceinline__ void umma_f16_ts2(uint32_t tmem_c, uint32_t tmem_a, uint32_t bd_lo, uint32_t bd_hi, uint32_t idesc,
                                             uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 d;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 d, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d, %4, p;\n"
      "}\n" ::"r"(tmem_c), "r"(tmem_a), "r"(bd_lo), "r"(bd_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// fast transcendental forms for the tensor-core path (their ~1e-6 error is far below the operand rounding)
// clock read that cannot be scheduled before `dep` is available (trace points inside arithmetic)
__device__ __forceinline__ long long clock_after(float dep) {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(dep) : "memory");
  return t;
}
__device__ __forceinline__ float fex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// every operation is pinned (__f*_rn is never contracted or re-associated): the NRG = 1 and NRG = 2 instantiations must
// round identically, or a row's result would depend on how many rows share its cluster (fp32 differences of one ulp are
// amplified to 2^-11 by the fp16 rounding of h)
__device__ __forceinline__ float fsigmoid(float x) { return frcp(__fadd_rn(1.0f, fex2(__fmul_rn(-1.4426950408889634f, x)))); }
__device__ __forceinline__ float ftanh(float x) {
  return __fmaf_rn(-2.0f, frcp(__fadd_rn(fex2(__fmul_rn(2.8853900817779268f, x)), 1.0f)), 1.0f);
}

// shared-memory matrix descriptors (sm_100 version bits).
//   swizzled  : K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO); LBO unused
//   unswizzled: K-major core matrices (8 rows x 16 B, 128 B contiguous); LBO = byte distance between core matrices
//               adjacent in K, SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_desc_none(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// byte offset of the 16-byte chunk holding elements [8*chunk, 8*chunk+8) of row `row` in a [rows][64 x 2B]
// 128B-swizzled block
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }
struct TcFwdParams {
  VmmtLstmDir d[2];
  const int64_t* lengths;
  int T, N, H, C, G, Ng, Kp;
  int floor;            // measurement only (env VMMT_LSTM_FLOOR): one k-block of MMAs and no transcendentals per step --
                        // what is left is the synchronisation floor (hand-off + mbarriers + tcgen05.ld + gate exchange)
  long long* trace;
};
struct TcBwdParams {
  VmmtLstmDirBwd d[2];
  const int64_t* lengths;
  int T, N, H, C, G, Ng, Kp;
  int floor;
  long long* trace;
};

// The slice lives in TMEM as the MMA's A operand (lane = row, one 32-bit column = two consecutive k).
// forward: row r = gate*32 + unit (lane quadrant = gate, lane = unit), columns k/2 for k in [0, Kp).
// Warp w loads gate w & 3, k-half w >> 2.  The global reads are coalesced (a quarter-warp reads 128 contiguous bytes
// of one W_hh row), the 32 x 32 block is transposed through a padded shared-memory tile so that lane = row, and the
// next block's loads are in flight while the current one is packed and stored.
constexpr int WST = 36;                // padded row stride (floats) of the per-warp staging tile: conflict-free LDS.128
__device__ __forceinline__ void load_w_tmem_fwd(const float* __restrict__ w_hh, int H, int Kp, int u0, int warp, int lane,
                                                float* stage_all) {
  const int q = warp & 3, hf = warp >> 2;
  float* st = stage_all + (size_t)warp * 32 * WST;
  const bool vec = (H & 3) == 0;
  const int kbeg = hf * (Kp / 2), nblk = Kp / 64;                 // 32-wide k blocks of this warp's half
  const int rr = lane >> 3, cc = (lane & 7) * 4;                  // this lane reads rows rr + 4 j, floats cc..cc+3
  auto fetch = [&](int blk, float4 (&v)[8]) {
    const int k = kbeg + blk * 32 + cc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int unit = u0 + rr + 4 * j;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (unit < H) {
        const float* row = w_hh + ((size_t)q * H + unit) * H;
        if (vec && k + 3 < H) x = __ldg(reinterpret_cast<const float4*>(row + k));
        else {
          if (k < H) x.x = __ldg(row + k);
          if (k + 1 < H) x.y = __ldg(row + k + 1);
          if (k + 2 < H) x.z = __ldg(row + k + 2);
          if (k + 3 < H) x.w = __ldg(row + k + 3);
        }
      }
      v[j] = x;
    }
  };
  float4 cur[8], nxt[8];
  fetch(0, cur);
  for (int blk = 0; blk < nblk; ++blk) {
    if (blk + 1 < nblk) fetch(blk + 1, nxt);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(st + (rr + 4 * j) * WST + cc) = cur[j];
    __syncwarp();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(st + lane * WST + 4 * i);
      pk[2 * i] = pack_h2(x.x, x.y);
      pk[2 * i + 1] = pack_h2(x.z, x.w);
    }
    tmem_st_x16(((uint32_t)(32 * q) << 16) + (uint32_t)((kbeg + blk * 32) / 2), pk);
#pragma unroll
    for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// backward: A = W_slice^T: tile mt, row = unit 128 mt + 32 q + lane, columns j/2 for this CTA's 128 gate rows j.
// Lanes are consecutive units = consecutive addresses of one W_hh row: coalesced as is.  Warp w serves lane quadrant
// w & 3 and the gates 2 (w >> 2), 2 (w >> 2) + 1; all 32 loads of a (tile, gate) block are issued before the first use.
__device__ __forceinline__ void load_w_tmem_bwd(const float* __restrict__ w_hh, int H, int u0, int n_mt, int warp, int lane) {
  const int q = warp & 3, hf = warp >> 2;
  for (int mt = 0; mt < n_mt; ++mt) {
    const int u = 128 * mt + 32 * q + lane;
#pragma unroll 1
    for (int gi = 0; gi < 2; ++gi) {
      const int gate = 2 * hf + gi;
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int ul = u0 + i;
        f[i] = (u < H && ul < H) ? __ldg(w_hh + ((size_t)gate * H + ul) * H + u) : 0.f;
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_h2(f[2 * i], f[2 * i + 1]);
      tmem_st_x16(((uint32_t)(32 * q) << 16) + (uint32_t)(mt * 64 + gate * 16), pk);
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}


// forward MMA issue with a compile-time k extent: k-step j of block kb accumulates into chain j
template <int NRG, int NKB>
__device__ __forceinline__ void issue_fwd_mmas(uint32_t bd, uint32_t bd_hi, uint32_t idesc) {
  constexpr uint32_t CHUNK = NRG * 128;
  constexpr int NPc = 8 * NRG;
#pragma unroll
  for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      umma_f16_ts2(TMEM_D + j * NPc, (uint32_t)(kb * 32 + j * 8), bd + (uint32_t)(((kb * 8 + j * 2) * CHUNK) >> 4), bd_hi, idesc,
                   kb > 0 ? 1u : 0u);
}
// ------------------------------------------------------------------------------------------------ forward
// NRG = 8-row groups of batch columns per cluster (MMA N = 8 NRG): 1 when the group has <= 8 rows, else 2.
// The bytes a CTA sends per step (C copies of 512 NRG B) are what bounds the hand-off (DSMEM ~20 B/clk/SM).
template <int NRG>
__global__ void __launch_bounds__(THREADS, 1) lstm_tc_fwd_kernel(const TcFwdParams P) {
  constexpr int NPc = 8 * NRG;                               // batch columns of the MMA
  constexpr int RM = NRG;                                    // cells per thread (rows warp + 8 r)
  constexpr int HN = NPc / 2;                                // accumulator columns one warp drains
  constexpr uint32_t CHUNK = NRG * 128;                      // one k-chunk (8 k) of the h tile: NRG core matrices
  constexpr uint32_t SLICE = 4 * CHUNK;                      // one CTA's h slice: 32 units = 4 k-chunks, contiguous
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int Kp = P.Kp, nkb = Kp / 64;
  const uint32_t tile_bytes = (uint32_t)(Kp / 8) * CHUNK;    // one h operand tile [Kp/8 chunks][NRG][8 rows][8 fp16]
  uint8_t* Bt = sm;                                          // [2][tile]
  // gate exchange, double buffered by step parity: the reads of step s and the writes of step s+1 are ordered only through
  // the mbarrier chain (h hand-off -> MMA -> accumulator), the writes of step s+2 also by two named barriers
  float* gsm0 = reinterpret_cast<float*>(Bt + 2 * tile_bytes);                // [2][4 gates][NPc][32]
  uint8_t* hstage = reinterpret_cast<uint8_t*>(gsm0 + 2 * 4 * NPc * 32);      // (unused tail kept for alignment of the barriers)
  uint64_t* bars = reinterpret_cast<uint64_t*>(hstage + 2 * SLICE);           // [0] mma_done, [1..2] h_full[buf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* wstage = reinterpret_cast<float*>(bars + 4);                         // [EW][32][WST] weight-load staging
  const uint32_t bt_addr = base;
  const uint32_t bar_mma = smem_u32(bars), bar_full0 = smem_u32(bars + 1);
  auto ht_off = [](int n, int k) -> uint32_t {               // byte offset of element (row n, k) in a tile / slice
    return (uint32_t)(k >> 3) * CHUNK + (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int C = P.C;
  const int cl = blockIdx.x / C;                             // cluster index = (dir, group)
  const int crank = (int)cluster_rank();
  const int dir = cl / P.G, grp = cl % P.G;
  const VmmtLstmDir& D = P.d[dir];
  const int T = P.T, N = P.N, H = P.H;
  const int n0 = grp * P.Ng;
  const int nn = min(P.Ng, N - n0);
  const int u0 = crank * UC;

  if (P.trace && blockIdx.x == 0 && tid == 0) { P.trace[8] = clock64(); unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); P.trace[11] = (long long)g; }
  // h_{-1}: every CTA fills its own operand tile (buffer 0) from h0; rows >= nn and k >= H are zero in both buffers
  for (int e = tid; e < 2 * (Kp / 8) * NPc; e += blockDim.x) {
    const int bufi = e / ((Kp / 8) * NPc), rem = e % ((Kp / 8) * NPc);
    const int chunk = rem / NPc, row = rem % NPc;
    __align__(16) __half hv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = chunk * 8 + i;
      const float v = (bufi == 0 && D.h0 && row < nn && k < H) ? D.h0[(size_t)(n0 + row) * H + k] : 0.0f;
      hv[i] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(Bt + (size_t)bufi * tile_bytes + ht_off(row, chunk * 8)) = *reinterpret_cast<uint4*>(hv);
  }
  if (tid == 0) {
    mbar_init(bar_mma, 1);
    mbar_init(bar_full0, 1);
    mbar_init(bar_full0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy smem writes -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (*tmem_slot != 0u) __trap();                            // the whole TMEM was requested: the base must be column 0
  if (warp < EW) {
    load_w_tmem_fwd(D.w_hh, H, Kp, u0, warp, lane, wstage);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  // cell ownership: thread (warp w < 8, lane) owns cells (n = w + 8 r, unit = lane), r < nr <= RM (nr is warp-uniform)
  float c[RM], h[RM], bias[RM][4];
  int len[RM];
  const int unit = u0 + lane;
  const bool ulive = warp < EW && unit < H;
  const int nr = (warp < EW && nn > warp) ? (nn - warp + 7) / 8 : 0;
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    const int n = warp + 8 * r;
    len[r] = T; c[r] = 0.f; h[r] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[r][g] = 0.f;
    if (ulive && n < nn) {
      const int gn = n0 + n;
      if (D.c0) c[r] = D.c0[(size_t)gn * H + unit];
      if (D.h0) h[r] = D.h0[(size_t)gn * H + unit];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const size_t j = (size_t)g * H + unit;
        bias[r][g] = __fadd_rn(__fadd_rn(D.b_ih ? D.b_ih[j] : 0.f, D.b_hh ? D.b_hh[j] : 0.f),
                               D.rowbias ? D.rowbias[(size_t)gn * 4 * H + j] : 0.f);
      }
      if (P.lengths) len[r] = (int)P.lengths[gn];
    }
  }
  const int ucl = min(unit, H - 1);                          // clamped: dead lanes read valid memory, results discarded
  // instruction descriptor: D f32, A/B fp16 K-major, N = NPc, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(NPc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // un-swizzled K-major B descriptor: core matrices adjacent in K are CHUNK apart (LBO), 8-row groups 128 B apart (SBO)
  const uint32_t bd_hi = (uint32_t)(128u >> 4) | (1u << 14);
  const uint32_t bd_lo0 = ((bt_addr >> 4) & 0x3FFFu) | ((CHUNK >> 4) << 16);
  const bool leader = elect_one();
  cluster_sync_all();          // every CTA of the cluster is initialised (tiles, mbarriers) before any remote copy lands
  if (P.trace && blockIdx.x == 0 && tid == 0) P.trace[9] = clock64();

  const int q = warp & 3, hf = warp >> 2;
  // this lane's two hand-off destinations (CTAs l & 7, (l & 7) + 8): their operand tiles and h_full barriers
  uint32_t rtile[2], rbar[2];
  bool dst_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int d = (lane & 7) + 8 * i;
    dst_ok[i] = d < C;
    rtile[i] = mapa(bt_addr, (uint32_t)min(d, C - 1));
    rbar[i] = mapa(bar_full0, (uint32_t)min(d, C - 1));
  }
  float gxn[RM][4];
  auto load_gx = [&](int tt) {
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const int n = min(warp + 8 * r, nn - 1);
      const float* gp = D.gx + ((size_t)tt * N + n0 + n) * 4 * H + ucl;
#pragma unroll
      for (int g = 0; g < 4; ++g) gxn[r][g] = (r < nr) ? __ldg(gp + (size_t)g * H) : 0.f;
    }
  };
  if (warp < EW) load_gx(D.reverse ? T - 1 : 0);
  for (int s = 0; s < T; ++s) {
    const int t = D.reverse ? T - 1 - s : s;
    const int buf = s & 1;
    const bool tr = P.trace != nullptr && blockIdx.x == 0 && s == 5;
    if (tr && tid == 0) P.trace[0] = clock64();
    if (warp == EW) {
      if (leader) {
        // arm next step's hand-off barrier: C slices will land in the other tile
        if (s + 1 < T) mbar_arrive_expect_tx(bar_full0 + 8u * (uint32_t)(buf ^ 1), (uint32_t)(C * nn) * 64u);
        if (s > 0) mbar_wait(bar_full0 + 8u * (uint32_t)buf, (uint32_t)(((s - 1) >> 1) & 1));   // h_{s-1} of all CTAs landed
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the peers' st.async data -> the MMA's operand reads
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tr) P.trace[14] = clock64();
        // k-step j of block kb accumulates into chain j: four independent accumulator chains, 2 k-chunks per k-step
        const uint32_t bd = bd_lo0 + (uint32_t)buf * (tile_bytes >> 4);
        if (P.floor) issue_fwd_mmas<NRG, 1>(bd, bd_hi, idesc);
        else if (nkb == 8) issue_fwd_mmas<NRG, 8>(bd, bd_hi, idesc);     // H in (448, 512]: every offset an immediate
        else if (nkb == 4) issue_fwd_mmas<NRG, 4>(bd, bd_hi, idesc);     // H in (192, 256]
        else {
          uint32_t a = 0, b = bd;
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_f16_ts2(TMEM_D + j * NPc, a + j * 8, b + j * ((2 * CHUNK) >> 4), bd_hi, idesc, 0u);
          for (int kb = 1; kb < nkb; ++kb) {
            a += 32; b += (8 * CHUNK) >> 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_f16_ts2(TMEM_D + j * NPc, a + j * 8, b + j * ((2 * CHUNK) >> 4), bd_hi, idesc, 1u);
          }
        }
        umma_commit(bar_mma);
        if (tr) P.trace[1] = clock64();
      }
      __syncwarp();
    } else {
      // input-projection terms: this step's were loaded during the previous step (gxn), the next step's loads are
      // issued now and complete under this step's MMAs / cell update
      float gx[RM][4];
#pragma unroll
      for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int g = 0; g < 4; ++g) gx[r][g] = __fadd_rn(gxn[r][g], bias[r][g]);
      if (s + 1 < T) load_gx(D.reverse ? T - 2 - s : s + 1);
      if (tr && tid == 0) P.trace[2] = clock64();
      mbar_wait(bar_mma, (uint32_t)(s & 1));
      if (tr && tid == 0) P.trace[3] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* gsm = gsm0 + (size_t)buf * 4 * NPc * 32;
      {
        // gate q of batch columns [hf HN, hf HN + HN): the four chains' partial sums
        uint32_t v[4][HN];
        const uint32_t ta = TMEM_D + ((uint32_t)(32 * q) << 16) + (uint32_t)(hf * HN);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (NRG == 1) tmem_ld_x4_nowait(ta + j * NPc, v[j]); else tmem_ld_x8_nowait(ta + j * NPc, v[j]);
        }
        tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // accumulator reads precede the next step's MMAs
#pragma unroll
        for (int n = 0; n < HN; ++n)
          gsm[(q * NPc + hf * HN + n) * 32 + lane] =
              __fadd_rn(__fadd_rn(__uint_as_float(v[0][n]), __uint_as_float(v[1][n])),
                        __fadd_rn(__uint_as_float(v[2][n]), __uint_as_float(v[3][n])));
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tr && tid == 0) P.trace[4] = clock64();
      float av[RM][4], hout[RM];                             // activated gates / masked output, stored after the hand-off
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        if (r < nr) {                                        // warp-uniform
          const int n = warp + 8 * r;
          const float Gi = __fadd_rn(gsm[(0 * NPc + n) * 32 + lane], gx[r][0]);
          const float Gf = __fadd_rn(gsm[(1 * NPc + n) * 32 + lane], gx[r][1]);
          const float Gg = __fadd_rn(gsm[(2 * NPc + n) * 32 + lane], gx[r][2]);
          const float Go = __fadd_rn(gsm[(3 * NPc + n) * 32 + lane], gx[r][3]);
          if (tr && tid == 0 && r == 0) P.trace[16] = clock_after(Gi + Gf + Gg + Go);
          if (P.floor) { av[r][0] = Gi; av[r][1] = Gf; av[r][2] = Gg; av[r][3] = Go; }
          else { av[r][0] = fsigmoid(Gi); av[r][1] = fsigmoid(Gf); av[r][2] = ftanh(Gg); av[r][3] = fsigmoid(Go); }
        }
      }
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        if (r < nr) {
          const int n = warp + 8 * r;
          const float cn = __fmaf_rn(av[r][1], c[r], __fmul_rn(av[r][0], av[r][2]));
          const float hn = P.floor ? cn : __fmul_rn(av[r][3], ftanh(cn));
          const bool m = t < len[r];
          c[r] = m ? cn : c[r];
          h[r] = m ? hn : h[r];
          hout[r] = m ? hn : 0.f;
          if (tr && tid == 0 && r == 0) P.trace[17] = clock_after(hn);
          if (s + 1 < T) {
            // Hand-off: the 8 lanes that hold one 16-byte chunk (8 consecutive units of row n, fp16) assemble it with
            // shuffles, then lane l stores it straight into the OTHER operand tile of CTAs l & 7 and (l & 7) + 8
            // (st.async: the store completes on the destination's h_full barrier; own CTA included).  Reuse is safe
            // without further synchronisation: a destination's tile (buf^1) was last read by its MMAs of step s-1,
            // which completed before that CTA produced the h_{s-1} values this CTA had to receive before computing h_s.
            const uint32_t hb = (uint32_t)__half_as_ushort(__float2half_rn(ulive ? h[r] : 0.f));
            const uint32_t pb = __shfl_xor_sync(0xffffffffu, hb, 1);
            const uint32_t w = (lane & 1) ? (pb | (hb << 16)) : (hb | (pb << 16));
            const int gb = lane & ~7;
            const uint32_t w0 = __shfl_sync(0xffffffffu, w, gb), w1 = __shfl_sync(0xffffffffu, w, gb + 2);
            const uint32_t w2 = __shfl_sync(0xffffffffu, w, gb + 4), w3 = __shfl_sync(0xffffffffu, w, gb + 6);
            const uint32_t off = (uint32_t)(buf ^ 1) * tile_bytes + (uint32_t)(crank * 4 + (lane >> 3)) * CHUNK +
                                 (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u;
            const uint32_t boff = 8u * (uint32_t)(buf ^ 1);
            if (tr && tid == 0 && r == 0) P.trace[18] = clock_after(__uint_as_float(w0 ^ w1 ^ w2 ^ w3));
            if (dst_ok[0]) st_async_v4(rtile[0] + off, w0, w1, w2, w3, rbar[0] + boff);
            if (dst_ok[1]) st_async_v4(rtile[1] + off, w0, w1, w2, w3, rbar[1] + boff);
          }
        }
      }
      if (tr && tid == 0) P.trace[13] = clock64();
      if (tr && tid == 0) P.trace[6] = clock64();
      // outputs and the tensors saved for backward leave after the hand-off: off the recurrence's critical path
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        const int n = warp + 8 * r;
        if (ulive && r < nr) {
          const size_t row = (size_t)t * N + n0 + n;
          D.out[row * D.out_ld + unit] = hout[r];
          if (D.gates) {
            float* gp = D.gates + row * 4 * H + unit;
            gp[0] = av[r][0]; gp[(size_t)H] = av[r][1]; gp[(size_t)2 * H] = av[r][2]; gp[(size_t)3 * H] = av[r][3];
          }
          if (D.cs) D.cs[row * H + unit] = c[r];
        }
      }
    }
    if (tr && tid == 0) P.trace[7] = clock64();
  }
  if (P.trace && blockIdx.x == 0 && tid == 0) { P.trace[10] = clock64(); unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); P.trace[12] = (long long)g; }
  if (warp < EW) {
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const int n = warp + 8 * r;
      if (ulive && r < nr) {
        if (D.hT) D.hT[(size_t)(n0 + n) * H + unit] = h[r];
        if (D.cT) D.cT[(size_t)(n0 + n) * H + unit] = c[r];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == EW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "n"(TMEM_COLS));
  cluster_sync_all();          // no CTA exits while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------ backward
template <int NRG>
__global__ void __launch_bounds__(THREADS, 1) lstm_tc_bwd_kernel(const TcBwdParams P) {
  constexpr int NPc = 8 * NRG;
  constexpr int RM = NRG;
  constexpr int KB_BYTES = 2 * NPc * 128;                    // one k-block (64 gate rows) of the dG tile: hi rows | lo rows
  constexpr int PART = NPc * 32 * 4;                         // one (source, owner) block of dh partials
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* Bt = sm;                                          // [2 k-blocks][2 NPc rows][128 B] swizzled: dG hi rows 0..NPc-1, lo rows NPc..
  float* recv = reinterpret_cast<float*>(Bt + 2 * KB_BYTES);              // [2][16 src][NPc][32]
  float* stage = recv + 2 * 16 * NPc * 32;                                // [2][16 owners][NPc][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2 * 16 * NPc * 32);// [0] mma_done, [1..2] recv_full[buf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const uint32_t recv_addr = smem_u32(recv), stage_addr = smem_u32(stage);
  const uint32_t bar_mma = smem_u32(bars), bar_recv0 = smem_u32(bars + 1);

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
  const uint32_t part_bytes = (uint32_t)nn * 128u;           // only the rows that exist travel

  if (P.trace && blockIdx.x == 0 && tid == 0) P.trace[8] = clock64();
  for (int e = tid; e < 2 * KB_BYTES / 16; e += blockDim.x) reinterpret_cast<uint4*>(Bt)[e] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar_mma, 1);
    mbar_init(bar_recv0, 1);
    mbar_init(bar_recv0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (*tmem_slot != 0u) __trap();
  const int n_mt = (P.C * UC + 127) / 128;                  // 128-unit output tiles that hold real units
  if (warp < EW) {
    load_w_tmem_bwd(D.w_hh, H, u0, n_mt, warp, lane);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  float dc[RM], dhr[RM];
  int len[RM];
  const int unit = u0 + lane;
  const bool ulive = warp < EW && unit < H;
  const int ucl = min(unit, H - 1);
  const int nr = (warp < EW && nn > warp) ? (nn - warp + 7) / 8 : 0;
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    const int n = warp + 8 * r;
    len[r] = T; dc[r] = 0.f; dhr[r] = 0.f;
    if (ulive && n < nn) {
      const int gn = n0 + n;
      if (D.dcT) dc[r] = D.dcT[(size_t)gn * H + unit];
      if (D.dhT) dhr[r] = D.dhT[(size_t)gn * H + unit];
      if (P.lengths) len[r] = (int)P.lengths[gn];
    }
  }
  // D f32, A fp16 from TMEM (W_slice^T), B fp16 K-major, N = 2 NPc (hi | lo), M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)((2 * NPc) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // 128-byte-swizzled K-major B descriptor (8-row groups 1024 B apart), as make_desc_sw128
  const uint32_t bd_hi = (uint32_t)(1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t bd_lo0 = ((base >> 4) & 0x3FFFu) | (1u << 16);
  const bool leader = elect_one();
  cluster_sync_all();
  if (P.trace && blockIdx.x == 0 && tid == 0) P.trace[9] = clock64();

  // Everything of a step that does not depend on the incoming dh is loaded one step ahead (the loads are issued
  // before the MMA wait, the arithmetic runs while the reduce-scatter is in flight):
  //   pf[r] = {a1, b0, b1, b2, b3, fg, dout},  dh = dh_rec + dout,  dct = dc + dh a1,
  //   dG = {dct b0, dct b1, dct b2, dh b3},  dc' = dct fg.
  float pf[RM][7];
  auto load_raw = [&](int sp) {                              // raw {ig, fg, gg, og, ct, cp, dout}: loads only (branch-free)
    const int tq = D.reverse ? sp : T - 1 - sp;
    const int tp = D.reverse ? tq + 1 : tq - 1;
    const bool has_prev = tp >= 0 && tp < T;
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      if (r < nr) {
        const int gn = n0 + warp + 8 * r;
        const size_t row = (size_t)tq * N + gn;
        const float* gp = D.gates + row * 4 * H + ucl;
        pf[r][0] = __ldg(gp); pf[r][1] = __ldg(gp + (size_t)H); pf[r][2] = __ldg(gp + (size_t)2 * H); pf[r][3] = __ldg(gp + (size_t)3 * H);
        pf[r][4] = __ldg(D.cs + row * H + ucl);
        pf[r][5] = has_prev ? __ldg(D.cs + ((size_t)tp * N + gn) * H + ucl) : (D.c0 ? __ldg(D.c0 + (size_t)gn * H + ucl) : 0.f);
        pf[r][6] = D.dout ? __ldg(D.dout + row * D.dout_ld + ucl) : 0.f;
      }
    }
  };
  auto derive = [&]() {                                      // raw -> {a1, b0, b1, b2, b3, fg, dout}
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      if (r < nr) {
        const float ig = pf[r][0], fg = pf[r][1], gg = pf[r][2], og = pf[r][3], ct = pf[r][4], cp = pf[r][5];
        const float tc = ftanh(ct);
        pf[r][0] = og * (1.f - tc * tc);
        pf[r][1] = gg * ig * (1.f - ig);
        pf[r][2] = cp * fg * (1.f - fg);
        pf[r][3] = ig * (1.f - gg * gg);
        pf[r][4] = tc * og * (1.f - og);
        pf[r][5] = fg;
      }
    }
  };
  if (warp < EW) { load_raw(0); derive(); }
  const int q = warp & 3, hf = warp >> 2;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};                      // bias gradient of this thread's unit: sum of dG over its rows and t
  float rsum[RM][4];                                         // per-row sums over t: gradient of the per-example additive term
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int g = 0; g < 4; ++g) rsum[r][g] = 0.f;

  for (int s = 0; s < T; ++s) {
    const int t = D.reverse ? s : T - 1 - s;
    const int buf = s & 1;
    const bool last = (s + 1 == T);
    const bool tr = P.trace != nullptr && blockIdx.x == 0 && s == 5;
    if (tr && tid == 0) P.trace[0] = clock64();
    float dGs[RM][4];
    if (warp < EW) {
      // ---- elementwise BPTT of this CTA's cells -> dG_t as fp16 hi/lo operand rows n / NPc+n
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        if (r < nr) {
          const int n = warp + 8 * r;
          const bool m = ulive && t < len[r];
          const float dh = dhr[r] + pf[r][6];
          const float dct = dc[r] + dh * pf[r][0];
          dGs[r][0] = m ? dct * pf[r][1] : 0.f;
          dGs[r][1] = m ? dct * pf[r][2] : 0.f;
          dGs[r][2] = m ? dct * pf[r][3] : 0.f;
          dGs[r][3] = m ? dh * pf[r][4] : 0.f;
          dc[r] = m ? dct * pf[r][5] : dc[r];
          dhr[r] = m ? 0.f : dhr[r];                        // consumed (the next value comes from the reduce below) / frozen: passes through
#pragma unroll
          for (int g = 0; g < 4; ++g) { bsum[g] += dGs[r][g]; rsum[r][g] += dGs[r][g]; }
          // operand tile: k = gate*32 + lane (this CTA's 128 gate rows), row n (hi) and NPc + n (lo)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            // fp16 hi + lo of dG * 2^12: 22 significand bits, magnitudes from 1.5e-11 to 16 (clamped beyond)
            const float xs = fminf(fmaxf(dGs[r][g] * DG_SCALE, -60000.f), 60000.f);
            const __half hi = __float2half_rn(xs);
            const __half lo = __float2half_rn(xs - __half2float(hi));
            const int k = g * 32 + lane, kb = k >> 6, ch = (k & 63) >> 3, el = k & 7;
            *reinterpret_cast<__half*>(Bt + kb * KB_BYTES + sw_off(n, ch) + el * 2) = hi;
            *reinterpret_cast<__half*>(Bt + kb * KB_BYTES + sw_off(NPc + n, ch) + el * 2) = lo;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    const bool stop = last && D.dh0 == nullptr;             // the last dh_prev only feeds dh0
    if (!stop) __syncthreads();
    if (tr && tid == 0) P.trace[1] = clock64();
    if (warp == EW) {
      if (stop) break;
      if (leader) {
        mbar_arrive_expect_tx(bar_recv0 + 8u * (uint32_t)buf, (uint32_t)(C - 1) * part_bytes);   // the C - 1 PEERS' blocks
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int mt = 0; mt < (P.floor ? 1 : n_mt); ++mt) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)                    // K = 128 gate rows = 8 steps of 16
            umma_f16_ts2(TMEM_D + mt * 32, (uint32_t)(mt * 64 + ks * 8),
                         bd_lo0 + (uint32_t)(((ks >> 2) * KB_BYTES + (ks & 3) * 32) >> 4), bd_hi, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_mma);
      }
      __syncwarp();
    } else {
      // dG_t leaves for global memory (it is d gx, consumed by the batched weight / input gradient GEMMs)
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        if (ulive && r < nr) {
          float* dg = D.dgates + ((size_t)t * N + n0 + warp + 8 * r) * 4 * H + unit;
          dg[0] = dGs[r][0]; dg[(size_t)H] = dGs[r][1]; dg[(size_t)2 * H] = dGs[r][2]; dg[(size_t)3 * H] = dGs[r][3];
        }
      }
      if (stop) break;
      if (!last) load_raw(s + 1);                           // next step's saved activations: issued now, consumed after the hand-off
      mbar_wait(bar_mma, (uint32_t)(s & 1));
      if (tr && tid == 0) P.trace[2] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // partial dh[unit = 128 mt + 32 q + lane, n] -> staged per owner CTA 4 mt + q as [n][lane]; warp w drains the
      // tiles mt = hf, hf + 2 of its lane quadrant
      float* st = stage + (size_t)buf * 16 * NPc * 32;
      {
        uint32_t v[2][2 * NPc];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int mt = hf + 2 * i;
          if (mt < n_mt) {
            const uint32_t ta = TMEM_D + ((uint32_t)(32 * q) << 16) + (uint32_t)(mt * 32);
            if (NRG == 1) tmem_ld_x16_nowait(ta, v[i]);
            else { tmem_ld_x16_nowait(ta, v[i]); tmem_ld_x16_nowait(ta + 16, v[i] + 16); }
          }
        }
        tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int mt = hf + 2 * i, owner = 4 * mt + q;
          if (mt < n_mt && owner < C) {
#pragma unroll
            for (int n = 0; n < NPc; ++n)
              st[(owner * NPc + n) * 32 + lane] = (__uint_as_float(v[i][n]) + __uint_as_float(v[i][NPc + n])) * (1.0f / DG_SCALE);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tr && tid == 0) P.trace[3] = clock64();
      // reduce-scatter: one block of nn rows per owner, into the owner's recv[buf][src = this CTA], completing on
      // the owner's recv_full[buf].  stage[buf] / recv[buf] are reused at step s+2: by then this CTA has consumed
      // every owner's step-(s+1) block, each sent after that owner consumed this step's copies.
      // (the block this CTA owns itself is read straight from `stage` below: a bulk copy targets ANOTHER CTA)
      const int dstc = (lane < 2) ? warp * 2 + lane : 99;
      if (dstc < C && dstc != crank) {
        const uint32_t dst = mapa(recv_addr + (uint32_t)((buf * 16 + crank) * PART), (uint32_t)dstc);
        const uint32_t bar = mapa(bar_recv0 + 8u * (uint32_t)buf, (uint32_t)dstc);
        bulk_copy_to_cluster(dst, stage_addr + (uint32_t)((buf * 16 + dstc) * PART), part_bytes, bar);
      }
      if (tr && tid == 0) P.trace[6] = clock64();
      if (!last) derive();                                  // while the reduce-scatter is in flight
      mbar_wait(bar_recv0 + 8u * (uint32_t)buf, (uint32_t)((s >> 1) & 1));
      if (tr && tid == 0) P.trace[4] = clock64();
      const float* rb = recv + (size_t)buf * 16 * NPc * 32;
      const float* own = st + (size_t)crank * NPc * 32;      // this CTA's own partial block never left shared memory
      auto part = [&](int src, int n) -> float {
        return src == crank ? own[n * 32 + lane] : rb[(src * NPc + n) * 32 + lane];
      };
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        if (r < nr) {
          const int n = warp + 8 * r;
          float s0 = dhr[r], s1 = 0.f, s2 = 0.f, s3 = 0.f;  // dhr: non-zero only for frozen (masked) cells
          for (int src = 0; src + 3 < C; src += 4) {
            s0 += part(src + 0, n);
            s1 += part(src + 1, n);
            s2 += part(src + 2, n);
            s3 += part(src + 3, n);
          }
          for (int src = C & ~3; src < C; ++src) s0 += part(src, n);
          dhr[r] = (s0 + s1) + (s2 + s3);
        }
      }
    }
    if (tr && tid == 0) P.trace[5] = clock64();
  }
  if (P.trace && blockIdx.x == 0 && tid == 0) P.trace[10] = clock64();
  if (warp < EW) {
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const int n = warp + 8 * r;
      if (ulive && r < nr) {
        if (D.dh0) D.dh0[(size_t)(n0 + n) * H + unit] = dhr[r];
        if (D.dc0) D.dc0[(size_t)(n0 + n) * H + unit] = dc[r];
        if (D.drow) {                                       // d(rowbias)[n, gate, unit] = sum_t dG: one owner per element
          float* dr = D.drow + (size_t)(n0 + n) * 4 * H + unit;
          dr[0] = rsum[r][0]; dr[(size_t)H] = rsum[r][1]; dr[(size_t)2 * H] = rsum[r][2]; dr[(size_t)3 * H] = rsum[r][3];
        }
      }
    }
  }
  // bias gradients db_ih = db_hh = sum over (t, n) of dG: the 8 cell-owner warps hold per-unit partial sums over their rows;
  // they are added in warp order through shared memory (the receive buffer is idle: the last hand-off has been consumed)
  // and one atomicAdd per (gate, unit) joins the batch groups' shares (replaces a column-sum pass over the 10 MB dG tensor)
  if (D.db_ih != nullptr || D.db_hh != nullptr) {
    __syncthreads();
    if (warp < EW) {
#pragma unroll
      for (int g = 0; g < 4; ++g) recv[(warp * 4 + g) * 32 + lane] = bsum[g];
    }
    __syncthreads();
    if (warp < 4 && u0 + lane < H) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < EW; ++w) tot += recv[(w * 4 + warp) * 32 + lane];
      if (D.db_ih) atomicAdd(D.db_ih + (size_t)warp * H + u0 + lane, tot);
      if (D.db_hh) atomicAdd(D.db_hh + (size_t)warp * H + u0 + lane, tot);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == EW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "n"(TMEM_COLS));
  cluster_sync_all();
}

struct TcPlan { int C, G, Ng, Kp, nrg; };

size_t fwd_smem(int Kp, int nrg) {
  return 128 + (size_t)2 * (Kp / 8) * nrg * 128 + (size_t)2 * 4 * 8 * nrg * 32 * 4 + (size_t)2 * 4 * nrg * 128 + 64 +
         (size_t)EW * 32 * WST * 4;
}
size_t bwd_smem(int nrg) { return 1024 + (size_t)2 * 2 * 8 * nrg * 128 + (size_t)4 * 16 * 8 * nrg * 32 * 4 + 64; }

template <typename K>
int prepare_kernel(K kernel, int C, size_t smem, int* max_clusters) {
  // function attributes are sticky: set them once per (kernel, device); the co-resident cluster count is cached.
  // Keyed by the kernel ADDRESS: the template instantiations share one function-pointer type.
  struct Slot { const void* fn; int dev; size_t smem; int maxc[17]; };
  static Slot slots[32];
  static int nslots = 0;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  Slot* sl = nullptr;
  for (int i = 0; i < nslots; ++i)
    if (slots[i].fn == (const void*)kernel && slots[i].dev == dev) sl = &slots[i];
  if (!sl) {
    if (nslots == 32) { vmmt_set_error("lstm_tc: kernel attribute cache full"); return VMMT_EINVAL; }
    sl = &slots[nslots++];
    sl->fn = (const void*)kernel; sl->dev = dev; sl->smem = 0;
    for (int i = 0; i < 17; ++i) sl->maxc[i] = 0;
    VMMT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  if (smem > sl->smem) {
    VMMT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sl->smem = smem;
    for (int i = 0; i < 17; ++i) sl->maxc[i] = 0;
  }
  if (sl->maxc[C] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
    sl->maxc[C] = n;
  }
  *max_clusters = sl->maxc[C];
  return VMMT_OK;
}

bool tc_shape_ok(int ndir, int N, int H) {
  return (ndir == 1 || ndir == 2) && H >= 32 && H <= 512 && N >= 1 && ceil_div(H, UC) <= 16;
}

// batch groups: as many clusters as can be co-resident (one wave; a cluster of 16 needs a whole GPC), at most 16
// rows per group; groups of <= 8 rows run the N = 8 variant (half the hand-off bytes)
void tc_plan(int ndir, int N, int H, int max_clusters, int cluster_budget, TcPlan* p) {
  p->C = ceil_div(H, UC);
  p->Kp = ceil_div(H, 64) * 64;
  const char* e = getenv("VMMT_LSTM_GROUPS");
  if (cluster_budget > 0) max_clusters = min(max_clusters, cluster_budget);
  int gmax = e ? atoi(e) : max(1, max_clusters / ndir);
  int G = max(1, min(gmax, N));
  int Ng = ceil_div(N, G);
  if (Ng > NP) Ng = NP;
  p->Ng = Ng;
  p->G = ceil_div(N, Ng);
  p->nrg = Ng <= 8 ? 1 : 2;
}

template <typename K, typename PT>
int cluster_launch(K kernel, const PT& params, int grid, int C, size_t smem, cudaStream_t s, const char* what) {
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

long long* trace_buf() {
  static long long* tbuf = nullptr;
  if (!tbuf) { cudaMalloc(&tbuf, 256); cudaMemset(tbuf, 0, 256); }
  return tbuf;
}

template <int NRG>
int launch_fwd(TcFwdParams& P, const TcPlan& p, int ndir, int maxc, cudaStream_t s) {
  const size_t smem = fwd_smem(p.Kp, NRG);
  int dummy = 0;
  int rc = prepare_kernel(lstm_tc_fwd_kernel<NRG>, p.C, smem, &dummy);
  if (rc) return rc;
  if (!getenv("VMMT_LSTM_TRACE"))
    return cluster_launch(lstm_tc_fwd_kernel<NRG>, P, ndir * p.G * p.C, p.C, smem, s, "lstm_tc_fwd_kernel");
  P.trace = trace_buf();
  rc = cluster_launch(lstm_tc_fwd_kernel<NRG>, P, ndir * p.G * p.C, p.C, smem, s, "lstm_tc_fwd_kernel");
  cudaStreamSynchronize(s);
  long long h[24];
  cudaMemcpy(h, P.trace, 192, cudaMemcpyDeviceToHost);
  fprintf(stderr, "[lstm fwd trace] cells: gates_read %lld h_ready %lld chunk_packed %lld\n", h[16] - h[0], h[17] - h[0], h[18] - h[0]);
  fprintf(stderr, "[lstm fwd trace] T=%d C=%d G=%d Ng=%d nrg=%d max_clusters=%d setup %lld cyc, loop %lld cyc (%lld per step), block0 wall %lld ns\n",
          P.T, p.C, p.G, p.Ng, NRG, maxc, h[9] - h[8], h[10] - h[9], (h[10] - h[9]) / P.T, h[12] - h[11]);
  fprintf(stderr, "[lstm fwd trace] step5: mma warp h_full %lld issued %lld | epi: gx_loaded %lld mma_done %lld gates_xchg %lld cells %lld fence+bar %lld copies_issued %lld end %lld\n",
          h[14] - h[0], h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[13] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0]);
  return rc;
}

template <int NRG>
int launch_bwd(TcBwdParams& P, const TcPlan& p, int ndir, int maxc, cudaStream_t s) {
  const size_t smem = bwd_smem(NRG);
  int dummy = 0;
  int rc = prepare_kernel(lstm_tc_bwd_kernel<NRG>, p.C, smem, &dummy);
  if (rc) return rc;
  if (!getenv("VMMT_LSTM_TRACE"))
    return cluster_launch(lstm_tc_bwd_kernel<NRG>, P, ndir * p.G * p.C, p.C, smem, s, "lstm_tc_bwd_kernel");
  P.trace = trace_buf();
  rc = cluster_launch(lstm_tc_bwd_kernel<NRG>, P, ndir * p.G * p.C, p.C, smem, s, "lstm_tc_bwd_kernel");
  cudaStreamSynchronize(s);
  long long h[16];
  cudaMemcpy(h, P.trace, 128, cudaMemcpyDeviceToHost);
  fprintf(stderr, "[lstm bwd trace] T=%d C=%d G=%d Ng=%d nrg=%d max_clusters=%d setup %lld cyc, loop %lld cyc (%lld per step)\n",
          P.T, p.C, p.G, p.Ng, NRG, maxc, h[9] - h[8], h[10] - h[9], (h[10] - h[9]) / P.T);
  fprintf(stderr, "[lstm bwd trace] step5: dG_ready %lld mma_done %lld staged %lld copies_issued %lld recv_full %lld end %lld\n",
          h[1] - h[0], h[2] - h[0], h[3] - h[0], h[6] - h[0], h[4] - h[0], h[5] - h[0]);
  return rc;
}

}  // namespace

bool vmmt_lstm_tc_supported(int ndir, int N, int H) { return tc_shape_ok(ndir, N, H); }

int vmmt_lstm_tc_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int cluster_budget,
                     cudaStream_t s) {
  if (!tc_shape_ok(ndir, N, H)) return VMMT_EINVAL;
  TcPlan p;
  const int Kp = ceil_div(H, 64) * 64;
  int maxc = 1;
  int rc = prepare_kernel(lstm_tc_fwd_kernel<2>, ceil_div(H, UC), fwd_smem(Kp, 2), &maxc);
  if (rc) return rc;
  tc_plan(ndir, N, H, maxc, cluster_budget, &p);
  TcFwdParams P;
  for (int d = 0; d < ndir; ++d) P.d[d] = dirs[d];
  if (ndir == 1) P.d[1] = dirs[0];
  P.lengths = lengths;
  P.T = T; P.N = N; P.H = H; P.C = p.C; P.G = p.G; P.Ng = p.Ng; P.Kp = p.Kp;
  P.floor = getenv("VMMT_LSTM_FLOOR") ? 1 : 0;
  P.trace = nullptr;
  return p.nrg == 1 ? launch_fwd<1>(P, p, ndir, maxc, s) : launch_fwd<2>(P, p, ndir, maxc, s);
}

int vmmt_lstm_tc_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int cluster_budget,
                     cudaStream_t s) {
  if (!tc_shape_ok(ndir, N, H)) return VMMT_EINVAL;
  TcPlan p;
  int maxc = 1;
  int rc = prepare_kernel(lstm_tc_bwd_kernel<2>, ceil_div(H, UC), bwd_smem(2), &maxc);
  if (rc) return rc;
  tc_plan(ndir, N, H, maxc, cluster_budget, &p);
  TcBwdParams P;
  for (int d = 0; d < ndir; ++d) P.d[d] = dirs[d];
  if (ndir == 1) P.d[1] = dirs[0];
  P.lengths = lengths;
  P.T = T; P.N = N; P.H = H; P.C = p.C; P.G = p.G; P.Ng = p.Ng; P.Kp = p.Kp;
  P.floor = getenv("VMMT_LSTM_FLOOR") ? 1 : 0;
  P.trace = nullptr;
  return p.nrg == 1 ? launch_bwd<1>(P, p, ndir, maxc, s) : launch_bwd<2>(P, p, ndir, maxc, s);
}
