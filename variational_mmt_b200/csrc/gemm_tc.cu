// tcgen05 (5th-gen tensor core) TF32 GEMM with fp32 storage, TMA operand staging and TMEM accumulators.
//
// C[M,N] (row-major, ldc) = act( op(A)[M,K] * op(B)[K,N] + bias[N] ) (+ C)        -- same contract as
// gemm_simt.cu; replaces nn.Linear / torch.mm on the hot path (GlobalAttention.py:71,78,113,188;
// NormalVariationalEncoder.py:18-25,35-43; the input projections inside nn.LSTM, Models.py:124-129;
// the generator projection, ModelConstructor.py:582-585) and their dgrad / wgrad counterparts.
//
// Design (one CTA per 128 x BN output tile, optional split-K over gridDim.z):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d global -> 128B-swizzled shared memory, STAGES-deep
//               mbarrier ring.  fp32 operands are loaded as TFLOAT32 (TMA rounds to nearest TF32).
//   warp 1      MMA issuer: one thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) from shared-memory
//               descriptors, fp32 accumulator in TMEM; tcgen05.commit frees ring slots / signals the epilogue.
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> per-warp smem
//               transpose -> bias / activation / accumulate -> coalesced 128-byte row stores.
// Both operands may be K-major ([rows,K], the nn.Linear layout) or MN-major ([K,rows]) -- forward,
// dgrad and wgrad all run on the same kernel with no transposing copy: the UMMA descriptors and the
// instruction descriptor's major bits select the layout.  Ragged M / N / K need no padding: TMA
// zero-fills out-of-bounds box elements and the epilogue masks its stores.
//
// bf16 variant (BF = true; BASELINE configs[1] "fp32 and bf16"): the SAME kernel on bf16 operands -- a 128-byte swizzle row
// holds 64 bf16 instead of 32 fp32 and one tcgen05.mma.kind::f16 consumes 16 of them (the same 32 bytes), so the shared
// memory geometry, the ring, the TMEM accumulators (fp32) and every epilogue are unchanged; only the tensor maps' element
// type / box, the instruction descriptor and the MN-major tile geometry (64-element rows, 8-row swizzle atoms) differ.
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>
#include <stdlib.h>
#include <stdio.h>
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                    // fp32 elements per stage along K = one 128-byte swizzle row (bf16: 64 elements)
constexpr int UMMA_K = 8;                 // tf32: 32 bytes per instruction along K (bf16: 16 elements)
constexpr int EPI_WARPS = 4;
constexpr int THREADS = 32 * (2 + EPI_WARPS);
constexpr int A_STAGE_BYTES = BM * BK * 4;               // 16 KB

__host__ __device__ constexpr int b_stage_bytes(int BN) { return BN * BK * 4; }
__host__ __device__ constexpr int epi_bytes() { return EPI_WARPS * 2 * 4096 + 1024; }   // 2 x 4 KB TMA staging per warp + bias tile
// Every GEMM CTA asks for at least EXCLUSIVE_SMEM bytes of dynamic shared memory, i.e. for an SM of its own as far as
// other tensor-memory users go: the LSTM recurrence kernels (lstm_tc.cu) allocate all 512 TMEM columns for the whole
// sequence, so a GEMM CTA placed beside one would sit in tcgen05.alloc until the recurrence ends (and a recurrence
// CTA placed beside a GEMM CTA stalls its whole cluster: measured 148 -> 222 us on the last encoder layer when the
// register footprints happened to allow co-residency).  21 KB (smallest recurrence CTA) + 208 KB > 227 KB.
constexpr size_t EXCLUSIVE_SMEM = 208 * 1024;
__host__ __device__ constexpr size_t smem_bytes(int BN, int STAGES) {
  const size_t need = 1024 /*align slack*/ + (size_t)STAGES * (A_STAGE_BYTES + b_stage_bytes(BN)) + epi_bytes() + 256;   // 256 >= 16*STAGES + 32 (barriers) + 8
  return need > EXCLUSIVE_SMEM ? need : EXCLUSIVE_SMEM;
}

// ---------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
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

// shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle
//   K-major : rows of 128 B (32 fp32 along K); 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: tf32 only supports the "128B swizzle with 32-byte atoms" layout (type 1; TMA
//             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 128-B rows hold 32 MN-elements of one k, 4 k-rows form a
//             512-B swizzle atom (SBO = distance between 4-row groups), 32-element MN groups are
//             `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                             uint32_t layout_type = 2 /*SWIZZLE_128B*/) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case VMMT_ACT_RELU: return fmaxf(v, 0.0f);
    case VMMT_ACT_TANH: return tanhf(v);
    case VMMT_ACT_SOFTPLUS: return softplusf_(v);
    case VMMT_ACT_SIGMOID: return sigmoidf_(v);
    default: return v;
  }
}

__device__ __forceinline__ void trace_mark(long long* tr, int slot) {
  if (tr) tr[(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 16 + slot] = clock64();
}

struct GemmArgs {
  float* C;
  int64_t ldc;
  int M, N, K;
  const float* bias;
  int act, accumulate;
  int kb_per_split;      // k-blocks (of BK) per gridDim.z slice
  int tma_store;         // 1: epilogue stores / reduces through the tmC tensor map
  int dbg;
  int tiles_m, tiles_n; // output tiles (the grid is 1-D over CTAs: persistent tile loop)
  int nkb2;             // k-blocks of the optional second operand pair (C = A B^T + A2 B2^T, both K-major), else 0
  long long* trace;     // debug: per-CTA timeline (8 slots), or null
  // fused generator epilogues (vmmt_internal.h: VmmtGenEpi): 0 none, 1 per-row log-sum-exp partials instead of C,
  // 2 C = softmax-NLL gradient of the tile
  int epi_mode;
  float4* lse_part;             // mode 1: [gridDim.x][M] {max, sum exp(x - max), best logit, best column}
  float* tgt_logit;             // mode 1: [M] logit of the target column (written by the tile that holds it)
  const int64_t* target;        // modes 1, 2: [M]
  const float* row_lse;         // mode 2: [M]
  const float* gscale;          // mode 2: device scalar or null
  float scale;                  // mode 2
  long long pad;                // mode 2: ignored target id
  int topk;                     // mode 3
  float2* tile_lse;             // mode 3: [tiles_n][M] {max, sum exp}
  float2* tile_cand;            // mode 3: [tiles_n][M][topk] {logit, column}
  int group_m;                  // row-tiles per rasterisation group (see tile_origin)
};

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  if (ACT == VMMT_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == VMMT_ACT_TANH) return tanhf(v);
  if (ACT == VMMT_ACT_SOFTPLUS) return softplusf_(v);
  if (ACT == VMMT_ACT_SIGMOID) return sigmoidf_(v);
  return v;
}

// One epilogue warp drains its 32 accumulator lanes (rows) x BN columns: TMEM -> registers -> per-warp smem
// transpose -> rows of 32 consecutive columns, 8 independent rows in flight per lane (coalesced 128-B accesses).
template <int BN, int ACT>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& g, uint32_t tmem_base, float* sc, int q, int lane,
                                              int m0, int n0, int mode) {
  const int row_base = m0 + 32 * q;
  const int nrows = min(32, g.M - row_base);       // may be <= 0 for fully out-of-range quarters
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    if (n0 + c * 32 >= g.N) break;
    const int col = n0 + c * 32 + lane;
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32), v);
    if (c == 0 && q == 2 && lane == 0) trace_mark(g.trace, 8);
#pragma unroll
    for (int j = 0; j < 32; ++j) sc[lane * 33 + j] = __uint_as_float(v[j]);
    __syncwarp();
    if (c == 0 && q == 2 && lane == 0) trace_mark(g.trace, 9);
    if (col < g.N && !(g.dbg & 1)) {
      const float bv = (g.bias != nullptr && mode != 3) ? __ldg(g.bias + col) : 0.0f;
      float* cbase = g.C + (int64_t)row_base * g.ldc + col;
#pragma unroll 1
      for (int r0 = 0; r0 < nrows; r0 += 8) {
        float x[8], old[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = sc[(r0 + i) * 33 + lane] + bv;
        if (mode == 1 || mode == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) old[i] = (r0 + i < nrows) ? cbase[(int64_t)(r0 + i) * g.ldc] : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (r0 + i < nrows) {
            float* cp = cbase + (int64_t)(r0 + i) * g.ldc;
            if (mode == 3) atomicAdd(cp, x[i]);
            else if (mode == 2) *cp = act_t<ACT>(x[i] + old[i]);
            else if (mode == 1) *cp = old[i] + act_t<ACT>(x[i]);
            else *cp = act_t<ACT>(x[i]);
          }
        }
      }
    }
    __syncwarp();
    if (c == 0 && q == 2 && lane == 0) trace_mark(g.trace, 10);
  }
}

// Epilogue through TMA: each warp finishes 32 rows x 32 columns in registers (bias, activation), writes them to
// a 128B-swizzled 4 KB staging tile (conflict-free 16-byte stores) and one lane issues a bulk tensor store -- or
// a bulk tensor reduce-add for C += (wgrad accumulation, split-K partial sums).  The TMA clips ragged edges.
// Generator forward: the tile's logits never leave the SM.  Each epilogue thread owns one row of the tile (TMEM lane)
// and folds its BN columns into {running max, sum of exponentials, best logit / column, target logit}.
template <int BN>
__device__ __forceinline__ void epilogue_lse(const GemmArgs& g, uint32_t tmem_base, const float* bias_s, int q, int lane,
                                             int m0, int n0) {
  const int row = m0 + 32 * q + lane;
  const long long tg = (row < g.M) ? (long long)g.target[row] : -1;
  float mx = -INFINITY, sum = 0.f, bestv = -INFINITY, tlogit = 0.f;
  int besti = 0x7fffffff;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 >= g.N) break;
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32), v);
    float x[32], cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      x[j] = (col0 + j < g.N) ? __uint_as_float(v[j]) + bias_s[c * 32 + j] : -INFINITY;
      if (x[j] > bestv) { bestv = x[j]; besti = col0 + j; }       // ascending columns: ties keep the lowest index
      cm = fmaxf(cm, x[j]);
      if ((long long)(col0 + j) == tg) tlogit = x[j];
    }
    const float nm = fmaxf(mx, cm);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += __expf(x[j] - nm);          // exp(-inf) = 0 for the columns beyond N
    sum = sum * __expf(mx - nm) + s;
    mx = nm;
  }
  if (row < g.M) {
    g.lse_part[(size_t)(n0 / BN) * g.M + row] = make_float4(mx, sum, bestv, __int_as_float(besti));
    if (tg >= n0 && tg < n0 + BN) g.tgt_logit[row] = tlogit;
  }
}

// Beam-search generator: per row of the tile {max, sum exp} and the tile's top-K logits (value, column), best first;
// ties keep the lowest column (the rule of beam_advance_kernel / the CPU oracle).  Nothing of size M x N is written.
template <int BN, int KL>      // KL = compile-time list length (>= g.topk): 5 covers the published beam size, 8 the maximum
__device__ __forceinline__ void epilogue_topk(const GemmArgs& g, uint32_t tmem_base, const float* bias_s, int q, int lane,
                                              int m0, int n0) {
  const int row = m0 + 32 * q + lane;
  const int K = g.topk;
  float mx = -INFINITY, sum = 0.f;
  float bv[KL];
  int bi[KL];
#pragma unroll
  for (int k = 0; k < KL; ++k) { bv[k] = -INFINITY; bi[k] = 0x7fffffff; }
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 >= g.N) break;
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32), v);
    float x[32], cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      x[j] = (col0 + j < g.N) ? __uint_as_float(v[j]) + bias_s[c * 32 + j] : -INFINITY;
      cm = fmaxf(cm, x[j]);
    }
    // the list always holds KL entries (compile-time register indices); the first K are written out
    if (cm > bv[KL - 1]) {                                        // some column of the chunk enters the list
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (x[j] > bv[KL - 1]) {                                  // ascending columns + strict '>' keep the lowest column
          bv[KL - 1] = x[j]; bi[KL - 1] = col0 + j;
#pragma unroll
          for (int k = KL - 1; k > 0; --k)
            if (bv[k] > bv[k - 1]) {
              const float tv = bv[k]; bv[k] = bv[k - 1]; bv[k - 1] = tv;
              const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
            }
        }
      }
    }
    const float nm = fmaxf(mx, cm);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += __expf(x[j] - nm);
    sum = sum * __expf(mx - nm) + s;
    mx = nm;
  }
  if (row < g.M) {
    const size_t t = (size_t)(n0 / BN) * g.M + row;
    g.tile_lse[t] = make_float2(mx, sum);
    float2* dst = g.tile_cand + t * K;
#pragma unroll
    for (int k = 0; k < KL; ++k)
      if (k < K) dst[k] = make_float2(bv[k], __int_as_float(bi[k]));
  }
}

template <int BN, int ACT, bool DL = false>
__device__ __forceinline__ void epilogue_tile_tma(const GemmArgs& g, const CUtensorMap* tmC, uint32_t tmem_base,
                                                  uint32_t stage_u32, const float* bias_s, int q, int lane, int m0,
                                                  int n0, int mode) {
  const int row0 = m0 + 32 * q;
  if (row0 >= g.M) return;
  // DL: the stored value is the softmax-NLL gradient (exp(x - lse[row]) - [col == target]) * scale, 0 for ignored rows
  float dl_lse = 0.f, dl_scale = 0.f;
  long long dl_tg = -1;
  if (DL) {
    const int row = row0 + lane;
    if (row < g.M) {
      dl_tg = (long long)g.target[row];
      dl_lse = g.row_lse[row];
      dl_scale = (dl_tg != g.pad) ? g.scale * (g.gscale ? g.gscale[0] : 1.0f) : 0.f;
    }
  }
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 >= g.N) break;
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32), v);
    if (c == 0 && q == 2 && lane == 0) trace_mark(g.trace, 8);
    const uint32_t buf = stage_u32 + (uint32_t)(c & 1) * 4096u;
    if (c >= 2) {      // the bulk store that last read this buffer must have finished reading shared memory
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
    }
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float x[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t = __uint_as_float(v[4 * j4 + e]);
        if (mode != 3) t += bias_s[c * 32 + 4 * j4 + e];
        if (DL) x[e] = (__expf(t - dl_lse) - ((long long)(col0 + 4 * j4 + e) == dl_tg ? 1.0f : 0.0f)) * dl_scale;
        else x[e] = act_t<ACT>(t);
      }
      const uint32_t addr = buf + (uint32_t)lane * 128u + (uint32_t)((j4 ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (c == 0 && q == 2 && lane == 0) trace_mark(g.trace, 9);
    if (lane == 0 && !(g.dbg & 1)) {
      if (mode == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(tmC), "r"(col0), "r"(row0), "r"(buf) : "memory");
      } else {
        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(tmC), "r"(col0), "r"(row0), "r"(buf) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (c == 0 && q == 2 && lane == 0) trace_mark(g.trace, 10);
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// Persistent tile loop: CTA `cta` of `ncta` processes tiles cta, cta + ncta, ... (with split-K: exactly one tile per
// CTA).  Two TMEM accumulators alternate between tiles, so the epilogue warps drain tile i (tcgen05.ld, bias /
// activation / LSE / top-K, TMA stores) while the MMA warp already contracts tile i+1 and the producer warp runs up to
// STAGES k-blocks ahead of it -- the smem ring and its phases simply continue across tiles.
// GEN = true instantiates the generator epilogues (LSE / dlogits / top-K); they roughly double the register count, so the
// plain GEMMs get their own instantiation (96 instead of 180 registers per thread: two 64-wide CTAs per SM, and room
// for a recurrence CTA beside a weight-gradient CTA on the same SM).
template <int BN, int STAGES, bool A_MN, bool B_MN, bool GEN, bool BF = false>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB2, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int B_STAGE = b_stage_bytes(BN);
  constexpr int STAGE = A_STAGE_BYTES + B_STAGE;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B atoms are 1024-B aligned
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* epi = reinterpret_cast<float*>(gen + (size_t)STAGES * STAGE);
  const uint32_t bar0 = base + STAGES * STAGE + epi_bytes();
  // barriers: full[s] at bar0 + 8 s, empty[s] after them, then tmem_full[2], tmem_empty[2]
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * STAGES, tfull0 = bar0 + 16 * STAGES, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (size_t)STAGES * STAGE + epi_bytes() + 16 * STAGES + 32);

  constexpr int BKE = BF ? 2 * BK : BK;          // ELEMENTS per k-block (one 128-byte row)
  constexpr int MNE = BF ? 64 : 32;              // MN-major: elements per 128-byte row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb_total = (g.K + BKE - 1) / BKE;
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int kb1 = min(nkb_total, kb0 + g.kb_per_split);
  const int nkb1 = kb1 - kb0;                            // k-blocks of the first operand pair per tile
  const int nkb = (g.dbg & 2) ? 0 : nkb1 + g.nkb2;       // >= 1 by construction of the grid
  const int tiles_total = g.tiles_m * g.tiles_n;
  const int cta = blockIdx.x, ncta = gridDim.x;

  // Tile rasterisation: tiles are numbered so that consecutive ones walk DOWN a group of GROUP_M row-tiles before moving
  // to the next column-tile: the ~148 co-resident CTAs share GROUP_M A-tiles and ~10 B-tiles out of L2 instead of one
  // A-tile and 148 B-tiles (the generator's [V,H] weight is larger than L2 at cfg5: without this every row-tile sweep
  // re-read all of it from HBM -- 39 GB of DRAM traffic for 0.3 GB of operands).  The group's A panel (GROUP_M x 128 rows x
  // K) has to stay L2-resident while B streams past it once per group: 64 row-tiles when K is small enough for a
  // <= 48 MB panel (generator forward: B re-read 5 x instead of 20 x at cfg5), 16 otherwise (host side: group_m).
  auto tile_origin = [&](int lin, int& m0, int& n0) {
    const int GROUP_M = g.group_m;
    const int grp = lin / (GROUP_M * g.tiles_n);
    const int first_m = grp * GROUP_M;
    const int gm = min(GROUP_M, g.tiles_m - first_m);
    const int within = lin - grp * GROUP_M * g.tiles_n;
    m0 = (first_m + within % gm) * BM;
    n0 = (within / gm) * BN;
  };

  if (threadIdx.x == 0) trace_mark(g.trace, 0);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (g.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    if (g.nkb2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // TMEM allocation (2 x BN fp32 accumulator columns), address written to shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_mark(g.trace, 1);

  // Producer and MMA warps run CONVERGED (all 32 lanes execute the loops, one elected lane issues): descriptor /
  // coordinate values then stay in uniform registers.  Inside an `if (lane == 0)` region ptxas has to move every
  // operand of UTMALDG / UTCHMMA into uniform registers with an ELECT + R2UR.BROADCAST loop (~65 cycles per MMA).
  const bool leader = elect_one();
  if (warp == 0) {
    int it = 0;                                          // k-blocks issued so far (ring position across tiles)
    for (int lin = cta; lin < tiles_total; lin += ncta) {
      int m0, n0;
      tile_origin(lin, m0, n0);
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t fb = full0 + 8 * s;
        const uint32_t sa = base + s * STAGE, sb = sa + A_STAGE_BYTES;
        const bool second = i >= nkb1;                             // k-blocks of the second operand pair follow the first
        const int k = second ? (i - nkb1) * BKE : (kb0 + i) * BKE;
        const CUtensorMap* pa = second ? &tmA2 : &tmA;
        const CUtensorMap* pb = second ? &tmB2 : &tmB;
        if (leader) {
          mbar_expect_tx(fb, STAGE);
          if (!A_MN) {
            tma_load_2d(sa, pa, k, m0, fb);                        // box {128 B of k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / MNE; ++j) tma_load_2d(sa + j * (BKE * 128), pa, m0 + MNE * j, k, fb);   // box {128 B of m, BKE k}
          }
          if (!B_MN) {
            tma_load_2d(sb, pb, k, n0, fb);                        // box {128 B of k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / MNE; ++j) tma_load_2d(sb + j * (BKE * 128), pb, n0 + MNE * j, k, fb);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D=f32, A=B=tf32 (kind::tf32 format 2) or bf16 (kind::f16 format 1), majors, N>>3, M>>4
    constexpr uint32_t FMT = BF ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((A_MN ? 1u : 0u) << 15) |
                           ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int it = 0, lt = 0;
    for (int lin = cta; lin < tiles_total; lin += ncta, ++lt) {
      const int acc = lt & 1;
      const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
      mbar_wait(tempty0 + 8 * acc, ((lt >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator (2 tiles ago)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        if (lt == 0 && i == 0 && lane == 0) trace_mark(g.trace, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * STAGE, sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < BK / UMMA_K; ++j) {
          // MN-major, one instruction = 8 (tf32) / 16 (bf16) k-rows of 128 B: tf32 uses the 32-byte-atom swizzle (4-row
          // atoms, SBO 512, layout type 1), bf16 the plain 128-byte swizzle (8-row atoms, SBO 1024, layout type 2); LBO is
          // the distance between the 128-byte-wide MN groups of the tile
          const uint64_t ad = A_MN ? (BF ? make_desc(sa + j * 2048, BKE * 128, 1024, 2) : make_desc(sa + j * 1024, BK * 128, 512, 1))
                                   : make_desc(sa + j * 32, 16, 1024);
          const uint64_t bd = B_MN ? (BF ? make_desc(sb + j * 2048, BKE * 128, 1024, 2) : make_desc(sb + j * 1024, BK * 128, 512, 1))
                                   : make_desc(sb + j * 32, 16, 1024);
          if (leader) {
            if (BF) umma_bf16(tacc, ad, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);
            else umma_tf32(tacc, ad, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);
          }
        }
        if (leader) umma_commit(empty0 + 8 * s);          // slot reusable once these MMAs have read it
        __syncwarp();
      }
      if (leader) umma_commit(tfull0 + 8 * acc);          // accumulator complete
      __syncwarp();
      if (lt == 0 && lane == 0) trace_mark(g.trace, 3);
    }
  } else {
    // ---------------- epilogue warps: TMEM lane quarter = warp % 4
    const int q = warp & 3;
    float* sc = epi + (warp - 2) * 2048;            // fallback-path transpose scratch (32 x 33 floats)
    const int mode = (gridDim.z > 1) ? 3 : g.accumulate;        // 0 store, 1 C += v, 2 act(C + v), 3 atomic
    int lt = 0;
    for (int lin = cta; lin < tiles_total; lin += ncta, ++lt) {
      int m0, n0;
      tile_origin(lin, m0, n0);
      const int acc = lt & 1;
      const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
      mbar_wait(tfull0 + 8 * acc, (lt >> 1) & 1);
      if (lt == 0 && threadIdx.x == 64) trace_mark(g.trace, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if ((g.tma_store && mode != 2) || (GEN && (g.epi_mode == 1 || g.epi_mode == 3))) {
        float* bias_s = epi + EPI_WARPS * 2 * 1024;                // [BN] staged once per tile by the 4 warps
        if (lt > 0) asm volatile("bar.sync 1, 128;" ::: "memory");  // every warp has finished with the previous tile's bias
        for (int i = threadIdx.x - 64; i < BN; i += 32 * EPI_WARPS)
          bias_s[i] = (g.bias != nullptr && n0 + i < g.N) ? __ldg(g.bias + n0 + i) : 0.0f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t st = smem_u32(epi) + (uint32_t)(warp - 2) * 8192u;
        if (GEN && g.epi_mode == 1) {
          if constexpr (GEN) epilogue_lse<BN>(g, tacc, bias_s, q, lane, m0, n0);
        } else if (GEN && g.epi_mode == 3) {
          if constexpr (GEN) {
            if (g.topk <= 5) epilogue_topk<BN, 5>(g, tacc, bias_s, q, lane, m0, n0);
            else epilogue_topk<BN, VMMT_TOPK_MAX>(g, tacc, bias_s, q, lane, m0, n0);
          }
        } else if (GEN && g.epi_mode == 2) {
          if constexpr (GEN) epilogue_tile_tma<BN, VMMT_ACT_NONE, true>(g, &tmC, tacc, st, bias_s, q, lane, m0, n0, mode);
        } else
        switch (g.act) {
          case VMMT_ACT_RELU: epilogue_tile_tma<BN, VMMT_ACT_RELU>(g, &tmC, tacc, st, bias_s, q, lane, m0, n0, mode); break;
          case VMMT_ACT_TANH: epilogue_tile_tma<BN, VMMT_ACT_TANH>(g, &tmC, tacc, st, bias_s, q, lane, m0, n0, mode); break;
          case VMMT_ACT_SOFTPLUS: epilogue_tile_tma<BN, VMMT_ACT_SOFTPLUS>(g, &tmC, tacc, st, bias_s, q, lane, m0, n0, mode); break;
          case VMMT_ACT_SIGMOID: epilogue_tile_tma<BN, VMMT_ACT_SIGMOID>(g, &tmC, tacc, st, bias_s, q, lane, m0, n0, mode); break;
          default: epilogue_tile_tma<BN, VMMT_ACT_NONE>(g, &tmC, tacc, st, bias_s, q, lane, m0, n0, mode); break;
        }
      } else {
        switch (g.act) {
          case VMMT_ACT_RELU: epilogue_tile<BN, VMMT_ACT_RELU>(g, tacc, sc, q, lane, m0, n0, mode); break;
          case VMMT_ACT_TANH: epilogue_tile<BN, VMMT_ACT_TANH>(g, tacc, sc, q, lane, m0, n0, mode); break;
          case VMMT_ACT_SOFTPLUS: epilogue_tile<BN, VMMT_ACT_SOFTPLUS>(g, tacc, sc, q, lane, m0, n0, mode); break;
          case VMMT_ACT_SIGMOID: epilogue_tile<BN, VMMT_ACT_SIGMOID>(g, tacc, sc, q, lane, m0, n0, mode); break;
          default: epilogue_tile<BN, VMMT_ACT_NONE>(g, tacc, sc, q, lane, m0, n0, mode); break;
        }
      }
      // this warp's tcgen05.ld of the accumulator are complete (wait::ld inside the loads): hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8 * acc) : "memory");
      if (lt == 0 && threadIdx.x == 64) trace_mark(g.trace, 5);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
  }
  if (threadIdx.x == 64) trace_mark(g.trace, 6);
}

// C = bias broadcast (or 0): pre-pass for split-K without accumulate
__global__ void tc_init_bias_kernel(float* C, int64_t ldc, int M, int N, const float* bias) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  C[(int64_t)m * ldc + n] = bias ? bias[n] : 0.0f;
}

// C = act(C) in place: finishing pass of a split-K GEMM with a non-linear epilogue
__global__ void tc_finish_act_kernel(float* C, int64_t ldc, int M, int N, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  float* p = C + (int64_t)(i / N) * ldc + (i % N);
  *p = act_apply(*p, act);
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  });
  return g_encode;
}

// tensor map over a row-major fp32 matrix [outer, inner] with row stride ld (elements); box {32, box_rows}
int make_map(CUtensorMap* tm, const float* p, int64_t inner, int64_t outer, int64_t ld, int box_rows, bool mn_major,
             bool plain_f32 = false, bool bf16 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    vmmt_set_error("gemm_tc: cuTensorMapEncodeTiled is unavailable");
    return VMMT_ELAUNCH;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {bf16 ? 64u : 32u, (cuuint32_t)box_rows};          // 128 bytes of the inner dimension
  cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                      : (plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32);
  const CUresult r = enc(tm, dt, 2, const_cast<float*>(p), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         (mn_major && !bf16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vmmt_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)r,
                   (long long)inner, (long long)outer, (long long)ld);
    return VMMT_ELAUNCH;
  }
  return VMMT_OK;
}

struct Maps { CUtensorMap a, b, c, a2, b2; };

template <int BN, int STAGES, bool A_MN, bool B_MN, bool GEN = false, bool BF = false>
int launch(const Maps& m, const GemmArgs& g, dim3 grid, cudaStream_t s) {
  auto kern = gemm_tf32_kernel<BN, STAGES, A_MN, B_MN, GEN, BF>;
  static bool attr_done = false;        // per instantiation
  constexpr size_t smem = smem_bytes(BN, STAGES);
  if (!attr_done) {
    VMMT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  kern<<<grid, THREADS, smem, s>>>(m.a, m.b, m.c, m.a2, m.b2, g);
  return vmmt_check_launch("gemm_tf32_kernel");
}

template <int BN, int STAGES, bool BF = false>
int launch_major(bool a_mn, bool b_mn, const Maps& m, const GemmArgs& g, dim3 grid, cudaStream_t s) {
  if (g.epi_mode != 0) {                 // generator epilogues: x [M,H] and W [V,H] are both K-major
    if (a_mn || b_mn || BN != 128) return VMMT_EINVAL;
    return launch<128, STAGES, false, false, true, BF>(m, g, grid, s);
  }
  if (!a_mn && !b_mn) return launch<BN, STAGES, false, false, false, BF>(m, g, grid, s);
  if (!a_mn && b_mn) return launch<BN, STAGES, false, true, false, BF>(m, g, grid, s);
  if (a_mn && !b_mn) return launch<BN, STAGES, true, false, false, BF>(m, g, grid, s);
  return launch<BN, STAGES, true, true, false, BF>(m, g, grid, s);
}

}  // namespace

bool vmmt_gemm_tc_eligible(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor,
                           const float* C, int64_t ldc, int M, int N, int K, int flags) {
  (void)C; (void)ldc; (void)a_kmajor; (void)b_kmajor;
  if (M < 1 || N < 1 || K < 1) return false;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return false;      // TMA: 16-byte aligned base ...
  if ((lda & 3) || (ldb & 3)) return false;                          // ... and 16-byte multiple row pitch
  // very thin outputs / contractions stay on the SIMT kernel (a 128-row MMA tile would be almost empty);
  // skinny-M problems (the batch-row MLPs of the latent / image networks) are weight-streaming: they run here
  // with split-K over all SMs and the TMA reduce-add epilogue.
  // (K >= 16: the K = batch-size weight-gradient products dW += dY^T X of the batch-row MLPs are 1-2 k-blocks deep but
  //  up to 2048 x 2048 wide -- 40+ us each on the SIMT kernel, a few us here)
  if (N < 64 || K < 16) return false;
  // batch-invariant calls (VMMT_F_NO_SPLITK): which kernel runs must not depend on the row count either, or a sentence
  // decoded alone (M = beam rows) would take the exact SIMT kernel and the same sentence in a batch the TF32 one
  if (!(flags & VMMT_F_NO_SPLITK) && (int64_t)M * N * K < (int64_t)64 * 64 * 64) return false;
  return true;
}

int vmmt_gemm_tc(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor, float* C,
                 int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate, int flags, cudaStream_t s) {
  return vmmt_gemm_tc_ex(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, bias, act, accumulate, nullptr, flags, s);
}

int vmmt_gemm_tc_ex(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor, float* C,
                    int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate,
                    const VmmtGenEpi* epi, int flags, cudaStream_t s) {
  return vmmt_gemm_tc_dual(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, bias, act, accumulate, epi, nullptr, flags, s);
}

int vmmt_gemm_tc_dual(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor, float* C,
                      int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate,
                      const VmmtGenEpi* epi, const VmmtGemmSecond* second, int flags, cudaStream_t s) {
  const bool a_mn = !a_kmajor, b_mn = !b_kmajor;
  const bool background = (flags & VMMT_F_BACKGROUND) != 0;
  const bool share = (flags & VMMT_F_SHARE_SMS) != 0;
  // VMMT_F_BF16 on this INTERNAL entry point: A / B (and the second pair) point to bf16 data, lda / ldb count bf16
  // elements (vmmt_gemm_bf16 and the generator's bf16 path cast the fp32 tensors first)
  const bool bf = (flags & VMMT_F_BF16) != 0;
  const int bke = bf ? 2 * BK : BK;                            // elements per k-block
  const int nsm = vmmt_num_sms();
  const int tiles_m = ceil_div(M, BM);
  const int nkb2 = second ? ceil_div(second->K2, bke) : 0;
  if (second && (a_mn || b_mn)) return VMMT_EINVAL;            // the second pair shares the K-major instantiation
  const int nkb = ceil_div(K, bke) + nkb2;                       // k-blocks per output tile (cost model); split-K only without a second pair
  // (tile width, split-K) from a small cost model in SM cycles: a CTA costs a fixed prologue + epilogue plus
  // its k-blocks; 128-wide tiles run one CTA per SM (shared-memory-bandwidth bound, ~450 cycles per k-block),
  // 64-wide tiles two per SM; split-K needs a linear epilogue (an activation is applied by a finishing pass)
  // VMMT_F_NO_SPLITK: one accumulation chain per output element in fixed K order (deterministic; a row's result does not
  // depend on how many rows the call has: inference compares a sentence decoded alone with the same sentence in a batch)
  const bool can_split = !(flags & VMMT_F_NO_SPLITK) && !share && epi == nullptr && second == nullptr && ((accumulate == 0) || (accumulate == 1 && bias == nullptr && act == VMMT_ACT_NONE));
  int BN = 128, splits = 1;
  double best = 1e30;
  for (int bn = (epi ? 128 : 64); bn <= 128; bn *= 2) {
    const int tiles = tiles_m * ceil_div(N, bn);
    const int slots = nsm;
    const int max_split = can_split ? min(32, max(1, nkb / 4)) : 1;
    for (int sp = 1; sp <= max_split; ++sp) {
      const int kb = ceil_div(nkb, sp);
      if (sp > 1 && ceil_div(nkb, kb) != sp) continue;
      const int ctas = tiles * sp;
      const double ckb = (bn == 128) ? 450.0 : (ctas > nsm ? 600.0 : 320.0);
      double cost = (double)ceil_div(ctas, slots) * (5000.0 + kb * ckb);
      if (sp > 1) cost += 4000.0 + (act != VMMT_ACT_NONE ? 4000.0 : 0.0);
      if (cost < best) { best = cost; BN = bn; splits = sp; }
    }
  }
  const int kb_per = ceil_div(nkb, splits);
  const bool finish_act = splits > 1 && act != VMMT_ACT_NONE;
  if (splits > 1 && !accumulate) {
    const int64_t tot = (int64_t)M * N;
    tc_init_bias_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(C, ldc, M, N, bias);
    const int rc0 = vmmt_check_launch("gemm_tc_init_bias");
    if (rc0) return rc0;
  }
  CUtensorMap ta, tb;
  int rc;
  // K-major operand [rows,K]: inner = K, outer = rows, box {32 k, tile rows}
  // MN-major operand [K,rows]: inner = rows, outer = K,  box {32 rows, 32 k}
  rc = a_mn ? make_map(&ta, A, M, K, lda, bke, true, false, bf) : make_map(&ta, A, K, M, lda, BM, false, false, bf);
  if (rc) return rc;
  rc = b_mn ? make_map(&tb, B, N, K, ldb, bke, true, false, bf) : make_map(&tb, B, K, N, ldb, BN, false, false, bf);
  if (rc) return rc;
  static int dbg = getenv("VMMT_GEMM_DBG") ? atoi(getenv("VMMT_GEMM_DBG")) : 0;
  // C through TMA (store / reduce-add) when its base and pitch are 16-byte aligned; else direct stores
  const bool lse_mode = epi != nullptr && (epi->mode == 1 || epi->mode == 3);        // no C at all
  const int tma_store = (!lse_mode && ((uintptr_t)C & 15) == 0 && (ldc & 3) == 0 && accumulate != 2 &&
                         (epi != nullptr || !getenv("VMMT_GEMM_NO_TMA_STORE"))) ? 1 : 0;
  if (epi != nullptr && epi->mode == 2 && !tma_store) return VMMT_EINVAL;   // the gradient epilogue lives in the TMA-store path
  CUtensorMap tc;
  if (tma_store) {
    rc = make_map(&tc, C, N, M, ldc, 32, false, true);
    if (rc) return rc;
  } else {
    tc = ta;
  }
  Maps maps{ta, tb, tc, ta, tb};
  if (second) {
    rc = make_map(&maps.a2, second->A2, second->K2, M, second->lda2, BM, false, false, bf);
    if (rc) return rc;
    rc = make_map(&maps.b2, second->B2, second->K2, N, second->ldb2, BN, false, false, bf);
    if (rc) return rc;
  }
  GemmArgs g{C, ldc, M, N, K, bias, finish_act ? VMMT_ACT_NONE : act, accumulate, kb_per, tma_store, dbg, tiles_m, ceil_div(N, BN), nkb2, nullptr,
             0, nullptr, nullptr, nullptr, nullptr, nullptr, 1.0f, 0, 0, nullptr, nullptr, 16};
  {
    static const int forced = getenv("VMMT_GEMM_GROUP_M") ? atoi(getenv("VMMT_GEMM_GROUP_M")) : 0;
    const size_t panel64 = (size_t)64 * BM * (size_t)(K + (second ? second->K2 : 0)) * (bf ? 2 : 4);
    g.group_m = forced > 0 ? forced : (panel64 <= ((size_t)48 << 20) ? 64 : 16);
  }
  if (epi) {
    g.epi_mode = epi->mode;
    g.lse_part = reinterpret_cast<float4*>(epi->lse_part);
    g.tgt_logit = epi->tgt_logit;
    g.target = epi->target;
    g.row_lse = epi->row_lse;
    g.gscale = epi->gscale;
    g.scale = epi->scale;
    g.pad = epi->pad;
    g.topk = epi->topk;
    g.tile_lse = reinterpret_cast<float2*>(epi->tile_lse);
    g.tile_cand = reinterpret_cast<float2*>(epi->tile_cand);
    if (epi->mode == 3 && (epi->topk < 1 || epi->topk > VMMT_TOPK_MAX)) return VMMT_EINVAL;
  }
  // persistent: one CTA per SM slot when there are more tiles than slots (never with split-K: one tile per CTA)
  const int tiles = tiles_m * ceil_div(N, BN);
  const int slots = nsm;                       // one CTA per SM (EXCLUSIVE_SMEM)
  static const int no_persist = getenv("VMMT_GEMM_NO_PERSIST") ? 1 : 0;
  // background launches (weight gradients on the low-priority side streams) keep one tile per CTA: a persistent CTA holds
  // its SM for the whole GEMM, and a recurrence kernel launched meanwhile cannot place its 16-CTA clusters until the GEMM
  // ends (measured: +80..100 us on an encoder layer's backward, at random); short-lived CTAs drain within one tile time
  // and the higher-priority cluster launch gets the SMs.
  static const int share_slots = getenv("VMMT_GEMM_SHARE_SLOTS") ? atoi(getenv("VMMT_GEMM_SHARE_SLOTS")) : 0;
  dim3 grid(share ? min(tiles, share_slots > 0 ? share_slots : max(1, nsm * 5 / 18))
                  : ((splits > 1 || no_persist || background) ? tiles : min(tiles, slots)), 1, splits);
  static long long* trace_buf = nullptr;
  const bool tracing = getenv("VMMT_GEMM_TRACE") != nullptr;
  if (tracing) {
    if (!trace_buf) cudaMalloc(&trace_buf, sizeof(long long) * 16 * 65536);
    cudaMemsetAsync(trace_buf, 0, sizeof(long long) * 16 * 65536, s);
    g.trace = trace_buf;
  }
  struct TraceDump {
    bool on; long long* buf; dim3 grid; cudaStream_t s; int BN;
    ~TraceDump() {
      if (!on) return;
      cudaStreamSynchronize(s);
      const int n = min(65536u, grid.x * grid.y * grid.z);
      long long* h = (long long*)malloc(sizeof(long long) * 16 * n);
      cudaMemcpy(h, buf, sizeof(long long) * 16 * n, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[gemm trace] grid=%ux%ux%u BN=%d (cycles since CTA start: setup, first_full, mma_issued, tfull, epi_end, exit)\n", grid.x, grid.y, grid.z, BN);
      for (int c : {0, 1, n / 2, n - 1}) {
        long long* t = h + 16 * c;
        fprintf(stderr, "  cta %5d: %lld %lld %lld %lld %lld %lld | epi chunk0: ld %lld sts %lld rows %lld\n", c, t[1] - t[0], t[2] - t[0], t[3] - t[0], t[4] - t[0], t[5] - t[0], t[6] - t[0], t[8] - t[4], t[9] - t[4], t[10] - t[4]);
      }
      free(h);
    }
  } dump{tracing, trace_buf, grid, s, BN};
  static const int deep = getenv("VMMT_GEMM_DEEP") ? atoi(getenv("VMMT_GEMM_DEEP")) : 0;
  if (bf)
    rc = (BN == 128) ? launch_major<128, 5, true>(a_mn, b_mn, maps, g, grid, s)
                     : launch_major<64, 3, true>(a_mn, b_mn, maps, g, grid, s);
  else if (deep)
    rc = (BN == 128) ? launch_major<128, 6>(a_mn, b_mn, maps, g, grid, s)
                     : launch_major<64, 8>(a_mn, b_mn, maps, g, grid, s);
  else
    rc = (BN == 128) ? launch_major<128, 5>(a_mn, b_mn, maps, g, grid, s)
                     : launch_major<64, 3>(a_mn, b_mn, maps, g, grid, s);
  if (rc) return rc;
  if (finish_act) {
    const int64_t tot = (int64_t)M * N;
    tc_finish_act_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(C, ldc, M, N, act);
    return vmmt_check_launch("gemm_tc_finish_act");
  }
  return VMMT_OK;
}


// ---------------------------------------------------------------------------------------------- bf16 variant
namespace {
// dst[r, 0..ld_dst) = bf16(src[r, 0..cols)), zero padded to ld_dst (a multiple of 8: 16-byte row pitch for the TMA)
__global__ void cast_bf16_kernel(const float* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst,
                                 int64_t ld_dst, int rows, int cols) {
  const int64_t per = ld_dst / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * per) return;
  const int r = (int)(i / per), c0 = (int)(i % per) * 8;
  const float* s = src + (int64_t)r * ld_src + c0;
  float v[8];
  if (c0 + 8 <= cols && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cols) ? __ldg(s + j) : 0.f;
  }
  __align__(16) __nv_bfloat162 o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(dst + (int64_t)r * ld_dst + c0) = *reinterpret_cast<const uint4*>(o);
}
}  // namespace

extern "C" int vmmt_cast_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int rows, int cols, void* stream) {
  VMMT_REQUIRE(src && dst && rows >= 0 && cols >= 0, "cast_bf16: bad arguments");
  VMMT_REQUIRE(ld_dst % 8 == 0 && ld_dst >= cols && ((uintptr_t)dst & 15) == 0,
               "cast_bf16: the destination pitch must be a multiple of 8 elements >= cols and the base 16-byte aligned");
  if (rows == 0 || cols == 0) return VMMT_OK;
  const int64_t n = (int64_t)rows * (ld_dst / 8);
  cast_bf16_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols);
  return vmmt_check_launch("cast_bf16_kernel");
}

// C[M,N] (fp32) = act(op(A) op(B) + bias) (+C) on bf16 operands (fp32 accumulate in tensor memory); same contract as
// vmmt_gemm, A / B are bf16 matrices with 16-byte aligned bases and pitches of a multiple of 8 elements.
extern "C" int vmmt_gemm_bf16(const void* A, int64_t lda, int a_kmajor, const void* B, int64_t ldb, int b_kmajor, float* C,
                              int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate, int flags,
                              void* stream) {
  VMMT_REQUIRE(M >= 1 && N >= 1 && K >= 1 && A && B && C, "gemm_bf16: bad arguments");
  VMMT_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && (lda & 7) == 0 && (ldb & 7) == 0,
               "gemm_bf16: operands need 16-byte aligned bases and pitches of a multiple of 8 elements");
  return vmmt_gemm_tc_dual(reinterpret_cast<const float*>(A), lda, a_kmajor, reinterpret_cast<const float*>(B), ldb, b_kmajor,
                           C, ldc, M, N, K, bias, act, accumulate, nullptr, nullptr, flags | VMMT_F_BF16,
                           (cudaStream_t)stream);
}
