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
#include <cuda.h>
#include <mutex>
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                    // fp32 elements per stage along K = one 128-byte swizzle row
constexpr int UMMA_K = 8;                 // tf32: 32 bytes per instruction along K
constexpr int EPI_WARPS = 4;
constexpr int THREADS = 32 * (2 + EPI_WARPS);
constexpr int A_STAGE_BYTES = BM * BK * 4;               // 16 KB

__host__ __device__ constexpr int b_stage_bytes(int BN) { return BN * BK * 4; }
__host__ __device__ constexpr int epi_bytes() { return EPI_WARPS * 32 * 33 * 4; }
__host__ __device__ constexpr size_t smem_bytes(int BN, int STAGES) {
  return 1024 /*align slack*/ + (size_t)STAGES * (A_STAGE_BYTES + b_stage_bytes(BN)) + epi_bytes() + 256;
}

// ---------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

struct GemmArgs {
  float* C;
  int64_t ldc;
  int M, N, K;
  const float* bias;
  int act, accumulate;
  int kb_per_split;      // k-blocks (of BK) per gridDim.z slice
};

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int B_STAGE = b_stage_bytes(BN);
  constexpr int STAGE = A_STAGE_BYTES + B_STAGE;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B atoms are 1024-B aligned
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* epi = reinterpret_cast<float*>(gen + (size_t)STAGES * STAGE);
  const uint32_t bar0 = base + STAGES * STAGE + epi_bytes();
  // barriers: full[s] at bar0 + 8 s, empty[s] at bar0 + 8 (STAGES + s), tmem_full after them
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * STAGES, tfull = bar0 + 16 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (size_t)STAGES * STAGE + epi_bytes() + 16 * STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb_total = (g.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int kb1 = min(nkb_total, kb0 + g.kb_per_split);
  const int nkb = kb1 - kb0;           // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // TMEM allocation (BN fp32 accumulator columns), address written to shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t fb = full0 + 8 * s;
        mbar_expect_tx(fb, STAGE);
        const uint32_t sa = base + s * STAGE, sb = sa + A_STAGE_BYTES;
        const int k = (kb0 + i) * BK;
        if (!A_MN) {
          tma_load_2d(sa, &tmA, k, m0, fb);                        // box {32 k, 128 rows}
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, m0 + 32 * j, k, fb);   // box {32 m, 32 k}
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmB, k, n0, fb);                        // box {32 k, BN rows}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, n0 + 32 * j, k, fb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * STAGE, sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < BK / UMMA_K; ++j) {
          const uint64_t ad = A_MN ? make_desc(sa + j * 1024, BK * 128, 512, 1) : make_desc(sa + j * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_desc(sb + j * 1024, BK * 128, 512, 1) : make_desc(sb + j * 32, 16, 1024);
          umma_tf32(tmem_base, ad, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * s);          // slot reusable once these MMAs have read it
      }
      umma_commit(tfull);                      // accumulator complete
    }
  } else {
    // ---------------- epilogue warps: TMEM lane quarter = warp % 4
    const int q = warp & 3;
    float* sc = epi + (warp - 2) * (32 * 33);
    mbar_wait(tfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool split = gridDim.z > 1;
    const int row_base = m0 + 32 * q;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int col = n0 + c * 32 + lane;
      if (n0 + c * 32 >= g.N) break;
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) sc[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const float bv = (g.bias != nullptr && col < g.N && !split) ? __ldg(g.bias + col) : 0.0f;
      if (col < g.N) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int row = row_base + r;
          if (row >= g.M) break;
          float x = sc[r * 33 + lane];
          float* cp = g.C + (int64_t)row * g.ldc + col;
          if (split) {
            atomicAdd(cp, x);
          } else {
            x += bv;
            if (g.accumulate == 2) x += *cp;
            x = act_apply(x, g.act);
            *cp = (g.accumulate == 1) ? (*cp + x) : x;
          }
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
  }
}

// C = bias broadcast (or 0): pre-pass for split-K without accumulate
__global__ void tc_init_bias_kernel(float* C, int64_t ldc, int M, int N, const float* bias) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  C[(int64_t)m * ldc + n] = bias ? bias[n] : 0.0f;
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
int make_map(CUtensorMap* tm, const float* p, int64_t inner, int64_t outer, int64_t ld, int box_rows, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    vmmt_set_error("gemm_tc: cuTensorMapEncodeTiled is unavailable");
    return VMMT_ELAUNCH;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(p), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vmmt_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)r,
                   (long long)inner, (long long)outer, (long long)ld);
    return VMMT_ELAUNCH;
  }
  return VMMT_OK;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, dim3 grid, cudaStream_t s) {
  auto kern = gemm_tf32_kernel<BN, STAGES, A_MN, B_MN>;
  static bool attr_done = false;        // per instantiation
  constexpr size_t smem = smem_bytes(BN, STAGES);
  if (!attr_done) {
    VMMT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  kern<<<grid, THREADS, smem, s>>>(ta, tb, g);
  return vmmt_check_launch("gemm_tf32_kernel");
}

template <int BN, int STAGES>
int launch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, dim3 grid,
                 cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, STAGES, false, false>(ta, tb, g, grid, s);
  if (!a_mn && b_mn) return launch<BN, STAGES, false, true>(ta, tb, g, grid, s);
  if (a_mn && !b_mn) return launch<BN, STAGES, true, false>(ta, tb, g, grid, s);
  return launch<BN, STAGES, true, true>(ta, tb, g, grid, s);
}

}  // namespace

bool vmmt_gemm_tc_eligible(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor,
                           const float* C, int64_t ldc, int M, int N, int K) {
  (void)C; (void)ldc; (void)a_kmajor; (void)b_kmajor;
  if (M < 1 || N < 1 || K < 1) return false;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return false;      // TMA: 16-byte aligned base ...
  if ((lda & 3) || (ldb & 3)) return false;                          // ... and 16-byte multiple row pitch
  if ((int64_t)M * N * K < (int64_t)64 * 64 * 64) return false;      // tiny problems: launch-bound either way
  return true;
}

int vmmt_gemm_tc(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor, float* C,
                 int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate, cudaStream_t s) {
  const bool a_mn = !a_kmajor, b_mn = !b_kmajor;
  const int nsm = vmmt_num_sms();
  // tile width: 128 when that still fills the machine, else 64
  const int tiles_m = ceil_div(M, BM);
  int BN = 128;
  if ((int64_t)tiles_m * ceil_div(N, 128) < nsm) BN = 64;
  const int tiles = tiles_m * ceil_div(N, BN);
  const int nkb = ceil_div(K, BK);
  // split-K when the tile grid leaves most SMs idle and the epilogue is linear
  int splits = 1;
  if (act == VMMT_ACT_NONE && accumulate != 2 && tiles * 2 <= nsm && nkb >= 8) {
    splits = min(min(nsm / tiles, nkb / 4), 32);
    if (splits < 1) splits = 1;
  }
  int kb_per = ceil_div(nkb, splits);
  splits = ceil_div(nkb, kb_per);
  if (splits > 1) {
    if (!accumulate) {
      const int64_t tot = (int64_t)M * N;
      tc_init_bias_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(C, ldc, M, N, bias);
      const int rc0 = vmmt_check_launch("gemm_tc_init_bias");
      if (rc0) return rc0;
    } else if (bias) {
      vmmt_set_error("vmmt_gemm_tc: bias with accumulate in split-K is unsupported");
      return VMMT_EINVAL;
    }
  }
  CUtensorMap ta, tb;
  int rc;
  // K-major operand [rows,K]: inner = K, outer = rows, box {32 k, tile rows}
  // MN-major operand [K,rows]: inner = rows, outer = K,  box {32 rows, 32 k}
  rc = a_mn ? make_map(&ta, A, M, K, lda, BK, true) : make_map(&ta, A, K, M, lda, BM, false);
  if (rc) return rc;
  rc = b_mn ? make_map(&tb, B, N, K, ldb, BK, true) : make_map(&tb, B, K, N, ldb, BN, false);
  if (rc) return rc;
  GemmArgs g{C, ldc, M, N, K, bias, act, accumulate, kb_per};
  dim3 grid(ceil_div(N, BN), tiles_m, splits);
  if (BN == 128) return launch_major<128, 5>(a_mn, b_mn, ta, tb, g, grid, s);
  return launch_major<64, 4>(a_mn, b_mn, ta, tb, g, grid, s);
}
