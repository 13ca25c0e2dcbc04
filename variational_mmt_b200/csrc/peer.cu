// Data-parallel optimiser step over NVLink peer memory: reduce-scatter of the flat gradient buffers,
// global-norm clip + Adam on the rank's own 1/N slice, all-gather of the updated parameters -- the
// exchange and the update are the SAME kernels (no NCCL on the data path).
//
// Reference semantics: N ranks == the reference's `-accum_count N`
// step (onmt/TrainerMultimodal.py:342-346, 625-718): gradients of the N batches ADD, then ONE
// clip_grad_norm(5) + Adam(eps 1e-9) update (onmt/Optim.py:69-70, 94-96).  Here rank r
//   1. barrier  (every rank's backward has finished writing its flat gradient buffer)
//   2. K_rs     reads slice r of all N gradient buffers (its own from HBM, N-1 through NVLink P2P
//               loads), adds them in rank order 0..N-1 (bitwise the same sum whoever computes it),
//               keeps the reduced slice locally and publishes its share of ||g||^2 to every peer
//   3. barrier  (all N partial norms have landed; all peers have finished reading my gradients)
//   4. K_adam   total = sum of the N partial norms in rank order (identical on every rank), clip
//               coefficient, Adam on slice r (moments exist only for the slice: 8 B/param/N), and the
//               new parameter values are stored into all N parameter buffers (P2P stores)
//   5. barrier  (every slice of my parameter buffer has been written by its owner)
// Per rank and step NVLink carries (N-1)/N * 4 B/param in and the same out -- the reduce-scatter +
// all-gather lower bound -- and the Adam pass touches 1/N of the optimiser state.
//
// NVLS (default when the host maps the segments through torch's symmetric memory: cuMem + cuMulticast): steps 2 and 4
// use multimem.ld_reduce / multimem.st on the segment's multicast address -- the NVSwitch forms the sum of the N
// gradient float4s and replicates the parameter store, a rank moves 4 B/param/N each way.  The P2P kernels below stay
// as the path for hosts without multicast support.
//
// Peer memory is plain cudaMalloc memory exported with cudaIpcGetMemHandle (one process per GPU);
// the 64-byte handles travel over whatever control plane the host has (torch.distributed object
// all-gather in variational_mmt_b200/distributed.py).  Cross-GPU barriers are flag words in that
// memory written with st.release.sys and polled with ld.acquire.sys; the barrier generation is a
// device-resident counter, so the whole sequence is CUDA-graph capturable.
#include "common.cuh"
#include "vmmt_internal.h"

namespace {

constexpr int kMaxRanks = 16;
// layout of the signal block at the head of every rank's peer segment (bytes)
//   [0,   64)   arrive[kMaxRanks]  u32, slot j written by rank j
//   [64,  68)   epoch              u32, local barrier generation
//   [128, 256)  norm2[0][kMaxRanks]  f64, slot j written by rank j (its share of ||g||^2 of exchange phase 0)
//   [256, 384)  norm2[1][kMaxRanks]  same for phase 1 (the early, overlapped reduce-scatter of a sub-range)
//   [384, 448)  arrive[kMaxRanks] of barrier CHANNEL 1, [448, 452) its epoch: barriers issued from a second stream
//               (the all-gather of the buffer's tail, overlapped with the next step's forward pass) must not share
//               the generation counter of the main stream's barriers
constexpr int kSigBytes = 512;
constexpr int kNormOff = 128;

struct PeerPtrs {
  void* p[kMaxRanks];
};

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

// One block.  Thread j < world signals rank j and waits for rank j.
__global__ void peer_barrier_kernel(PeerPtrs sig, int rank, int world, int channel) {
  const int aw = channel ? 96 : 0, ew = channel ? 112 : 16;       // word offsets of the channel's arrive[] / epoch
  uint32_t* mine = reinterpret_cast<uint32_t*>(sig.p[rank]);
  const uint32_t e = mine[ew] + 1;                       // epoch
  __syncthreads();
  if (threadIdx.x == 0) mine[ew] = e;
  const int j = threadIdx.x;
  if (j < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(sig.p[j]) + aw + rank, e);
    // a peer is at most one generation ahead of me, so ">= e" (wrap-safe) is the arrival test
    while ((int32_t)(ld_acquire_sys(mine + aw + j) - e) < 0) __nanosleep(40);
  }
  __syncthreads();
}

// Reduce-scatter + squared norm of the reduced slice.  grads.p[j] = base of rank j's flat gradient
// buffer; the slice is [lo4, hi4) in float4 units.
template <int W>
__global__ void __launch_bounds__(256)
peer_reduce_scatter_kernel(PeerPtrs grads, int world, int64_t lo4, int64_t hi4,
                           float4* __restrict__ gsum, double* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  const int nw = (W > 0) ? W : world;
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v[(W > 0) ? W : kMaxRanks];
#pragma unroll
    for (int j = 0; j < ((W > 0) ? W : kMaxRanks); ++j)
      if (j < nw) v[j] = reinterpret_cast<const float4*>(grads.p[j])[i];
    float4 a = v[0];
#pragma unroll
    for (int j = 1; j < ((W > 0) ? W : kMaxRanks); ++j)
      if (j < nw) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
    gsum[i - lo4] = a;
    s = fmaf(a.x, a.x, s); s = fmaf(a.y, a.y, s); s = fmaf(a.z, a.z, s); s = fmaf(a.w, a.w, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = (double)s;
}

// Sums the block partials and publishes this rank's share of ||g||^2 into slot `rank` of every peer.
__global__ void peer_norm_publish_kernel(const double* __restrict__ partial, int nblk, PeerPtrs sig,
                                         int rank, int world, int phase) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if ((int)threadIdx.x < world) {
    double* slot = reinterpret_cast<double*>(reinterpret_cast<char*>(sig.p[threadIdx.x]) + kNormOff) +
                   phase * kMaxRanks + rank;
    *reinterpret_cast<volatile double*>(slot) = sh[0];
    __threadfence_system();
  }
}

// Clip + Adam on the slice, new parameters stored to every rank's buffer.
template <int W>
__global__ void __launch_bounds__(256)
peer_adam_allgather_kernel(PeerPtrs params, int world, int rank, int64_t lo4, int64_t hi4,
                           const float4* __restrict__ gsum, float4* __restrict__ m4,
                           float4* __restrict__ v4, const double* __restrict__ norm2, int nphase,
                           float* __restrict__ sq_out, float max_norm, float step_size, float b1,
                           float b2, float inv_bc2_sqrt, float eps) {
  const int nw = (W > 0) ? W : world;
  double tot = 0.0;
  for (int ph = 0; ph < nphase; ++ph)                       // fixed order: identical on every rank
    for (int j = 0; j < nw; ++j) tot += norm2[ph * kMaxRanks + j];
  const float sq = (float)tot;
  if (blockIdx.x == 0 && threadIdx.x == 0 && sq_out) sq_out[0] = sq;
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.0f, max_norm / (sqrtf(sq) + 1e-6f));
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gg = gsum[i - lo4];
    float4 pp = reinterpret_cast<const float4*>(params.p[rank])[i];
    float4 mm = m4[i - lo4];
    float4 vv = v4[i - lo4];
    float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x = ga[e] * coef;
      ma[e] = b1 * ma[e] + (1.f - b1) * x;
      va[e] = b2 * va[e] + (1.f - b2) * x * x;
      pa[e] -= step_size * ma[e] / (sqrtf(va[e]) * inv_bc2_sqrt + eps);
    }
    m4[i - lo4] = mm;
    v4[i - lo4] = vv;
#pragma unroll
    for (int j = 0; j < ((W > 0) ? W : kMaxRanks); ++j)
      if (j < nw) reinterpret_cast<float4*>(params.p[j])[i] = pp;
  }
}

// ---- NVLS forms (mc = the segment's MULTICAST address, cuMulticast* mapping over all N ranks' segments): the NVSwitch
// adds the N ranks' float4s in flight (multimem.ld_reduce) and replicates one store to all N buffers (multimem.st), so a
// rank moves its 1/N slice once in each direction -- 4 B/param/N in and out instead of (N-1)/N * 4 B/param.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc_addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256)
peer_reduce_scatter_mc_kernel(const float4* __restrict__ mc_grads, int64_t lo4, int64_t hi4, float4* __restrict__ gsum,
                              double* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  // a switch round trip per load: four independent multimem loads in flight per thread
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += 4 * stride) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < hi4) a[u] = multimem_ld_reduce_add(mc_grads + i + u * stride);   // sum over the N ranks, formed in the switch
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < hi4) {
        gsum[i + u * stride - lo4] = a[u];
        s = fmaf(a[u].x, a[u].x, s); s = fmaf(a[u].y, a[u].y, s); s = fmaf(a[u].z, a[u].z, s); s = fmaf(a[u].w, a[u].w, s);
      }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = (double)s;
}

__global__ void __launch_bounds__(256)
peer_adam_allgather_mc_kernel(const float4* __restrict__ my_params, float4* __restrict__ mc_params, int world, int64_t lo4,
                              int64_t hi4, const float4* __restrict__ gsum, float4* __restrict__ m4, float4* __restrict__ v4,
                              const double* __restrict__ norm2, int nphase, float* __restrict__ sq_out, float max_norm,
                              float step_size, float b1, float b2, float inv_bc2_sqrt, float eps) {
  double tot = 0.0;
  for (int ph = 0; ph < nphase; ++ph)                       // fixed order: identical on every rank
    for (int j = 0; j < world; ++j) tot += norm2[ph * kMaxRanks + j];
  const float sq = (float)tot;
  if (blockIdx.x == 0 && threadIdx.x == 0 && sq_out) sq_out[0] = sq;
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.0f, max_norm / (sqrtf(sq) + 1e-6f));
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gg = gsum[i - lo4];
    float4 pp = my_params[i];
    float4 mm = m4[i - lo4];
    float4 vv = v4[i - lo4];
    float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x = ga[e] * coef;
      ma[e] = b1 * ma[e] + (1.f - b1) * x;
      va[e] = b2 * va[e] + (1.f - b2) * x * x;
      pa[e] -= step_size * ma[e] / (sqrtf(va[e]) * inv_bc2_sqrt + eps);
    }
    m4[i - lo4] = mm;
    v4[i - lo4] = vv;
    multimem_st(mc_params + i, pp);                         // one store, replicated by the switch into all N buffers
  }
}

int fill(PeerPtrs* out, void* const* in, int world, size_t byte_off) {
  for (int j = 0; j < kMaxRanks; ++j)
    out->p[j] = (j < world) ? (void*)((char*)in[j] + byte_off) : nullptr;
  return 0;
}

}  // namespace

extern "C" size_t vmmt_peer_signal_bytes(void) { return kSigBytes; }
extern "C" int vmmt_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int vmmt_peer_alloc(size_t bytes, void** ptr, void* handle_out) {
  VMMT_REQUIRE(ptr && handle_out && bytes >= (size_t)kSigBytes, "peer_alloc: bad arguments");
  VMMT_CUDA(cudaMalloc(ptr, bytes));
  VMMT_CUDA(cudaMemset(*ptr, 0, bytes));
  VMMT_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  VMMT_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle_out, &h, sizeof(h));
  return VMMT_OK;
}

extern "C" int vmmt_peer_open(const void* handle, void** ptr) {
  VMMT_REQUIRE(handle && ptr, "peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  VMMT_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VMMT_OK;
}

extern "C" int vmmt_peer_close(void* ptr) {
  VMMT_CUDA(cudaIpcCloseMemHandle(ptr));
  return VMMT_OK;
}

extern "C" int vmmt_peer_free(void* ptr) {
  VMMT_CUDA(cudaFree(ptr));
  return VMMT_OK;
}

extern "C" int vmmt_peer_barrier(void* const* segments, int rank, int world, int channel, void* stream) {
  VMMT_REQUIRE(segments && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world,
               "peer_barrier: bad rank/world");
  VMMT_REQUIRE(channel == 0 || channel == 1, "peer_barrier: channel must be 0 or 1");
  PeerPtrs sig;
  fill(&sig, segments, world, 0);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sig, rank, world, channel);
  return vmmt_check_launch("peer_barrier");
}

extern "C" size_t vmmt_peer_adam_workspace_bytes(void) { return 2048 * sizeof(double); }

// [*lo, *hi) = the floats of the n-float range starting at `base` that rank `rank` reduces and updates
static int64_t range_slice(int64_t base, int64_t n, int world, int rank, int64_t* lo, int64_t* hi) {
  const int64_t n4 = n / 4;
  const int64_t per = (n4 + world - 1) / world;
  int64_t a = per * rank, b = per * (rank + 1);
  if (a > n4) a = n4;
  if (b > n4) b = n4;
  if (lo) *lo = base + a * 4;
  if (hi) *hi = base + b * 4;
  return per * 4;                       // capacity (floats) every rank allocates for its slice state
}

extern "C" int64_t vmmt_peer_slice(int64_t n, int world, int rank, int64_t* lo, int64_t* hi) {
  return range_slice(0, n, world, rank, lo, hi);
}

// barrier -> reduce-scatter of the sub-range [begin, begin + n) of the flat gradient buffers (this rank's slice of
// it) -> this rank's share of ||g||^2 published into norm slot array `phase` (0 or 1) of every peer.
// Two phases let a step exchange a sub-range EARLY (beside the rest of the backward pass) and the remainder at the end.
extern "C" int vmmt_peer_reduce_scatter(void* const* segments, void* mc_base, size_t grad_off, int rank, int world,
                                        int64_t begin, int64_t n, float* gsum, int phase, void* workspace, void* stream) {
  VMMT_REQUIRE(segments && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world,
               "peer_reduce_scatter: bad rank/world");
  VMMT_REQUIRE(n > 0 && n % 4 == 0 && begin >= 0 && begin % 4 == 0, "peer_reduce_scatter: range must be 16-byte granular");
  VMMT_REQUIRE(grad_off % 16 == 0 && ((uintptr_t)gsum & 15) == 0, "peer_reduce_scatter: buffers must be 16-byte aligned");
  VMMT_REQUIRE(phase == 0 || phase == 1, "peer_reduce_scatter: phase must be 0 or 1");
  cudaStream_t s = (cudaStream_t)stream;
  PeerPtrs sig, grads;
  fill(&sig, segments, world, 0);
  fill(&grads, segments, world, grad_off);
  int64_t lo, hi;
  range_slice(begin, n, world, rank, &lo, &hi);
  const int64_t lo4 = lo / 4, hi4 = hi / 4, cnt4 = hi4 - lo4;
  int rc;
  peer_barrier_kernel<<<1, 32, 0, s>>>(sig, rank, world, 0);
  if ((rc = vmmt_check_launch("peer_barrier"))) return rc;
  int nblk = ceil_div(cnt4 > 0 ? cnt4 : 1, 256 * 2);
  // phase 1 runs BESIDE the encoders' backward recurrences: two light blocks per SM leave room for their thread-block
  // clusters (a grid that fills every SM delays the cluster placement by its whole duration); phase 0 has the GPU alone
  const int cap = vmmt_num_sms() * (phase == 1 ? 2 : 8);
  if (nblk > cap) nblk = cap;
  if (nblk > 2048) nblk = 2048;
  if (nblk < 1) nblk = 1;
  double* partial = (double*)workspace;
  if (mc_base != nullptr && world > 1) {
    peer_reduce_scatter_mc_kernel<<<nblk, 256, 0, s>>>(reinterpret_cast<const float4*>((const char*)mc_base + grad_off), lo4, hi4,
                                                        (float4*)gsum, partial);
    if ((rc = vmmt_check_launch("peer_reduce_scatter_mc"))) return rc;
    peer_norm_publish_kernel<<<1, 256, 0, s>>>(partial, nblk, sig, rank, world, phase);
    return vmmt_check_launch("peer_norm_publish");
  }
  switch (world) {
#define RS_CASE(W)                                                                               \
  case W:                                                                                        \
    peer_reduce_scatter_kernel<W><<<nblk, 256, 0, s>>>(grads, world, lo4, hi4, (float4*)gsum, partial); \
    break;
    RS_CASE(1) RS_CASE(2) RS_CASE(4) RS_CASE(8)
#undef RS_CASE
    default:
      peer_reduce_scatter_kernel<0><<<nblk, 256, 0, s>>>(grads, world, lo4, hi4, (float4*)gsum, partial);
  }
  if ((rc = vmmt_check_launch("peer_reduce_scatter"))) return rc;
  peer_norm_publish_kernel<<<1, 256, 0, s>>>(partial, nblk, sig, rank, world, phase);
  return vmmt_check_launch("peer_norm_publish");
}

// [barrier ->] clip (total norm = sum over `nphase` slot arrays) + Adam on this rank's slice of [begin, begin + n),
// new parameters stored into all N parameter buffers [-> barrier].  The barrier before is needed once after the last
// reduce-scatter of a step, the barrier after once after the last update.
extern "C" int vmmt_peer_adam_allgather(void* const* segments, void* mc_base, size_t param_off, int rank, int world,
                                        int64_t begin, int64_t n, const float* gsum, float* exp_avg, float* exp_avg_sq,
                                        float* sqnorm_out, int nphase, float max_norm, float lr, float beta1, float beta2,
                                        float eps, int64_t step, int barrier_before, int barrier_after, int channel,
                                        void* stream) {
  VMMT_REQUIRE(segments && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world,
               "peer_adam_allgather: bad rank/world");
  VMMT_REQUIRE(n > 0 && n % 4 == 0 && begin >= 0 && begin % 4 == 0, "peer_adam_allgather: range must be 16-byte granular");
  VMMT_REQUIRE(param_off % 16 == 0, "peer_adam_allgather: offsets must be 16-byte aligned");
  VMMT_REQUIRE(step >= 1 && (nphase == 1 || nphase == 2), "peer_adam_allgather: bad step / nphase");
  VMMT_REQUIRE(channel == 0 || channel == 1, "peer_adam_allgather: channel must be 0 or 1");
  VMMT_REQUIRE((((uintptr_t)gsum | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
               "peer_adam_allgather: slice buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  PeerPtrs sig, params;
  fill(&sig, segments, world, 0);
  fill(&params, segments, world, param_off);
  int64_t lo, hi;
  range_slice(begin, n, world, rank, &lo, &hi);
  const int64_t lo4 = lo / 4, hi4 = hi / 4, cnt4 = hi4 - lo4;
  int rc;
  if (barrier_before) {
    peer_barrier_kernel<<<1, 32, 0, s>>>(sig, rank, world, channel);
    if ((rc = vmmt_check_launch("peer_barrier"))) return rc;
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const double* norm2 = reinterpret_cast<const double*>((const char*)segments[rank] + kNormOff);
  // channel 1 = the overlapped all-gather of the tail: it runs BESIDE the next step's encoder recurrences, whose
  // thread-block clusters need whole GPC slices free at once -- one light block per SM (no shared / tensor memory: it
  // co-resides with a recurrence or GEMM CTA) instead of a grid that floods the block scheduler
  int ablk = ceil_div(cnt4 > 0 ? cnt4 : 1, 256);
  if (channel == 1) ablk = min(ablk, vmmt_num_sms());
  if (mc_base != nullptr && world > 1) {
    peer_adam_allgather_mc_kernel<<<ablk, 256, 0, s>>>(
        reinterpret_cast<const float4*>((const char*)segments[rank] + param_off),
        reinterpret_cast<float4*>((char*)mc_base + param_off), world, lo4, hi4, (const float4*)gsum, (float4*)exp_avg,
        (float4*)exp_avg_sq, norm2, nphase, sqnorm_out, max_norm, step_size, beta1, beta2, inv_bc2_sqrt, eps);
  } else
  switch (world) {
#define AD_CASE(W)                                                                                \
  case W:                                                                                         \
    peer_adam_allgather_kernel<W><<<ablk, 256, 0, s>>>(params, world, rank, lo4, hi4, (const float4*)gsum, \
        (float4*)exp_avg, (float4*)exp_avg_sq, norm2, nphase, sqnorm_out, max_norm, step_size, beta1, beta2, \
        inv_bc2_sqrt, eps);                                                                       \
    break;
    AD_CASE(1) AD_CASE(2) AD_CASE(4) AD_CASE(8)
#undef AD_CASE
    default:
      peer_adam_allgather_kernel<0><<<ablk, 256, 0, s>>>(params, world, rank, lo4, hi4, (const float4*)gsum,
          (float4*)exp_avg, (float4*)exp_avg_sq, norm2, nphase, sqnorm_out, max_norm, step_size, beta1, beta2,
          inv_bc2_sqrt, eps);
  }
  if ((rc = vmmt_check_launch("peer_adam_allgather"))) return rc;
  if (barrier_after) {
    peer_barrier_kernel<<<1, 32, 0, s>>>(sig, rank, world, channel);
    return vmmt_check_launch("peer_barrier");
  }
  return VMMT_OK;
}

// The whole flat buffer in one go: barrier -> reduce-scatter + norm -> barrier -> clip + Adam + all-gather -> barrier.
extern "C" int vmmt_peer_adam_step(void* const* segments, void* mc_base, size_t param_off, size_t grad_off, int rank,
                                   int world, int64_t n, float* gsum, float* exp_avg,
                                   float* exp_avg_sq, float* sqnorm_out, float max_norm, float lr,
                                   float beta1, float beta2, float eps, int64_t step, void* workspace,
                                   void* stream) {
  int rc = vmmt_peer_reduce_scatter(segments, mc_base, grad_off, rank, world, 0, n, gsum, 0, workspace, stream);
  if (rc) return rc;
  return vmmt_peer_adam_allgather(segments, mc_base, param_off, rank, world, 0, n, gsum, exp_avg, exp_avg_sq, sqnorm_out, 1,
                                  max_norm, lr, beta1, beta2, eps, step, 1, 1, 0, stream);
}
