// Step-wise LSTM recurrence for LARGE batches / hidden sizes (e.g. the stress configuration B = 512, H = 1024):
// when the per-step contraction [N,H] x [H,4H] is a healthy GEMM by itself, every step is one tensor-core GEMM
// (vmmt_gemm: tcgen05 TF32, or exact fp32 in SIMT mode) plus one fused cell kernel; the persistent / cluster
// kernels (lstm.cu, lstm_tc.cu) cover the small-batch, latency-bound shapes.
//
// Reference semantics: torch nn.LSTM (onmt/Models.py:124-149, 892-893; onmt/VI_Model1.py:106), gate order
// i,f,g,o; rows are frozen past their length and their outputs are zero (what unpacking a packed sequence gives).
#include "common.cuh"
#include "vmmt_internal.h"
#include "lstm_tc.h"

extern "C" int vmmt_gemm(const float*, int64_t, int, const float*, int64_t, int, float*, int64_t, int, int, int,
                         const float*, int, int, int, void*);

namespace {

// gates = act(gpre + gx_t + b_ih + b_hh + rowbias); c' = f c + i g; h' = o tanh(c'); masked rows keep their state
__global__ void lstm_step_fwd_kernel(const float* __restrict__ gpre /*[N,4H] or null (first step, zero state)*/,
                                     const float* __restrict__ gx_t, const float* __restrict__ b_ih,
                                     const float* __restrict__ b_hh, const float* __restrict__ rowbias,
                                     float* __restrict__ h_state, float* __restrict__ c_state,
                                     const int64_t* __restrict__ lengths, int t, float* __restrict__ out_t,
                                     int64_t out_ld, float* __restrict__ gates_t, float* __restrict__ cs_t, int N, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H) return;
  const int n = i / H, u = i % H;
  float G[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const size_t j = (size_t)g * H + u, nj = (size_t)n * 4 * H + j;
    G[g] = (gpre ? gpre[nj] : 0.f) + gx_t[nj] + (b_ih ? b_ih[j] : 0.f) + (b_hh ? b_hh[j] : 0.f) +
           (rowbias ? rowbias[nj] : 0.f);
  }
  const float ig = sigmoidf_(G[0]), fg = sigmoidf_(G[1]), gg = tanhf(G[2]), og = sigmoidf_(G[3]);
  const float cn = fg * c_state[i] + ig * gg;
  const float hn = og * tanhf(cn);
  const bool m = lengths == nullptr || t < (int)lengths[n];
  if (m) { c_state[i] = cn; h_state[i] = hn; }
  out_t[(size_t)n * out_ld + u] = m ? hn : 0.f;
  if (gates_t) {
    float* gp = gates_t + (size_t)n * 4 * H + u;
    gp[0] = ig; gp[(size_t)H] = fg; gp[(size_t)2 * H] = gg; gp[(size_t)3 * H] = og;
  }
  if (cs_t) cs_t[i] = m ? cn : c_state[i];
}

// dG_t from (dh_rec + dout_t, dc); dc <- dct f; dh_rec <- 0 for live rows (the GEMM that follows accumulates
// dG_t W_hh into it), unchanged for frozen rows (their gradient passes through this step)
__global__ void lstm_step_bwd_kernel(const float* __restrict__ gates_t, const float* __restrict__ cs_t,
                                     const float* __restrict__ c_prev /*cs[t_prev] or c0 or null*/,
                                     const float* __restrict__ dout_t, int64_t dout_ld, float* __restrict__ dh_rec,
                                     float* __restrict__ dc, const int64_t* __restrict__ lengths, int t,
                                     float* __restrict__ dgates_t, int N, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H) return;
  const int n = i / H, u = i % H;
  const bool m = lengths == nullptr || t < (int)lengths[n];
  float dG[4] = {0.f, 0.f, 0.f, 0.f};
  if (m) {
    const float* gp = gates_t + (size_t)n * 4 * H + u;
    const float ig = gp[0], fg = gp[(size_t)H], gg = gp[(size_t)2 * H], og = gp[(size_t)3 * H];
    const float ct = cs_t[i];
    const float cp = c_prev ? c_prev[i] : 0.f;
    const float dh = dh_rec[i] + (dout_t ? dout_t[(size_t)n * dout_ld + u] : 0.f);
    const float tc = tanhf(ct);
    const float dct = dc[i] + dh * og * (1.f - tc * tc);
    dG[0] = dct * gg * ig * (1.f - ig);
    dG[1] = dct * cp * fg * (1.f - fg);
    dG[2] = dct * ig * (1.f - gg * gg);
    dG[3] = dh * tc * og * (1.f - og);
    dc[i] = dct * fg;
    dh_rec[i] = 0.f;
  }
  float* dg = dgates_t + (size_t)n * 4 * H + u;
  dg[0] = dG[0]; dg[(size_t)H] = dG[1]; dg[(size_t)2 * H] = dG[2]; dg[(size_t)3 * H] = dG[3];
}

__global__ void copy_or_zero_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src ? src[i] : 0.f;
}

int copy_or_zero(float* dst, const float* src, int64_t n, cudaStream_t s) {
  copy_or_zero_kernel<<<ceil_div(n, 256), 256, 0, s>>>(dst, src, n);
  return vmmt_check_launch("lstm_step_init");
}

}  // namespace

size_t vmmt_lstm_step_workspace_floats(int ndir, int N, int H) { return (size_t)ndir * 6 * N * H; }

int vmmt_lstm_step_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int flags,
                       float* ws, cudaStream_t s) {
  const int64_t NH = (int64_t)N * H;
  for (int d = 0; d < ndir; ++d) {
    const VmmtLstmDir& D = dirs[d];
    float* gpre = ws + (size_t)d * 6 * NH;             // [N,4H]
    float* hst = gpre + 4 * NH;                        // [N,H] running state
    float* cst = hst + NH;
    int rc = copy_or_zero(hst, D.h0, NH, s);
    if (rc) return rc;
    rc = copy_or_zero(cst, D.c0, NH, s);
    if (rc) return rc;
    for (int st = 0; st < T; ++st) {
      const int t = D.reverse ? T - 1 - st : st;
      const bool have_h = st > 0 || D.h0 != nullptr;
      if (have_h) {
        rc = vmmt_gemm(hst, H, 1, D.w_hh, H, 1, gpre, 4 * H, N, 4 * H, H, nullptr, VMMT_ACT_NONE, 0, flags, (void*)s);
        if (rc) return rc;
      }
      lstm_step_fwd_kernel<<<ceil_div(NH, 256), 256, 0, s>>>(
          have_h ? gpre : nullptr, D.gx + (size_t)t * N * 4 * H, D.b_ih, D.b_hh, D.rowbias, hst, cst, lengths, t,
          D.out + (size_t)t * N * D.out_ld, D.out_ld, D.gates ? D.gates + (size_t)t * N * 4 * H : nullptr,
          D.cs ? D.cs + (size_t)t * NH : nullptr, N, H);
      rc = vmmt_check_launch("lstm_step_fwd");
      if (rc) return rc;
    }
    if (D.hT) { rc = copy_or_zero(D.hT, hst, NH, s); if (rc) return rc; }
    if (D.cT) { rc = copy_or_zero(D.cT, cst, NH, s); if (rc) return rc; }
  }
  return VMMT_OK;
}

int vmmt_lstm_step_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int flags,
                       float* ws, cudaStream_t s) {
  const int64_t NH = (int64_t)N * H;
  for (int d = 0; d < ndir; ++d) {
    const VmmtLstmDirBwd& D = dirs[d];
    float* dh = ws + (size_t)d * 6 * NH;
    float* dc = dh + NH;
    int rc = copy_or_zero(dh, D.dhT, NH, s);
    if (rc) return rc;
    rc = copy_or_zero(dc, D.dcT, NH, s);
    if (rc) return rc;
    for (int st = 0; st < T; ++st) {
      const int t = D.reverse ? st : T - 1 - st;
      const int tp = D.reverse ? t + 1 : t - 1;
      const float* c_prev = (tp >= 0 && tp < T) ? D.cs + (size_t)tp * NH : D.c0;
      float* dg_t = D.dgates + (size_t)t * N * 4 * H;
      lstm_step_bwd_kernel<<<ceil_div(NH, 256), 256, 0, s>>>(
          D.gates + (size_t)t * N * 4 * H, D.cs + (size_t)t * NH, c_prev,
          D.dout ? D.dout + (size_t)t * N * D.dout_ld : nullptr, D.dout_ld, dh, dc, lengths, t, dg_t, N, H);
      rc = vmmt_check_launch("lstm_step_bwd");
      if (rc) return rc;
      if (st + 1 < T || D.dh0) {                        // dh_{t-1} += dG_t W_hh   ([N,4H] x [4H,H])
        rc = vmmt_gemm(dg_t, 4 * H, 1, D.w_hh, H, 0, dh, H, N, H, 4 * H, nullptr, VMMT_ACT_NONE, 1, flags, (void*)s);
        if (rc) return rc;
      }
    }
    if (D.dh0) { rc = copy_or_zero(D.dh0, dh, NH, s); if (rc) return rc; }
    if (D.dc0) { rc = copy_or_zero(D.dc0, dc, NH, s); if (rc) return rc; }
  }
  return VMMT_OK;
}
