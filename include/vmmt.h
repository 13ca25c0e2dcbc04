/* vmmt.h -- C ABI of libvmmt.so: the sm_100a kernels behind the VI-model-1 hot path.
 *
 * The reference (iacercalixto/variational_mmt) is pure Python over PyTorch and has no FFI of its own:
 * every entry point below replaces one or more torch library calls made by the reference modules
 * (cited per function as file:line relative to the reference root).  The Python host
 * (variational_mmt_b200/_lib.py) binds these with ctypes and passes tensor.data_ptr() values and the
 * current CUDA stream; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - all floating-point buffers are fp32 device memory owned by the caller (PyTorch); kernels never
 *     allocate or free; token ids / lengths are int64 device memory, as the reference holds them;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 *     synchronisation and is CUDA-graph capturable;
 *   - return value 0 = ok; > 0 = cudaError_t; < 0 = VMMT_E*; vmmt_last_error() has the message
 *     (thread-local).  There is no CPU fallback of any kind.
 *   - sequence tensors are time-major [T, B, *], the layout the reference's RNN code uses.
 */
#ifndef VMMT_H_
#define VMMT_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VMMT_ACT_NONE = 0, VMMT_ACT_RELU = 1, VMMT_ACT_TANH = 2, VMMT_ACT_SOFTPLUS = 3, VMMT_ACT_SIGMOID = 4 };

const char* vmmt_last_error(void);
int vmmt_version(void);
/* number of kernels libvmmt has launched in this process (bench.py's gpu_launches). */
unsigned long long vmmt_launch_count(void);

/* Per-call arithmetic / scheduling flags (`flags` arguments below).  The library keeps NO process-global mode: the only
 * state it holds are per-device caches (function attributes, SM counts) behind a mutex. */
#define VMMT_F_EXACT 1       /* exact-fp32 SIMT contractions / recurrences instead of tensor cores (parity debugging) */
#define VMMT_F_BF16 2        /* tensor-core contractions on bf16 operands (fp32 accumulate) instead of TF32 */
#define VMMT_F_BACKGROUND 4  /* optimiser-only work (weight gradients on a low-priority side stream): one tile per CTA instead
                                of the persistent per-SM tile loop, so that SMs free up for critical-path kernels between tiles */
#define VMMT_F_SHARE_SMS 16  /* a LARGE optimiser-only GEMM (more tiles than SMs) issued while short critical-path kernels run and no
                                recurrence cluster is waiting: persistent tile loop on 5/18 of the SMs (41 of 148), so that the critical chain
                                always finds free SMs (a grid that fills every SM makes later launches wait a whole tile time) */
#define VMMT_F_NO_SPLITK 8   /* one accumulation chain per output element in fixed K order: run-to-run deterministic and
                                batch-invariant (a row's result does not depend on how many rows the call has) */

/* C[M,N] (ldc) = act(op(A) op(B) + bias[N]) (+C if accumulate).
 * a_kmajor: A stored [M,K] (1) or [K,M] (0);  b_kmajor: B stored [N,K] (1, nn.Linear weight) or [K,N] (0).
 * Replaces nn.Linear / torch.mm / addmm on the path: GlobalAttention.py:71,78,113,188;
 * NormalVariationalEncoder.py:18-25,35-43; the input projections inside nn.LSTM (Models.py:124-129). */
int vmmt_gemm(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor,
              float* C, int64_t ldc, int M, int N, int K, const float* bias, int act, int accumulate,
              int flags, void* stream);

/* C[M,N] = act(A1 B1^T + A2 B2^T + bias): two operand pairs ([rows,K] row-major, the nn.Linear layout) contracted into one
 * accumulator in one launch.  Replaces the pairs of nn.Linear-on-a-concatenation the reference spells with torch.cat:
 * the LSTM cell's x W_ih^T + h W_hh^T (nn.LSTM, VI_Model1.py:106, one decode step) and linear_out([c ; q])
 * (GlobalAttention.py:187-190). */
int vmmt_gemm_dual(const float* A1, int64_t lda1, const float* B1, int64_t ldb1, int K1, const float* A2, int64_t lda2,
                   const float* B2, int64_t ldb2, int K2, float* C, int64_t ldc, int M, int N, const float* bias, int act,
                   int flags, void* stream);

/* bf16 variant (BASELINE configs[1] "fp32 and bf16"): dst[r, 0..ld_dst) = bf16(src[r, 0..cols)), zero padded to a pitch of a
 * multiple of 8 elements; vmmt_gemm_bf16 = vmmt_gemm's contract on such bf16 operands (tcgen05.mma.kind::f16, fp32
 * accumulate in tensor memory, fp32 output).  With VMMT_F_BF16 the generator entry points cast x and W into their
 * workspace themselves (the two products with the fp32 softmax gradient stay TF32). */
int vmmt_cast_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int rows, int cols, void* stream);
int vmmt_gemm_bf16(const void* A, int64_t lda, int a_kmajor, const void* B, int64_t ldb, int b_kmajor, float* C, int64_t ldc,
                   int M, int N, int K, const float* bias, int act, int accumulate, int flags, void* stream);

/* Embedding gather / dense scatter-add (Embeddings.py:169-188; nn.Embedding padding_idx row gets no grad). */
/* `rows` = rows of the table: an index outside [0, rows) traps on the device (nn.Embedding / numpy fancy indexing raise) */
int vmmt_embedding_fwd(const int64_t* idx, int64_t n, const float* table, int64_t rows, int E, float* out, void* stream);
int vmmt_embedding_bwd(const int64_t* idx, int64_t n, const float* dout, int E, int64_t pad_idx,
                       float* dtable, int64_t rows, void* stream);

/* ---- LSTM recurrence, one layer, one or two directions per launch (nn.LSTM: Models.py:124-149,
 * 892-893; VI_Model1.py:106).  gx = x W_ih^T (no bias) is computed by vmmt_gemm beforehand. */
typedef struct VmmtLstmDir {
  const float* gx;       /* [T,N,4H] */
  const float* w_hh;     /* [4H,H]  gate order i,f,g,o */
  const float* b_ih;     /* [4H] or NULL */
  const float* b_hh;     /* [4H] or NULL */
  const float* rowbias;  /* [N,4H] or NULL: per-example additive term (decoder: z W_ih[:,E:]^T) */
  const float* h0;       /* [N,H] or NULL (zeros) */
  const float* c0;       /* [N,H] or NULL */
  float* out;            /* out[(t*N+n)*out_ld + u]; zero past lengths[n] */
  int64_t out_ld;
  float* hT;             /* [N,H] or NULL: state at each row's last valid step */
  float* cT;
  float* gates;          /* [T,N,4H] activated gates saved for backward, or NULL (inference) */
  float* cs;             /* [T,N,H] cell states saved for backward, or NULL */
  int32_t reverse;       /* 1: run t = T-1 .. 0 */
  int32_t pad_;
} VmmtLstmDir;

typedef struct VmmtLstmDirBwd {
  const float* w_hh;     /* [4H,H] */
  const float* gates;    /* saved by forward */
  const float* cs;
  const float* c0;       /* or NULL */
  const float* dout;     /* dL/dout, dout[(t*N+n)*dout_ld + u], or NULL */
  int64_t dout_ld;
  const float* dhT;      /* [N,H] or NULL */
  const float* dcT;
  float* dgates;         /* [T,N,4H] OUT: gradient wrt the pre-activation gates (= d gx) */
  float* dh0;            /* [N,H] OUT or NULL */
  float* dc0;
  float* db_ih;          /* [4H] or NULL: bias gradient sum_{t,n} dgates ACCUMULATED (+=) by the launch -- only when */
  float* db_hh;          /* vmmt_lstm_seq_bwd_fuses_bias() says so (else ignored: use vmmt_colsum_acc on dgates) */
  float* drow;           /* [N,4H] or NULL: gradient of `rowbias` = sum_t dgates, WRITTEN under the same condition */
  int32_t reverse;
  int32_t pad_;
} VmmtLstmDirBwd;

size_t vmmt_lstm_workspace_bytes(int ndir, int N, int H);
int vmmt_lstm_seq_supported(int ndir, int N, int H);
/* cluster_budget: cap on the thread-block clusters this launch may occupy (0 = as many as are co-resident).  Two
 * independent recurrences issued on two streams (source encoder / target encoder) share the GPU with it. */
int vmmt_lstm_seq_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int flags,
                      int cluster_budget, void* workspace, size_t workspace_bytes, void* stream);
/* 1 when vmmt_lstm_seq_bwd with these arguments accumulates db_ih / db_hh itself (the tensor-core cluster recurrence) */
int vmmt_lstm_seq_bwd_fuses_bias(int ndir, int N, int H, int flags);
int vmmt_lstm_seq_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int flags,
                      int cluster_budget, void* workspace, size_t workspace_bytes, void* stream);
/* single-step cell on pre-summed gate pre-activations (decode with large sentences x beam). */
int vmmt_lstm_cell_fwd(const float* gates_pre, const float* b_ih, const float* b_hh, const float* rowbias,
                       const float* c_prev, float* h_out, float* c_out, int N, int H, void* stream);

/* ---- GlobalAttention core (GlobalAttention.py:108-113,169-184): qp = linear_in(q) [T,B,H],
 * ctx [S,B,H] -> align [T,B,S] (masked softmax), cvec [T,B,H].  S <= 128. */
int vmmt_attention_fwd(const float* qp, const float* ctx, const int64_t* lengths, float* align,
                       float* cvec, int T, int B, int S, int H, void* stream);
int vmmt_attention_bwd(const float* dcvec, const float* qp, const float* ctx, const float* align,
                       const int64_t* lengths, float* dscore_ws /*[T,B,S]*/, float* dqp, float* dctx,
                       int accumulate_dctx, int T, int B, int S, int H, void* stream);
/* the same in two launches: the query side (dscore, dqp: the decoder's backward chain waits for it) and the context side
 * (dctx: only the encoders' backward reads it), so that the caller may put the latter on another stream */
int vmmt_attention_bwd_query(const float* dcvec, const float* ctx, const float* align, const int64_t* lengths,
                             float* dscore_ws, float* dqp, int T, int B, int S, int H, void* stream);
int vmmt_attention_bwd_ctx(const float* dcvec, const float* qp, const float* align, const float* dscore_ws,
                           float* dctx, int accumulate_dctx, int T, int B, int S, int H, void* stream);

/* ---- inference networks (NormalVariationalEncoder.py:65-84, 12-43; Dists.py:21-26; VILoss.py:439-460) */
/* x[t,b,k] at x + t*stride_t + b*stride_b + k: time-major [T,B,H] (stride_t = B*H, stride_b = H) or the transposed view of a
 * [B,T,H] tensor (stride_t = H, stride_b = T*H: the target encoder's output, Models.py:905-911, without a transposing copy) */
int vmmt_masked_mean_fwd(const float* x, int64_t stride_t, int64_t stride_b, const int64_t* lengths, float* out,
                         int64_t out_ld, int T, int B, int H, void* stream);
int vmmt_masked_mean_bwd(const float* dout, int64_t dout_ld, const int64_t* lengths, float* dx, int64_t stride_t,
                         int64_t stride_b, int accumulate, int T, int B, int H, void* stream);
int vmmt_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act, void* stream);
/* stats8 = {nll, n_words, n_correct, kl, img_logprob, img_cos, -, -}: writes loss1[0] = stats8[7] = nll - img_logprob +
 * kl_weight * kl and stats8[6] = kl_weight * kl (VILoss.py:462-496) in one launch. */
int vmmt_loss_finalize(float* stats8, float kl_weight, float* loss1, void* stream);
/* out[N] (+= column sums of a[M,N]); out2 (optional) receives the same sums (nn.LSTM's b_ih / b_hh pair). */
int vmmt_colsum_acc(const float* a, int64_t lda, int M, int N, float* out, float* out2, void* stream);
int vmmt_axpy(float* y, const float* x, float alpha, int64_t n, void* stream);
/* Philox streams: effective offset = offset + (base ? *base : 0); `base` is a device-resident counter so that a
 * captured CUDA graph draws fresh masks / noise on every replay (advance it with vmmt_counter_add). */
int vmmt_counter_add(uint64_t* ctr, uint64_t inc, void* stream);
/* debug / measurement: buf[slot] = %globaltimer (ns) when the stream reaches this point (a one-thread kernel: phase
 * boundaries of a captured step without a profiler attached, tools/phase_stamps.py) */
int vmmt_stamp(uint64_t* buf, int slot, void* stream);
int vmmt_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, uint64_t offset,
                 const uint64_t* base /*device or NULL*/, void* stream);
int vmmt_normal_sample(const float* mu, const float* sd, const float* eps /*or NULL: Philox*/, float* z,
                       int64_t n, uint64_t seed, uint64_t offset, const uint64_t* base /*device or NULL*/,
                       void* stream);
int vmmt_kl_fwd(const float* mu_q, const float* sd_q, const float* mu_p /*NULL: 0*/,
                const float* sd_p /*NULL: 1*/, float* out1, int B, int Z, void* stream);
int vmmt_kl_bwd(const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p, float* dmu_q,
                float* dsd_q, float* dmu_p /*or NULL*/, float* dsd_p /*or NULL*/,
                const float* gscale /*device scalar or NULL*/, float scale, int B, int Z, void* stream);

/* ---- batch-row linear layers in exact fp32 (csrc/rowlin.cu): the location / scale MLPs of the prior, posterior and
 * image networks (NormalVariationalEncoder.py:12-43,93-110,164-228,286-304) have M = batch rows against K up to 3048:
 * weight-bandwidth bound, cluster split-K (deterministic), the same arithmetic per row whatever M is.
 *   w_transposed = 0: out_p[M,N] = act_p(x_p[M,K] W_p[N,K]^T + b_p)
 *   w_transposed = 1: out_p[M,N] = x_p[M,K] W_p[K,N]  with x_p := x_p * act'(y_p) applied on load (input gradient of a
 *                     layer whose activation output is y_p); xt_out, if set, receives that transformed x
 * x_p is the column-wise concatenation of nseg <= 3 matrices (segment i holds logical columns [k0_i, k0_{i+1})).
 * nprob = 2: two independent problems, or (sum_outputs) both accumulated into problem 0's output. */
typedef struct VmmtRowLinSeg {
  const float* p;
  int64_t ld;
  int32_t k0;
  int32_t pad_;
} VmmtRowLinSeg;
typedef struct VmmtRowLin {
  VmmtRowLinSeg seg[3];
  int32_t nseg;
  int32_t act;           /* activation of the output (forward form) */
  const float* y;        /* gradient form: activation output whose derivative scales x on load, or NULL */
  int64_t ldy;
  int32_t yact;
  int32_t pad_;
  const float* w;
  int64_t ldw;
  const float* bias;     /* [N] or NULL */
  float* out;
  int64_t ldo;
  float* xt_out;         /* [M,K] or NULL */
  int64_t ld_xt;
} VmmtRowLin;
int vmmt_rowlin(const VmmtRowLin* probs, int nprob, int sum_outputs, int w_transposed, int M, int N, int K,
                void* stream);

/* ---- image-feature head (NormalVariationalEncoder.py:286-299) and its loss (VILoss.py:22-56,317-332) */
int vmmt_gate_fwd(const float* z, const float* w, const float* bias, float* gate /*[B]*/, float* gated,
                  int B, int Z, void* stream);
int vmmt_gate_bwd(const float* dgated, const float* z, const float* gate, float* dpre_ws /*[B]*/, float* dw,
                  float* db, int B, int Z, void* stream);
int vmmt_image_loss_fwd(const float* loc, const float* v, float* rowstats /*[B,4]*/,
                        float* out2 /*{logprob, cosine}*/, int B, int D, void* stream);
int vmmt_image_loss_bwd(const float* loc, const float* v, const float* rowstats, float* dloc,
                        const float* gscale /*device scalar or NULL*/, float scale, int legacy_passthrough,
                        int B, int D, void* stream);

/* ---- generator + criterion (ModelConstructor.py:582-585; VILoss.py:228,243,515-531) */
size_t vmmt_generator_workspace_bytes(int M, int H, int V);
int vmmt_generator_nll_fwd(const float* x, const float* W, const float* b, const int64_t* target,
                           int64_t pad_idx, int M, int H, int V, float* lse /*[M]*/,
                           float* stats3 /*{nll_sum, n_words, n_correct}*/, void* workspace,
                           size_t workspace_bytes, int flags, void* stream);
int vmmt_generator_nll_bwd(const float* x, const float* W, const float* b, const int64_t* target,
                           int64_t pad_idx, const float* lse, const float* gscale /*device scalar or NULL*/,
                           float scale, int M, int H, int V,
                           float* dx /*or NULL*/, float* dW /*accumulated, or NULL*/, float* db /*accumulated, or NULL*/,
                           void* workspace, size_t workspace_bytes, int flags, void* stream);
/* dW += dlogits^T x, db += colsum(dlogits) from the dlogits vmmt_generator_nll_bwd left in `workspace` (when it was
 * called with dW = db = NULL): lets the host issue the weight gradient on another stream. */
int vmmt_generator_nll_wgrad(const float* x, const void* workspace, int M, int H, int V, float* dW, float* db,
                             int flags, void* stream);
int vmmt_generator_logprobs(const float* x, const float* W, const float* b, int M, int H, int V,
                            float* out /*[M,V]*/, float* lse_ws /*[M]*/, int flags, void* stream);

/* Beam-search form of the generator (TranslatorMultimodalVI.py:199 `self.model.generator.forward(dec_out)` followed by
 * Beam.advance's topk over beam x vocabulary, Beam.py:64-104): the GEMM epilogue keeps, per row and 128-column tile,
 * {max, sum exp} and the tile's K best logits; the [M,V] log-prob matrix is never written.  Requires the tensor-core
 * GEMM (vmmt_generator_topk_supported); otherwise use vmmt_generator_logprobs + vmmt_beam_advance. */
size_t vmmt_generator_topk_workspace_bytes(int M, int V, int K);
int vmmt_generator_topk_supported(const float* x, const float* W, int M, int H, int V, int flags);
int vmmt_generator_topk(const float* x, const float* W, const float* b, int M, int H, int V, int K, void* workspace,
                        size_t workspace_bytes, int flags, void* stream);

/* up to 5 device-to-device copies in one launch (a captured step's static input buffers; TrainerMultimodal.py:632-677 builds
 * the batch tensors per step): SM threads instead of a copy engine. */
int vmmt_copy_list(const void* const* src, void* const* dst, const int64_t* bytes, int n, void* stream);

/* ---- optimiser: global-norm clip + Adam on flat buffers (Optim.py:69-70,94-96) */
/* model.zero_grad() (TrainerMultimodal.py:627-628) on the flat gradient buffer with at most max_blocks resident blocks
 * (0 = one per SM), so that it can run beside latency-critical kernels without starving them of SM slots or HBM. */
int vmmt_fill_zero(float* p, int64_t n, int max_blocks, void* stream);
size_t vmmt_sqnorm_workspace_bytes(void);
int vmmt_sqnorm(const float* g, int64_t n, float* out1, int accumulate, void* workspace, void* stream);
int vmmt_adam_clip_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* sqnorm /*device scalar*/, float max_norm, float grad_scale, float lr,
                        float beta1, float beta2, float eps, int64_t step, void* stream);

/* ---- data-parallel optimiser step over NVLink peer memory (csrc/peer.cu).  The reference has no multi-GPU path
 * (train_mm_vi_model1.py:73-75); an N-rank step must equal its `-accum_count N` step (TrainerMultimodal.py:342-346,
 * 625-718: gradients ADD, then one clip_grad_norm + Adam, Optim.py:69-70,94-96).  Every rank owns one segment
 *   [signal block (vmmt_peer_signal_bytes) | ... flat params at param_off ... | ... flat grads at grad_off ...]
 * allocated with vmmt_peer_alloc (cudaMalloc + cudaIpcGetMemHandle; the library owns this memory because it must be
 * IPC-exportable) and maps the others with vmmt_peer_open.  `segments[j]` = base of rank j's segment as mapped here. */
size_t vmmt_peer_signal_bytes(void);
int vmmt_peer_handle_bytes(void);
int vmmt_peer_alloc(size_t bytes, void** ptr, void* handle_out /* vmmt_peer_handle_bytes() bytes */);
int vmmt_peer_open(const void* handle, void** ptr);
int vmmt_peer_close(void* ptr);
int vmmt_peer_free(void* ptr);
/* cross-GPU barrier kernel on `stream` (flag words in the signal blocks, device-resident generation counter).  Two
 * independent barrier CHANNELS (0, 1): barriers issued from two streams that may run concurrently (the all-gather of the
 * buffer's tail overlapped with the next step) must use different channels. */
int vmmt_peer_barrier(void* const* segments, int rank, int world, int channel, void* stream);
size_t vmmt_peer_adam_workspace_bytes(void);
/* [*lo, *hi) = the floats of an n-float flat buffer rank `rank` reduces and updates; returns the slice capacity
 * (floats) every rank allocates for gsum / exp_avg / exp_avg_sq. */
int64_t vmmt_peer_slice(int64_t n, int world, int rank, int64_t* lo, int64_t* hi);
/* The two halves of the step over a sub-range [begin, begin + n) of the flat buffers, so that a step can exchange the
 * gradients that are final early (generator, latent / image networks) BESIDE the rest of the backward pass (phase 1)
 * and the remainder at the end (phase 0); the clip norm is the sum over `nphase` published slot arrays. */
/* mc_base: the segments' MULTICAST address (all N segments bound to one cuMulticast object, e.g. by torch's symmetric
 * memory) or NULL.  With it the reduce-scatter is multimem.ld_reduce (the NVSwitch adds the N ranks' values in flight)
 * and the all-gather is multimem.st (one store replicated by the switch): 4 B/param/N per rank each way instead of
 * (N-1)/N * 4 B/param of P2P loads / stores. */
int vmmt_peer_reduce_scatter(void* const* segments, void* mc_base, size_t grad_off, int rank, int world, int64_t begin,
                             int64_t n, float* gsum, int phase, void* workspace, void* stream);
int vmmt_peer_adam_allgather(void* const* segments, void* mc_base, size_t param_off, int rank, int world, int64_t begin,
                             int64_t n, const float* gsum, float* exp_avg, float* exp_avg_sq, float* sqnorm_out, int nphase,
                             float max_norm, float lr, float beta1, float beta2, float eps, int64_t step,
                             int barrier_before, int barrier_after, int channel, void* stream);
/* barrier -> reduce-scatter (P2P loads, rank-ordered sum) + ||g||^2 share -> barrier -> clip + Adam on the slice,
 * new parameters stored into all N parameter buffers (P2P stores) -> barrier.  sqnorm_out (optional) receives the
 * squared global norm of the summed gradient.  No NCCL, no host synchronisation, CUDA-graph capturable. */
int vmmt_peer_adam_step(void* const* segments, void* mc_base, size_t param_off, size_t grad_off, int rank, int world,
                        int64_t n,
                        float* gsum, float* exp_avg, float* exp_avg_sq, float* sqnorm_out, float max_norm, float lr,
                        float beta1, float beta2, float eps, int64_t step, void* workspace, void* stream);

/* ---- beam search bookkeeping (Beam.py:64-123; Models.py:589-594) */
/* step index = *step_dev when step_dev != NULL (device-resident: the decode step is replayed from a CUDA graph), else
 * `step`; tok_cur / prev_cur ([K,B], optional) receive the newest tokens / back pointers at fixed addresses. */
int vmmt_beam_advance(const float* logp, int B, int K, int V, int step, const int64_t* step_dev, int64_t* tok_cur,
                      int32_t* prev_cur, int64_t eos, float* scores,
                      int64_t* next_ys, int32_t* prev_ks, float* fin_score, int32_t* fin_t, int32_t* fin_k,
                      int32_t* n_fin, int32_t* done, int32_t* n_active, void* stream);
/* same step on the workspace vmmt_generator_topk(M = K*B rows, K candidates per tile) left behind. */
int vmmt_beam_advance_topk(const void* gen_workspace, int B, int K, int V, int step, const int64_t* step_dev,
                           int64_t* tok_cur, int32_t* prev_cur, int64_t eos, float* scores, int64_t* next_ys,
                           int32_t* prev_ks, float* fin_score, int32_t* fin_t, int32_t* fin_k, int32_t* n_fin,
                           int32_t* done, int32_t* n_active, void* stream);
/* hist[*step_dev][0:n] = cur[0:n]: the step's attention rows into the history the hypotheses are read from. */
int vmmt_beam_record(const float* cur, float* hist, const int64_t* step_dev, int64_t n, void* stream);
int vmmt_beam_reorder(const float* src, float* dst, const int32_t* prev_k_step, const int32_t* done, int L,
                      int K, int B, int H, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMMT_H_ */
