// Internal: cluster / tcgen05 LSTM recurrence (lstm_tc.cu), dispatched from vmmt_lstm_seq_{fwd,bwd}.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vmmt.h"

bool vmmt_lstm_tc_supported(int ndir, int N, int H);
// cluster_budget: cap on the clusters of this launch (0 = as many as are co-resident); two recurrences issued on two
// streams (source / target encoder) share the GPU with it
int vmmt_lstm_tc_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int cluster_budget,
                     cudaStream_t s);
int vmmt_lstm_tc_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int cluster_budget,
                     cudaStream_t s);

// step-wise path for large batches / hidden sizes (lstm_step.cu): one GEMM + one fused cell kernel per step
size_t vmmt_lstm_step_workspace_floats(int ndir, int N, int H);
int vmmt_lstm_step_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int flags, float* ws,
                       cudaStream_t s);
int vmmt_lstm_step_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, int flags, float* ws,
                       cudaStream_t s);
