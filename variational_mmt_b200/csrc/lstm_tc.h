// Internal: cluster / tcgen05 LSTM recurrence (lstm_tc.cu), dispatched from vmmt_lstm_seq_{fwd,bwd}.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vmmt.h"

bool vmmt_lstm_tc_supported(int ndir, int N, int H);
int vmmt_lstm_tc_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, cudaStream_t s);
int vmmt_lstm_tc_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, cudaStream_t s);

// step-wise path for large batches / hidden sizes (lstm_step.cu): one GEMM + one fused cell kernel per step
size_t vmmt_lstm_step_workspace_floats(int ndir, int N, int H);
int vmmt_lstm_step_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, float* ws, cudaStream_t s);
int vmmt_lstm_step_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, float* ws, cudaStream_t s);
