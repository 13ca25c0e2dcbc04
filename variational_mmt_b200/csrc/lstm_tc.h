// Internal: cluster / tcgen05 LSTM recurrence (lstm_tc.cu), dispatched from vmmt_lstm_seq_{fwd,bwd}.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vmmt.h"

bool vmmt_lstm_tc_supported(int ndir, int N, int H);
int vmmt_lstm_tc_fwd(const VmmtLstmDir* dirs, int ndir, const int64_t* lengths, int T, int N, int H, cudaStream_t s);
int vmmt_lstm_tc_bwd(const VmmtLstmDirBwd* dirs, int ndir, const int64_t* lengths, int T, int N, int H, cudaStream_t s);
