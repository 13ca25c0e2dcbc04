#!/bin/bash
# LSTM recurrence timing + per-step cycle trace for the three cfg1 stacks (run through gpurun).
for a in "30 40 500 500 1" "40 32 500 250 2" "31 40 1000 500 1"; do
  python tools/lstm_bench.py $a
  VMMT_LSTM_TRACE=1 python tools/lstm_bench.py $a 2>&1 | grep trace | tail -4
done
