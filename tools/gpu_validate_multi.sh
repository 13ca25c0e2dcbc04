#!/bin/bash
# Multi-GPU validation on one NVSwitch box:  gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_validate_multi.sh TAG'
TAG=${1:-x}
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -x -q > gpurun_out/pytest_peer_$TAG.log 2>&1; echo "pytest peer rc=$?"; tail -4 gpurun_out/pytest_peer_$TAG.log
for n in 8 4 2; do
  [ $n -le $NG ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2974$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_${n}gpu_$TAG.json 2> gpurun_out/bench_${n}gpu_$TAG.err; echo "bench$n rc=$?"; tail -1 gpurun_out/bench_${n}gpu_$TAG.json | cut -c1-230
done
