set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_peer.py -x -q -k "8" > gpurun_out/pytest_peer8.log 2>&1; echo "pytest peer8 rc=$?"; tail -5 gpurun_out/pytest_peer8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"; tail -1 gpurun_out/bench_8gpu.json | cut -c1-1500; tail -5 gpurun_out/bench_8gpu.err
