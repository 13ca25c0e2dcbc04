set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -x -q > gpurun_out/pytest_peer29.log 2>&1; echo "pytest peer rc=$?"; tail -4 gpurun_out/pytest_peer29.log
for n in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2974$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_${n}gpu_final.json 2> gpurun_out/bench_${n}gpu_final.err; echo "bench$n rc=$?"; tail -1 gpurun_out/bench_${n}gpu_final.json | cut -c1-230
done
