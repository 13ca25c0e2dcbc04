set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -x -q > gpurun_out/pytest_peer15.log 2>&1; echo "pytest peer rc=$?"; tail -4 gpurun_out/pytest_peer15.log
for n in 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_${n}gpu_u.json 2> gpurun_out/bench_${n}gpu_u.err; echo "bench$n rc=$?"; tail -1 gpurun_out/bench_${n}gpu_u.json | cut -c1-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29732 bench.py --impl reference --gpus $n --steps 2 --warmup 1 > gpurun_out/bench_ref_${n}gpu_u.json 2> gpurun_out/bench_ref_${n}gpu_u.err; echo "ref$n rc=$?"; tail -1 gpurun_out/bench_ref_${n}gpu_u.json | cut -c1-200
done
