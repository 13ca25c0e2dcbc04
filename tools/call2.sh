set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -6
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -x -q > gpurun_out/pytest_peer.log 2>&1; echo "pytest peer rc=$?"; tail -15 gpurun_out/pytest_peer.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu2.log
for mode in auto nccl; do
VMMT_DP_EXCHANGE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu_$mode.json 2> gpurun_out/bench_2gpu_$mode.err; echo "bench2 $mode rc=$?"; tail -1 gpurun_out/bench_2gpu_$mode.json | cut -c1-900; tail -3 gpurun_out/bench_2gpu_$mode.err
done
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_k.json 2>gpurun_out/bench_k.err; tail -1 gpurun_out/bench_k.json | cut -c1-400
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_k.json 2>gpurun_out/bench_cfg5_k.err; tail -1 gpurun_out/bench_cfg5_k.json | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 2 -o gpurun_out/gen_cfg5_full_k python tools/gen_one.py 40448 1024 32000 > gpurun_out/ncu_gen5k.log 2>&1; echo "ncu gen5 rc=$?"
