set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"; cat gpurun_out/bench_j.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_j.json 2>&1; cat gpurun_out/bench_ref_j.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_j.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 3 -o gpurun_out/gen_cfg1_full python tools/gen_one.py 1240 500 10000 > gpurun_out/ncu_gen1.log 2>&1; echo "ncu gen1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 3 -o gpurun_out/gen_cfg5_full python tools/gen_one.py 40448 1024 32000 > gpurun_out/ncu_gen5.log 2>&1; echo "ncu gen5 rc=$?"
ls -la gpurun_out
