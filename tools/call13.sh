set -x
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu28.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu13.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke28.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke13.log
timeout 400 python bench.py > gpurun_out/bench_af.json 2> gpurun_out/bench_af.err; echo "bench rc=$?"; cat gpurun_out/bench_af.json | cut -c1-600
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_af.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_af.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 2 -o gpurun_out/gen_cfg1_full_af python tools/gen_one.py 1240 500 10000 > gpurun_out/ncu_gen1af.log 2>&1; echo "ncu gen1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 2 -o gpurun_out/gen_cfg5_full_af python tools/gen_one.py 40448 1024 32000 > gpurun_out/ncu_gen5af.log 2>&1; echo "ncu gen5 rc=$?"
timeout 300 python bench.py --workload decode --steps 4 --warmup 4 > gpurun_out/bench_decode_af.json 2>gpurun_out/bench_decode_af.err; tail -1 gpurun_out/bench_decode_af.json | cut -c1-300
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_af.json 2>gpurun_out/bench_cfg5_af.err; tail -1 gpurun_out/bench_cfg5_af.json | cut -c1-300
