set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q > gpurun_out/pytest_decode6.log 2>&1; echo "pytest decode rc=$?"; tail -15 gpurun_out/pytest_decode6.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu6.log
timeout 300 python bench.py --workload decode --steps 4 --warmup 3 > gpurun_out/bench_decode_m.json 2>gpurun_out/bench_decode_m.err; tail -1 gpurun_out/bench_decode_m.json | cut -c1-1800; tail -3 gpurun_out/bench_decode_m.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_decode_m.csv python bench.py --workload decode --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_decode_m.log 2>&1; echo "ncu decode rc=$?"
