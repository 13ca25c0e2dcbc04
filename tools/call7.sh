set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu7.log
timeout 300 python bench.py --workload decode --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_decode_n.json 2>gpurun_out/bench_decode_n.err; tail -1 gpurun_out/bench_decode_n.json | cut -c1-300; grep -o '"roofline": {[^}]*}' gpurun_out/bench_decode_n.json; tail -3 gpurun_out/bench_decode_n.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_decode_n.csv python bench.py --workload decode --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_decode_n.log 2>&1; echo "ncu decode rc=$?"
