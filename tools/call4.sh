set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu4.log
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_l.json 2>gpurun_out/bench_l.err; tail -1 gpurun_out/bench_l.json | cut -c1-300; grep -o '"roofline": {[^}]*}' gpurun_out/bench_l.json
timeout 300 python bench.py --workload decode --steps 4 --warmup 3 > gpurun_out/bench_decode_l.json 2>gpurun_out/bench_decode_l.err; tail -1 gpurun_out/bench_decode_l.json | cut -c1-300
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_decode_l.csv python bench.py --workload decode --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_decode.log 2>&1; echo "ncu decode rc=$?"
