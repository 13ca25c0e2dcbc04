set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu8.log
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_o.json 2>gpurun_out/bench_o.err; tail -1 gpurun_out/bench_o.json | cut -c1-300; grep -o '"roofline": {[^}]*}' gpurun_out/bench_o.json; tail -3 gpurun_out/bench_o.err
timeout 300 python bench.py --workload decode --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/bench_decode_o.json 2>gpurun_out/bench_decode_o.err; tail -1 gpurun_out/bench_decode_o.json | cut -c1-300; grep -o '"roofline": {[^}]*}' gpurun_out/bench_decode_o.json; tail -3 gpurun_out/bench_decode_o.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_o.json 2>gpurun_out/bench_cfg5_o.err; tail -1 gpurun_out/bench_cfg5_o.json | cut -c1-300; grep -o '"roofline": {[^}]*}' gpurun_out/bench_cfg5_o.json; tail -3 gpurun_out/bench_cfg5_o.err
