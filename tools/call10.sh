set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu21.log
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_z.json 2>gpurun_out/bench_z.err; tail -1 gpurun_out/bench_z.json | cut -c1-300; tail -3 gpurun_out/bench_z.err
python tools/timeline.py gpurun_out/timeline_z.json > gpurun_out/timeline_z.txt 2>gpurun_out/timeline_z.err; head -8 gpurun_out/timeline_z.txt
