set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu23.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu23.log
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_ab.json 2>gpurun_out/bench_ab.err; tail -1 gpurun_out/bench_ab.json | cut -c1-300; tail -3 gpurun_out/bench_ab.err
python tools/timeline.py gpurun_out/timeline_ab.json > gpurun_out/timeline_ab.txt 2>gpurun_out/timeline_ab.err; head -8 gpurun_out/timeline_ab.txt
