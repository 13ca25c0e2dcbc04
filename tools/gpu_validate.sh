#!/bin/bash
# Round-end style validation on one B200 (run through gpurun): GPU parity tests, smoke, the default bench line, the CPU
# reference arm, the ncu launch list of the bench command, ncu --set full of the roofline kernel, decode and cfg5 benches.
#   gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh TAG'
TAG=${1:-x}
set -x
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cut -c1-300 gpurun_out/bench_ref_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 2 -o gpurun_out/gen_cfg1_full_$TAG python tools/gen_one.py 1240 500 10000 > gpurun_out/ncu_gen1_$TAG.log 2>&1; echo "ncu gen1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 2 -o gpurun_out/gen_cfg5_full_$TAG python tools/gen_one.py 40448 1024 32000 > gpurun_out/ncu_gen5_$TAG.log 2>&1; echo "ncu gen5 rc=$?"
timeout 300 python bench.py --workload decode --steps 4 --warmup 4 > gpurun_out/bench_decode_$TAG.json 2>gpurun_out/bench_decode_$TAG.err; cut -c1-300 gpurun_out/bench_decode_$TAG.json
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_$TAG.json 2>gpurun_out/bench_cfg5_$TAG.err; cut -c1-300 gpurun_out/bench_cfg5_$TAG.json
python tools/timeline.py gpurun_out/timeline_$TAG.json > gpurun_out/timeline_$TAG.txt 2>/dev/null; head -9 gpurun_out/timeline_$TAG.txt
