#!/bin/bash
# End-of-round evidence on one B200 (through gpurun): sanitizer summaries, ncu launch list of the bench command, ncu metrics of
# the attention kernels, timeline + phase stamps, decode / cfg5 / bf16 bench lines.   gpurun --timeout 1500 -- 'bash tools/gpu_evidence2.sh TAG'
TAG=${1:-r2b}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitizer_${tool}_$TAG.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_${tool}_$TAG.txt | tail -3
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_ -s 9 -c 3 -o gpurun_out/attn_full_$TAG python tools/attn_one.py > /dev/null 2>&1; echo "ncu attn rc=$?"
python tools/timeline.py gpurun_out/timeline_$TAG.json > gpurun_out/timeline_$TAG.txt 2>/dev/null; head -14 gpurun_out/timeline_$TAG.txt | tail -12
python tools/phase_stamps.py > gpurun_out/phase_stamps_$TAG.txt 2>&1; tail -12 gpurun_out/phase_stamps_$TAG.txt
timeout 300 python bench.py --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16_$TAG.json 2> gpurun_out/bench_bf16_$TAG.err; cut -c1-200 gpurun_out/bench_bf16_$TAG.json
timeout 300 python bench.py --workload decode --steps 4 --warmup 4 > gpurun_out/bench_decode_$TAG.json 2>gpurun_out/bench_decode_$TAG.err; cut -c1-200 gpurun_out/bench_decode_$TAG.json
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_$TAG.json 2>gpurun_out/bench_cfg5_$TAG.err; cut -c1-200 gpurun_out/bench_cfg5_$TAG.json
timeout 300 python bench.py --workload cfg5 --dtype bf16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_bf16_$TAG.json 2>gpurun_out/bench_cfg5_bf16_$TAG.err; cut -c1-200 gpurun_out/bench_cfg5_bf16_$TAG.json
