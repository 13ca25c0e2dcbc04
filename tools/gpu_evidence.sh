#!/bin/bash
# Round-2 evidence run on one B200 (through gpurun): sanitizer summaries, ncu launch list of the bench command,
# ncu --set full capture of the dominant kernel.   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh TAG'
TAG=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitizer_${tool}_$TAG.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" gpurun_out/sanitizer_${tool}_$TAG.txt | tail -8
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_fwd -s 3 -c 1 -o gpurun_out/lstm_fwd_full_$TAG python tools/lstm_bench.py 30 40 500 500 1 > gpurun_out/ncu_lstm_$TAG.log 2>&1; echo "ncu lstm fwd rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_bwd -s 3 -c 1 -o gpurun_out/lstm_bwd_full_$TAG python tools/lstm_bench.py 30 40 500 500 1 >> gpurun_out/ncu_lstm_$TAG.log 2>&1; echo "ncu lstm bwd rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:rowlin -s 6 -c 2 -o gpurun_out/rowlin_full_$TAG python tools/rowlin_bench.py > /dev/null 2>&1; echo "ncu rowlin rc=$?"
