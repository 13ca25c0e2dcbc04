python -m pytest tests/test_gpu_kernels.py -q -k lstm 2>&1 | tail -2
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_target.py > gpurun_out/sanitizer_${tool}_r2.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY| ok" gpurun_out/sanitizer_${tool}_r2.txt | tail -8
done
python tools/lstm_bench.py 30 40 500 500 1
