mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python bench.py --steps 40 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run', d['ms_per_step'], d['value'])"; done
python tools/timeline.py gpurun_out/timeline_ae.json > gpurun_out/timeline_ae.txt 2>/dev/null; awk 'NR>9 && $2>=60' gpurun_out/timeline_ae.txt | cut -c1-90
