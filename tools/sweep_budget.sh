#!/bin/bash
# step time vs the cluster budgets of the concurrent source / target encoder recurrences (run through gpurun)
for pair in "3 8" "4 8" "5 8" "6 4" "7 4" "5 6" "6 6" "4 6" "0 0"; do
  set -- $pair
  r=$(VMMT_ENC_BUDGET=$1 VMMT_TGT_BUDGET=$2 python bench.py --no-cpu-baseline --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4f ms  %.0f tok/s' % (d['ms_per_step'], d['value']))")
  echo "ENC=$1 TGT=$2: $r"
done
