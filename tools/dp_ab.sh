#!/bin/bash
# A/B of the data-parallel exchange variants at N GPUs: tools/dp_ab.sh N   (run through gpurun --gpus N)
N=${1:-8}
for nv in 1 0; do for ov in 1 0; do
  r=$(VMMT_DP_NVLS=$nv VMMT_DP_AG_OVERLAP=$ov python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2977$nv bench.py --gpus $N --steps 30 --warmup 5 --no-dp-parity 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4))")
  echo "N=$N NVLS=$nv OVERLAP=$ov: $r"
done; done
