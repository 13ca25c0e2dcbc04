"""Kernel timeline of one graph-replayed training step (torch.profiler / CUPTI; no nsys in this image):
per-stream busy time, idle gaps on the union of all streams, and the longest kernels in start order."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import variational_mmt_b200 as vm
from variational_mmt_b200 import synthetic
opt = synthetic.make_opt(conditional=True, dropout=0.5)
fields = synthetic.make_fields(10000, 10000)
torch.manual_seed(0)
model = vm.make_vi_model_mmt(opt, fields, gpu=True); model.train()
loss = vm.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
optim = vm.Optim("adam", 0.002, 5); optim.set_parameters(model.parameters())
if world > 1:
    optim.enable_early_exchange(model)
batch = [t.cuda() for t in synthetic.random_batch(10000, 10000, 40, 2048, seed=1, full_length=(30, 30))]
g = vm.GraphedTrainStep(model, loss, shard_size=32, optim=optim)
def step():
    g(*batch, 40 * world); optim.step()
for _ in range(5): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json"
if rank != 0:
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
prof.export_chrome_trace(out + ".trace.json")
tr = json.load(open(out + ".trace.json"))
ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
# split into steps at the (last) Adam kernel of each step
ends = [i for i, e in enumerate(ev) if "adam_clip" in e["name"] or "peer_adam_allgather" in e["name"]]
if world > 1:
    ends = [i for j, i in enumerate(ends) if j + 1 == len(ends) or ends[j + 1] - i > 20]      # the step's last update kernel
if len(ends) >= 2:
    ev = ev[ends[-2] + 1: ends[-1] + 1]
t0 = ev[0]["ts"]
rows = [(e["ts"] - t0, e["dur"], e["args"].get("stream", -1), e["name"][:70]) for e in ev]
json.dump(rows, open(out, "w"))
span = max(r[0] + r[1] for r in rows)
print("kernels in step: %d, span %.1f us, sum of durations %.1f us" % (len(rows), span, sum(r[1] for r in rows)))
streams = {}
for r in rows: streams.setdefault(r[2], []).append(r)
for s, rs in streams.items():
    print("stream %s: %d kernels, busy %.1f us" % (s, len(rs), sum(r[1] for r in rs)))
# union busy / idle
iv = sorted((r[0], r[0] + r[1]) for r in rows)
busy, cur_s, cur_e, gaps = 0.0, iv[0][0], iv[0][1], []
for a, b in iv[1:]:
    if a > cur_e:
        busy += cur_e - cur_s; gaps.append((a - cur_e, cur_e)); cur_s, cur_e = a, b
    else:
        cur_e = max(cur_e, b)
busy += cur_e - cur_s
print("union busy %.1f us, idle %.1f us in %d gaps (>=2us: %d, total %.1f us)" % (
    busy, span - busy, len(gaps), sum(1 for g_, _ in gaps if g_ >= 2), sum(g_ for g_, _ in gaps if g_ >= 2)))
print("--- timeline (start us, dur us, stream, name)")
for r in rows:
    print("%8.1f %7.1f  s%-3s %s" % r)

if world > 1:
    dist.barrier(); dist.destroy_process_group()
