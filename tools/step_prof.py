"""Per-entry-point (and per-GEMM-shape) device time of one eager training step (cfg1 by default)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import variational_mmt_b200 as vm
from variational_mmt_b200 import synthetic, _lib
opt = synthetic.make_opt(conditional=True, dropout=0.5)
fields = synthetic.make_fields(10000, 10000)
torch.manual_seed(0)
model = vm.make_vi_model_mmt(opt, fields, gpu=True); model.train()
loss = vm.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
optim = vm.Optim("adam", 0.002, 5); optim.set_parameters(model.parameters())
src, sl, tgt, tl, img = [t.cuda() for t in synthetic.random_batch(10000, 10000, 40, 2048, seed=1, full_length=(30, 30))]
class B: pass
def step():
    model.zero_grad()
    out, attns, _ = model(src.unsqueeze(2), tgt.unsqueeze(2), sl, tl, img)
    b = B(); b.tgt = tgt; b.batch_size = 40
    loss.sharded_compute_loss(b, out, attns, 0, tgt.size(0), 32, 40)
    optim.step()
for _ in range(3): step()
prof = []
_lib.set_profile(prof); step(); torch.cuda.synchronize(); _lib.set_profile(None)
agg = {}
for name, args, e0, e1 in prof:
    key = name
    if name == "vmmt_gemm":
        key = "gemm M=%d N=%d K=%d a_k=%d b_k=%d acc=%d act=%d" % (args[8], args[9], args[10], args[2], args[5], args[13], args[12])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print("total %.3f ms in %d calls" % (tot, len(prof)))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%8.1f us total  %7.1f us/call x%-3d %s" % (1e3 * t, 1e3 * t / c, c, k))
