"""Phase boundaries of the replayed training step WITHOUT a profiler: %globaltimer stamps (VMMT_STAMPS=1) averaged over steps."""
import os, sys
os.environ["VMMT_STAMPS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import variational_mmt_b200 as vm
from variational_mmt_b200 import synthetic, ops
opt = synthetic.make_opt(conditional=True, dropout=0.5)
fields = synthetic.make_fields(10000, 10000)
torch.manual_seed(0)
model = vm.make_vi_model_mmt(opt, fields, gpu=True); model.train()
loss = vm.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
optim = vm.Optim("adam", 0.002, 5); optim.set_parameters(model.parameters())
batches = [[t.cuda() for t in synthetic.random_batch(10000, 10000, 40, 2048, seed=s, full_length=(30, 30))] for s in range(4)]
g = vm.GraphedTrainStep(model, loss, shard_size=32, optim=optim)
names = {0: "forward start", 1: "source encoder done", 2: "target encoder done", 3: "latent block done (decoder may start)",
         4: "decoder + attention done", 5: "loss forward done", 14: "d(decoder output) ready (generator backward done)",
         11: "d(context) ready (decoder backward done)", 12: "d(target context) ready", 15: "backward main chain done",
         16: "weight-gradient lanes joined", 20: "after clip + Adam"}
acc, n = {}, 0
for i in range(30):
    g(*batches[i % 4], 40)
    optim.step()
    ops.stamp(20)
    if i >= 10:
        torch.cuda.synchronize()
        b = ops.stamp_buffer().tolist()
        for k in names:
            acc[k] = acc.get(k, 0.0) + (b[k] - b[0]) / 1e3
        n += 1
for k in sorted(names, key=lambda k: acc[k]):
    print("%9.1f us  %s" % (acc[k] / n, names[k]))
