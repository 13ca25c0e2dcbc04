"""Dump the captured training-step graph as DOT (cudaGraphDebugDotPrint) and list the first nodes' dependencies."""
import os, sys, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import variational_mmt_b200 as vm
from variational_mmt_b200 import synthetic
opt = synthetic.make_opt(conditional=True, dropout=0.5)
fields = synthetic.make_fields(10000, 10000)
torch.manual_seed(0)
model = vm.make_vi_model_mmt(opt, fields, gpu=True); model.train()
loss = vm.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
optim = vm.Optim("adam", 0.002, 5); optim.set_parameters(model.parameters())
batch = [t.cuda() for t in synthetic.random_batch(10000, 10000, 40, 2048, seed=1, full_length=(30, 30))]
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_graph.dot"
os.environ["VMMT_GRAPH_DOT"] = out
g = vm.GraphedTrainStep(model, loss, shard_size=32, optim=optim)
g(*batch, 40); optim.step(); torch.cuda.synchronize()
txt = open(out).read()
nodes = dict(re.findall(r'"?(graph_\w+_node_\d+|\d+)"?\s*\[.*?label="\{?([^"]*)"', txt))
print("nodes", len(nodes), "bytes", len(txt))
