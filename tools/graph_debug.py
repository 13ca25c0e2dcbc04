"""Find the first libvmmt call that invalidates a stream capture (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import variational_mmt_b200 as vm
from variational_mmt_b200 import _lib
from conftest import load_golden
from gpu_helpers import build_cuda_model, to_device
from oracle import synth

name = sys.argv[1] if len(sys.argv) > 1 else "tiny_cond_train"
meta, arr = load_golden(name)
cfg = synth.ModelConfig(**meta["cfg"])
params = synth.make_params(cfg, meta["param_seed"], meta["param_scale"])
batch = synth.make_batch(cfg, **meta["batch"])
model, fields = build_cuda_model(cfg, params)
model.train()
b = to_device(batch)
loss = vm.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
orig = _lib.call
state = {"bad": None, "n": 0}
def call(name, *args):
    orig(name, *args)
    state["n"] += 1
    if state["bad"] is None and state.get("on"):
        try:
            ok = torch.cuda.is_current_stream_capturing()
        except Exception as e:
            ok = False
        if not ok:
            state["bad"] = (state["n"], name, [a for a in args if isinstance(a, int) and abs(a) < 100000])
            print("capture invalidated after call", state["bad"], flush=True)
_lib.call = call
import variational_mmt_b200.ops as ops
ops.L.call = call
if os.environ.get("BASE_FIRST"):
    ops.rng_base(torch.device("cuda:0"))
import contextlib
if os.environ.get("EAGER_FIRST"):
    sd = torch.cuda.Stream() if os.environ.get("EAGER_SIDE") else None
    if sd is not None: sd.wait_stream(torch.cuda.current_stream())
    with vm.Normal.inject_noise(b.eps), (torch.cuda.stream(sd) if sd is not None else contextlib.nullcontext()):
        model.zero_grad()
        if os.environ.get("EAGER_NOGRAD"):
            with torch.no_grad():
                out, attns, _ = model(b.src, b.tgt_in, b.src_lengths, b.tgt_lengths, b.img_feats)
                st = loss.monolithic_compute_loss(b, out, attns)
        else:
            out, attns, _ = model(b.src, b.tgt_in, b.src_lengths, b.tgt_lengths, b.img_feats)
            if os.environ.get("EAGER_NOBWD"):
                st = loss.monolithic_compute_loss(b, out, attns)
            else:
                st = loss.sharded_compute_loss(b, out, attns, 0, b.tgt.size(0), 32, b.batch_size)
    if sd is not None: torch.cuda.current_stream().wait_stream(sd)
    print("eager done", st.nmt_loss, flush=True)
    if os.environ.get("DEL_EAGER"):
        del out, attns, st
    if os.environ.get("SYNC_BEFORE"):
        torch.cuda.synchronize()
step = vm.GraphedTrainStep(model, loss, shard_size=32)
real_capture = torch.cuda.graph.__enter__
def enter(self):
    r = real_capture(self); state["on"] = True; state["n"] = 0; return r
torch.cuda.graph.__enter__ = enter
try:
    with vm.Normal.inject_noise(b.eps):
        v = step(b.src, b.src_lengths, b.tgt_ids, b.tgt_lengths, b.img_feats, b.batch_size)
    print("ok", v.cpu())
except Exception as e:
    print("FAILED:", type(e).__name__, str(e)[:200]); print("calls during capture:", state["n"], "first bad:", state["bad"])
