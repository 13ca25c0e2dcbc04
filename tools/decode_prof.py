"""Per-entry-point device time of one eager decode step (cfg1 model, beam 5)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import variational_mmt_b200 as vm
from variational_mmt_b200 import synthetic, _lib
Bd = int(sys.argv[1]) if len(sys.argv) > 1 else 250
opt = synthetic.make_opt(conditional=True, dropout=0.5)
fields = synthetic.make_fields(10000, 10000)
model = vm.make_vi_model_mmt(opt, fields, gpu=True); model.eval()
tr = vm.TranslatorMultimodalVI(model, fields, beam_size=5, max_length=100, global_scorer=vm.GNMTGlobalScorer(0., -0.), cuda=True,
                               test_img_feats=np.zeros((1, 2048), np.float32), multimodal_model_type="vi-model1")
tr.use_graph = False
src, sl, *_ = synthetic.random_batch(10000, 10000, Bd, 8, seed=1)
class B: pass
b = B(); b.batch_size = Bd; b.src = (src.cuda(), sl.cuda())
tr.max_length = 12
tr.translate_batch(b)
prof = []
_lib.set_profile(prof)
tr.translate_batch(b)
torch.cuda.synchronize()
_lib.set_profile(None)
agg = {}
for name, args, e0, e1 in prof:
    key = name + ((" M=%d N=%d K=%d" % (args[8], args[9], args[10])) if name == "vmmt_gemm" else "")
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print("S=%d R=%d total %.3f ms over 12 steps (+encoder)" % (src.size(0), 5 * Bd, tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("%8.1f us/call x%-3d %s" % (1e3 * t / c, c, k))
