"""Batch invariance of the decode path: encoder context / final state / prior mean of a sentence inside a 250-sentence
batch vs the same sentence alone (bitwise), stage by stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import variational_mmt_b200 as vm
from variational_mmt_b200 import synthetic, ops

opt = synthetic.make_opt(conditional=True, dropout=0.5)
fields = synthetic.make_fields(10000, 10000)
torch.manual_seed(3435)
model = vm.make_vi_model_mmt(opt, fields, gpu=True)
model.eval()
src, sl, _t, _tl, _img = synthetic.random_batch(10000, 10000, 250, 8, seed=77)
src, sl = src.cuda(), sl.cuda()
with torch.no_grad(), ops.batch_invariant():
    emb_b = model.encoder.embeddings(src.unsqueeze(2))
    (hb, cb), ctx_b = model.encoder(src.unsqueeze(2), sl)
    q0, _ = model.gen_net_global(ctx_b, sl)
    zb = q0.mean()
    for i in [0, 1, 57, 249]:
        n = int(sl[i])
        s1 = src[:n, i:i + 1].contiguous()
        emb_1 = model.encoder.embeddings(s1.unsqueeze(2))
        (h1, c1), ctx_1 = model.encoder(s1.unsqueeze(2), sl[i:i + 1].contiguous())
        q1, _ = model.gen_net_global(ctx_1, sl[i:i + 1].contiguous())
        z1 = q1.mean()
        # layer-0 input projection alone
        w = model.encoder.rnn.weight_ih_l0
        g_b = torch.empty(src.shape[0] * 250, w.shape[0], device="cuda")
        ops.gemm(emb_b.view(-1, 500), w, g_b, src.shape[0] * 250, w.shape[0], 500)
        g_1 = torch.empty(n, w.shape[0], device="cuda")
        ops.gemm(emb_1.view(-1, 500), w, g_1, n, w.shape[0], 500)
        d = lambda a, b: float((a - b).abs().max())
        print(f"sentence {i} len {n}: emb {d(emb_b[:n, i], emb_1[:, 0]):.2e} gx {d(g_b.view(-1, 250, 2000)[:n, i], g_1):.2e} "
              f"ctx {d(ctx_b[:n, i], ctx_1[:, 0]):.2e} h {d(hb[:, i], h1[:, 0]):.2e} c {d(cb[:, i], c1[:, 0]):.2e} z {d(zb[i], z1[0]):.2e}")
