"""Measured parity of the BENCHMARKED mode (TF32 tcgen05 GEMMs + fp16-operand recurrence) against the CPU oracle at
config-1 size: max RELATIVE error of the attention weights over unmasked entries (north_star: 1e-3), per-stage max-abs
errors, loss / KL relative errors.  Prints one JSON object; `python tools/parity_probe.py [golden-name ...]`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from conftest import load_golden
from gpu_helpers import build_cuda_model, to_device
from oracle import synth
from oracle import vi_model1_ref as R


def probe(name, gemm_mode, dtype_mode="f32"):
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import _lib, ops
    meta, arr = load_golden(name)
    cfg = synth.ModelConfig(**meta["cfg"])
    params = synth.make_params(cfg, meta["param_seed"], meta["param_scale"])
    batch = synth.make_batch(cfg, **meta["batch"])
    train = name.endswith("train")
    ops.set_gemm_mode(gemm_mode)
    if hasattr(vm, "set_compute_dtype"):
        vm.set_compute_dtype(dtype_mode)
    try:
        model, fields = build_cuda_model(cfg, params)
        model.train(train)
        cap = {}
        model.encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("context", o[1].detach()))
        model.decoder.rnn.register_forward_hook(lambda m, i, o: cap.__setitem__("rnn_out", o[0].detach()))
        b = to_device(batch)
        loss = vm.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
        if train:
            model.zero_grad()
            with vm.Normal.inject_noise(b.eps):
                out, attns, _ = model(b.src, b.tgt_in, b.src_lengths, b.tgt_lengths, b.img_feats)
            st = loss.sharded_compute_loss(b, out, attns, 0, b.tgt.size(0), 32, b.batch_size)
            _, ostats, ofwd = R.train_step_grads(params, cfg, batch)
        else:
            with torch.no_grad():
                out, attns, _ = model(b.src, b.tgt_in, b.src_lengths, b.tgt_lengths, b.img_feats)
                st = loss.monolithic_compute_loss(b, out, attns)
            ofwd, ostats, _ = R.eval_step(params, cfg, batch)
    finally:
        ops.set_gemm_mode(0)
        if hasattr(vm, "set_compute_dtype"):
            vm.set_compute_dtype("f32")
    a = attns["std"].detach().cpu().numpy().astype(np.float64)
    ar = ofwd["attn"].detach().numpy().astype(np.float64)
    S = a.shape[2]
    live = np.arange(S)[None, None, :] < np.asarray(batch.src_lengths)[None, :, None]
    live = np.broadcast_to(live, a.shape)
    rel = np.abs(a - ar)[live] / ar[live]
    o = out.detach().cpu().numpy().astype(np.float64)
    res = dict(case=name, gemm_mode=gemm_mode, dtype=dtype_mode,
               attn_max_rel=float(rel.max()), attn_median_rel=float(np.median(rel)),
               attn_min_ref=float(ar[live].min()), attn_max_abs=float(np.abs(a - ar).max()),
               out_max_abs=float(np.abs(o - ofwd["out"].detach().numpy()).max()),
               nll_rel=abs(st.nmt_loss - ostats["nmt"]) / abs(ostats["nmt"]),
               kl_rel=abs(st.td_kl_before - ostats["td_kl_before"]) / abs(ostats["td_kl_before"]),
               img_rel=abs(st.image_feats_loss - ostats["img_feats_loss"]) / abs(ostats["img_feats_loss"]))
    for k in ("context", "rnn_out"):
        res[k + "_max_abs"] = float(np.abs(cap[k].cpu().numpy() - ofwd[k].detach().numpy()).max())
    if "--per-step" in sys.argv:
        for k in ("context", "rnn_out"):
            e = np.abs(cap[k].cpu().numpy() - ofwd[k].detach().numpy()).max(axis=(1, 2))
            res[k + "_err_by_t"] = [float("%.2e" % v) for v in e]
            res[k + "_absmax_by_t"] = [float("%.2e" % v) for v in np.abs(ofwd[k].detach().numpy()).max(axis=(1, 2))]
    res["z_max_abs"] = float(np.abs(attns["z0_sample"][0].detach().cpu().numpy() - ofwd["z"].detach().numpy()).max())
    for k in ("mu_q", "sd_q", "mu_p", "sd_p"):
        src = attns["z_latent"][0] if k.endswith("q") else attns["p_latent"][0]
        v = src.params()[0 if k.startswith("mu") else 1].detach().cpu().numpy()
        res[k + "_max_abs"] = float(np.abs(v - ofwd[k].detach().numpy()).max())
    return res


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg1_eval", "cfg1_train", "cfg1_fixed_train"]
    dts = ["f32", "bf16"] if "--bf16" in sys.argv else ["f32"]
    outp = []
    for n in names:
        for dt in dts:
            for gm in ((1, 0) if dt == "f32" else (0,)):
                r = probe(n, gm, dt)
                outp.append(r)
                print(json.dumps(r), flush=True)
