"""Generate tests/golden/*.npz by EXECUTING the unmodified reference in the build container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run:  python -m oracle.make_golden
The reference has no tests or golden vectors of its own, so these fixtures --
outputs of the reference's own code on deterministic synthetic inputs (oracle/synth.py) -- are what
pins the CPU restatement (tests/test_oracle_golden.py) and, through it, the CUDA path.
Fixtures store inputs' *recipe* (config, seeds) and the reference's outputs; large tensors are
stored as strided samples + norms.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BIG = 4096           # tensors above this many elements are stored as samples


def sample_of(a, n=257):
    a = np.asarray(a).reshape(-1)
    if a.size <= BIG:
        return a.copy()
    idx = (np.arange(n, dtype=np.int64) * 7919) % a.size
    return a[idx].copy()


class RefBatch:
    pass


def ref_batch(ns, b):
    t = ns.torch
    rb = RefBatch()
    rb.tgt = t.as_tensor(b.tgt)
    rb.batch_size = b.batch_size
    return rb


def run_forward(ns, model, b, training, eps=None):
    t = ns.torch
    src = t.as_tensor(b.src).unsqueeze(2)
    tgt = t.as_tensor(b.tgt).unsqueeze(2)
    args = (src, tgt, t.as_tensor(b.src_lengths), t.as_tensor(b.tgt_lengths),
            t.as_tensor(b.img_feats).clone())
    if training:
        model.train()
        with ns.inject_noise(t.as_tensor(eps)):
            return model(*args)
    model.eval()
    with t.no_grad():
        return model(*args)


def stats_dict(st):
    f = lambda x: float(x.reshape(-1)[0]) if hasattr(x, "reshape") else float(x)
    return dict(nmt=f(st.nmt_loss), td_kl_before=f(st.td_kl_before), td_kl_after=f(st.td_kl_after),
                img_feats_loss=f(st.image_feats_loss), img_feats_cos=f(st.image_feats_cos),
                elbo=f(st.elbo_loss), n_words=int(st.n_words), n_correct=int(st.n_correct))


def case_eval(ns, name, cfg, batch_kw, pseed=3435, scale=0.1):
    params = synth.make_params(cfg, pseed, scale)
    b = synth.make_batch(cfg, **batch_kw)
    model, fields = ns.build_model(cfg, params)
    out, attns, _ = run_forward(ns, model, b, training=False)
    # copies before the loss normalises the image tensors in place (H3)
    rec = dict(out=out.numpy().copy(), attn=attns["std"].numpy().copy(),
               mu_q=attns["z_latent"][0].params()[0].numpy().copy(),
               sd_q=attns["z_latent"][0].params()[1].numpy().copy(),
               mu_p=attns["p_latent"][0].params()[0].numpy().copy(),
               sd_p=attns["p_latent"][0].params()[1].numpy().copy(),
               z=attns["z0_sample"][0].numpy().copy(),
               img_loc=attns["p_global_image_features"][0].params()[0].numpy().copy())
    loss = ns.make_loss(model, fields)
    with ns.torch.no_grad():
        st = loss.monolithic_compute_loss(ref_batch(ns, b), out, attns)
    save(name, cfg, batch_kw, pseed, scale, rec, stats_dict(st), mode="eval")


def case_train(ns, name, cfg, batch_kw, pseed=3435, scale=0.1, lr=0.002):
    t = ns.torch
    params = synth.make_params(cfg, pseed, scale)
    b = synth.make_batch(cfg, **batch_kw)
    model, fields = ns.build_model(cfg, params)
    out, attns, _ = run_forward(ns, model, b, training=True, eps=b.eps)
    rec = dict(out=out.detach().numpy().copy(), attn=attns["std"].detach().numpy().copy(),
               mu_q=attns["z_latent"][0].params()[0].detach().numpy().copy(),
               sd_q=attns["z_latent"][0].params()[1].detach().numpy().copy(),
               z=attns["z0_sample"][0].detach().numpy().copy())
    loss = ns.make_loss(model, fields)
    model.zero_grad()
    T = b.tgt.shape[0]
    st = loss.sharded_compute_loss(ref_batch(ns, b), out, attns, 0, T, 32, b.batch_size)
    gnorm, gsample, nograd = {}, {}, []
    for k, prm in model.named_parameters():
        if k.startswith("encoder_tgt.embeddings"):
            continue
        if prm.grad is None:
            nograd.append(k)
            continue
        g = prm.grad.detach().numpy()
        gnorm[k] = float(np.sqrt((g.astype(np.float64) ** 2).sum()))
        gsample[k] = sample_of(g)
    for k in gnorm:
        rec["grad/" + k] = gsample[k]
    # one optimiser update (Optim.py:78-96): clip 5 + Adam(eps 1e-9)
    optim = ns.onmt.Optim("adam", lr, 5, lr_decay=0.5, start_decay_at=8)
    optim.set_parameters(model.parameters())
    optim.step()
    for k, prm in model.named_parameters():
        if k.startswith("encoder_tgt.embeddings") or k in nograd:
            continue
        rec["delta/" + k] = sample_of(prm.detach().numpy().astype(np.float64)
                                      - params[k].astype(np.float64)).astype(np.float32)
    extra = dict(grad_norm=gnorm, no_grad=nograd,
                 total_grad_norm=float(np.sqrt(sum(v * v for v in gnorm.values()))), lr=lr)
    save(name, cfg, batch_kw, pseed, scale, rec, stats_dict(st), mode="train", extra=extra)


def case_decode(ns, name, cfg, n_sent, beam, pseed=3435, scale=0.5, max_length=30):
    t = ns.torch
    params = synth.make_params(cfg, pseed, scale)
    model, fields = ns.build_model(cfg, params)
    model.eval()
    scorer = ns.onmt.translate.GNMTGlobalScorer(0., -0.)
    tr = ns.onmt.translate.TranslatorMultimodalVI(
        model, fields, beam_size=beam, n_best=1, max_length=max_length, global_scorer=scorer,
        copy_attn=False, cuda=False, test_img_feats=np.zeros((n_sent, cfg.img_dim), np.float32),
        multimodal_model_type="vi-model1")
    b = synth.make_batch(cfg, batch_size=n_sent, seed=11)
    rec, toks = {}, []

    class Data:
        data_type = "text"
    for i in range(n_sent):
        L = int(b.src_lengths[i])
        rb = RefBatch()
        rb.batch_size = 1
        rb.src = (t.as_tensor(b.src[:L, i:i + 1]), t.as_tensor([L]))
        with t.no_grad():
            ret = tr.translate_batch(rb, Data(), i)
        hyp = [int(x) for x in ret["predictions"][0][0]]
        rec[f"tokens/{i}"] = np.asarray(hyp, np.int64)
        rec[f"score/{i}"] = np.asarray(float(ret["scores"][0][0]), np.float64)
        rec[f"attn/{i}"] = ret["attention"][0][0].numpy().copy()
        toks.append(hyp)
    save(name, cfg, dict(batch_size=n_sent, seed=11), pseed, scale, rec, {}, mode="decode",
         extra=dict(beam=beam, max_length=max_length, n_sent=n_sent))


def save(name, cfg, batch_kw, pseed, scale, rec, stats, mode, extra=None):
    meta = dict(name=name, mode=mode, cfg=cfg.to_dict(), batch=batch_kw, param_seed=pseed,
                param_scale=scale, stats=stats, extra=extra or {},
                generator="oracle/make_golden.py (executed reference, torch %s)" % _torch_version())
    arrays = {}
    for k, v in rec.items():
        v = np.asarray(v)
        arrays[k] = v if v.size <= BIG else sample_of(v)
        if v.size > BIG:
            arrays[k + "@norm"] = np.asarray(np.sqrt((v.astype(np.float64) ** 2).sum()))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), __meta__=np.asarray(json.dumps(meta)), **arrays)
    print("wrote", name, {k: round(v, 6) if isinstance(v, float) else v for k, v in stats.items()})


def _torch_version():
    import torch
    return torch.__version__


def main():
    assert ref_shims.available(), "reference not mounted; golden fixtures can only be made in the build container"
    ns = ref_shims.load()
    ns.torch.set_num_threads(8)
    only = sys.argv[1:] or None

    def want(n):
        return only is None or n in only
    ragged5 = dict(batch_size=5, seed=1)
    if want("tiny_cond_eval"):
        case_eval(ns, "tiny_cond_eval", synth.TINY, ragged5)
    if want("tiny_fixed_eval"):
        case_eval(ns, "tiny_fixed_eval", synth.TINY_FIXED, ragged5)
    # tgt up to 40 positions > shard 32: exercises the first-shard truncation (H4)
    long6 = dict(batch_size=6, seed=2, t_force=40)
    if want("tiny_cond_train"):
        case_train(ns, "tiny_cond_train", synth.TINY, long6)
    if want("tiny_fixed_train"):
        case_train(ns, "tiny_fixed_train", synth.TINY_FIXED, long6)
    if want("tiny_cond_beam5"):
        case_decode(ns, "tiny_cond_beam5", synth.TINY, 4, 5)
    if want("tiny_cond_greedy"):
        case_decode(ns, "tiny_cond_greedy", synth.TINY, 4, 1)
    if want("tiny_fixed_beam5"):
        case_decode(ns, "tiny_fixed_beam5", synth.TINY_FIXED, 3, 5)
    # bidirectional source encoder (-encoder_type brnn, opts.py:54-58; Models.py:107-109,1158-1165)
    if want("tiny_brnn_eval"):
        case_eval(ns, "tiny_brnn_eval", synth.TINY_BRNN, ragged5)
    if want("tiny_brnn_train"):
        case_train(ns, "tiny_brnn_train", synth.TINY_BRNN, dict(batch_size=6, seed=2, t_force=20))
    if want("cfg1_eval"):
        case_eval(ns, "cfg1_eval", synth.CFG1, dict(batch_size=40, seed=3))
    if want("cfg1_train"):
        case_train(ns, "cfg1_train", synth.CFG1, dict(batch_size=40, seed=4, full_length=(30, 30)))
    if want("cfg1_fixed_train"):
        case_train(ns, "cfg1_fixed_train", synth.CFG_FIXED, dict(batch_size=40, seed=5))


if __name__ == "__main__":
    main()
