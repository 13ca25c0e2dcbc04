"""CPU baseline of the VI-model-1 training step: the reported `cpu_baseline` / `--impl reference` arm.

TEST INFRASTRUCTURE (see oracle/__init__.py); only bench.py and tests/ import it.

The reference is a Python package over PyTorch that lives in /root/reference and cannot travel to
the GPU box, so the CPU arm is a *port*: the same step the reference executes on a CPU
(NMTVIModel.forward -> sharded_compute_loss incl. backward -> Optim.step; onmt/Models.py:850-1011,
onmt/Loss.py:88-132, onmt/Optim.py:78-96) built from the same torch library calls the reference
makes -- ``nn.LSTM`` (oneDNN RNN on CPU) with ``pack_padded_sequence`` for the encoder
(onmt/Models.py:139-147), ``nn.LSTM`` for the decoder and the bidirectional target encoder, dense
``torch.optim.Adam(eps=1e-9)`` + ``clip_grad_norm_(5)`` (onmt/Optim.py:69-70,94-95) -- and from the
oracle restatement (oracle/vi_model1_ref.py) for everything else.  tests/test_oracle_golden.py checks
that this fast path equals the explicit-loop restatement (and therefore the executed reference).
"""
import time

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from . import synth
from . import vi_model1_ref as R


class CpuStep(nn.Module):
    """Parameters under the reference state_dict keys; LSTMs as nn.LSTM modules sharing them."""

    def __init__(self, cfg, params, dropout=0.0):
        super().__init__()
        self.cfg = cfg
        self.p = nn.ParameterDict()
        self._names = {}
        for k, v in params.items():
            kk = k.replace(".", "/")
            self._names[k] = kk
            self.p[kk] = nn.Parameter(torch.as_tensor(np.asarray(v)).float().clone())
        E, H, Z, L = cfg.emb, cfg.hidden, cfg.z_dim, cfg.layers
        self.dropout = dropout
        self.enc = nn.LSTM(E, H, L, dropout=dropout)
        self.dec = nn.LSTM(E + Z, H, L, dropout=dropout)
        self._tie(self.enc, "encoder.rnn")
        self._tie(self.dec, "decoder.rnn")
        if cfg.conditional:
            self.tgt = nn.LSTM(E, H // 2, L, dropout=dropout, bidirectional=True)
            self._tie(self.tgt, "encoder_tgt.rnn")

    def _tie(self, rnn, prefix):
        for name, _ in list(rnn.named_parameters()):
            setattr(rnn, name, self.p[self._names[f"{prefix}.{name}"]])
        rnn.flatten_parameters()

    def P(self):
        return {k: self.p[kk] for k, kk in self._names.items()}

    def forward_loss(self, b, training=True, shard_size=32, normalization=None):
        """-> (loss / normalization, stats)."""
        cfg, p = self.cfg, self.P()
        src, lengths, tgt, tl, v = b["src"], b["src_lengths"], b["tgt"], b["tgt_lengths"], b["img_feats"]
        drop = (lambda x: torch.dropout(x, self.dropout, True)) if (training and self.dropout > 0) else (lambda x: x)
        x = p["encoder.embeddings.make_embedding.emb_luts.0.weight"][src]
        packed = pack_padded_sequence(x, lengths.cpu())
        out, (h_enc, c_enc) = self.enc(packed)
        ctx = pad_packed_sequence(out)[0]
        hx = R.masked_mean(ctx, lengths)
        if cfg.conditional:
            mu_p = R.mlp2(p, "gen_net_global.location", hx)
            sd_p = R.mlp2(p, "gen_net_global.scale", hx, softplus=True)
            y = p["decoder.embeddings.make_embedding.emb_luts.0.weight"][tgt].transpose(0, 1)
            yctx = self.tgt(y)[0].transpose(0, 1)
            hq = torch.cat([R.masked_mean(ctx.detach(), lengths), R.masked_mean(yctx, tl), v], 1)
        else:
            hq = R.masked_mean(ctx.detach(), lengths)
        mu_q = R.mlp2(p, "inf_net_global.location", hq)
        sd_q = R.mlp2(p, "inf_net_global.scale", hq, softplus=True)
        if not cfg.conditional:
            mu_p, sd_p = torch.zeros_like(mu_q), torch.ones_like(mu_q)
        if training:
            eps = b["eps"] if b.get("eps") is not None else torch.randn_like(mu_q)
            z = (mu_q + sd_q * eps).detach()
        else:
            z = (mu_p if cfg.conditional else mu_q).detach()
        e = p["decoder.embeddings.make_embedding.emb_luts.0.weight"][tgt[:-1]]
        u = torch.cat([e, z.unsqueeze(0).expand(e.shape[0], -1, -1)], 2)
        q, _ = self.dec(u, (h_enc, c_enc))
        attn_h, align = R.global_attention(q, ctx, lengths, p["decoder.attn.linear_in.weight"],
                                           p["decoder.attn.linear_out.weight"])
        attn_h = drop(attn_h)
        loc_v, _ = R.image_head(p, z)
        fwd = dict(out=attn_h, attn=align, mu_q=mu_q, sd_q=sd_q, mu_p=mu_p, sd_p=sd_p, img_loc=loc_v, z=z)
        loss, stats, _ = R.compute_loss(p, cfg, fwd, b, shard_size=shard_size if training else None)
        norm = normalization if normalization is not None else src.shape[1]
        return loss / norm, stats, fwd


def _cfg_of(mk):
    return synth.ModelConfig(v_src=mk["v"], v_tgt=mk["v"], emb=mk["emb"], hidden=mk["hidden"], z_dim=mk["z"],
                             conditional=mk["conditional"])


def time_train_steps(mk, bk, steps=3, warmup=1, threads=None, budget_s=25.0, dropout=0.5, device="cpu"):
    """Times full training steps (fwd + loss + bwd + clip + Adam) of the port on synthetic batches of the
    workload; stops early once `budget_s` seconds of timed work have been spent (>= 1 step).
    ``device="cuda"``: the same torch library calls on the GPU (cuDNN nn.LSTM + cuBLAS: the library path the reference
    used on its GPUs, README.md:49) -- bench.py's informational ``gpu_torch_baseline`` yardstick."""
    if threads:
        torch.set_num_threads(int(threads))
    cfg = _cfg_of(mk)
    params = synth.make_params(cfg, 3435, 0.1)
    model = CpuStep(cfg, params, dropout=dropout)
    model.train()
    on_gpu = str(device).startswith("cuda")
    if on_gpu:
        model = model.to(device)
        for rnn in (model.enc, model.dec, getattr(model, "tgt", None)):
            if rnn is not None:
                rnn.flatten_parameters()
    opt = torch.optim.Adam(model.parameters(), lr=0.002, betas=(0.9, 0.999), eps=1e-9)
    shard = 32 if mk["hidden"] < 1024 else 128
    batches = [R.to_torch_batch(synth.make_batch(cfg, batch_size=bk["batch_size"], seed=i,
                                                 full_length=bk.get("full_length"),
                                                 src_max=80 if mk["hidden"] >= 1024 else 50,
                                                 tgt_max=80 if mk["hidden"] >= 1024 else 50))
               for i in range(2)]
    if on_gpu:
        batches = [{k: v.to(device) for k, v in b.items()} for b in batches]

    def sync():
        if on_gpu:
            torch.cuda.synchronize()

    def one(b):
        opt.zero_grad(set_to_none=True)
        loss, stats, _ = model.forward_loss(b, training=True, shard_size=shard)
        loss.backward()
        emb_keys = [k for k in model._names if "emb_luts" in k]
        for k in emb_keys:                                   # padding_idx row gets no gradient
            g = model.p[model._names[k]].grad
            if g is not None:
                g[synth.PAD] = 0
        torch.nn.utils.clip_grad_norm_([q for q in model.parameters() if q.grad is not None], 5.0)
        opt.step()
        return stats["n_words"]

    for i in range(warmup):
        one(batches[i % 2])
    sync()
    done, tok, t0 = 0, 0, time.perf_counter()
    for i in range(max(steps, 1)):
        tok += one(batches[i % 2])
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    sync()
    dt = time.perf_counter() - t0
    if on_gpu:
        return {"tokens_per_s": tok / dt, "ms_per_step": dt / done * 1e3, "steps": done, "warmup": warmup,
                "tokens_per_step": tok / done,
                "sample": "%d full training steps of the oracle port on %s: torch %s eager, cuDNN nn.LSTM (allow_tf32=%s) + "
                          "cuBLAS matmul (allow_tf32=%s), B=%d, host-timed with synchronize" % (
                              done, torch.cuda.get_device_name(), torch.__version__, torch.backends.cudnn.allow_tf32,
                              torch.backends.cuda.matmul.allow_tf32, bk["batch_size"])}
    return {"tokens_per_s": tok / dt, "ms_per_step": dt / done * 1e3, "steps": done, "warmup": warmup,
            "threads": torch.get_num_threads(), "tokens_per_step": tok / done,
            "sample": "%d full training steps (fwd+loss+bwd+clip+Adam) of the same workload, B=%d, "
                      "torch %s CPU (oneDNN nn.LSTM), %d threads" % (done, bk["batch_size"], torch.__version__,
                                                                     torch.get_num_threads())}
