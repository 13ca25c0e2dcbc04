"""CPU restatement of one VI-model-1 step (forward, loss, gradients, clip+Adam).

TEST INFRASTRUCTURE (see oracle/__init__.py): checker and reported CPU baseline only.

Plain PyTorch on the host, explicit time loops, any float dtype (fp32 = what the reference computes
in; fp64 = tolerance budgeting).  Every function cites the reference lines it restates; paths are
relative to /root/reference.  The restatement is pinned to the executed reference by
tests/golden/*.npz (oracle/make_golden.py, tests/test_oracle_golden.py).

Parameter dict keys are the reference state_dict keys (oracle/synth.py:param_shapes).
"""
import math

import torch
import torch.nn.functional as F

PAD = 1


def _t(params, dtype=torch.float32, requires_grad=False):
    out = {}
    for k, v in params.items():
        t = torch.as_tensor(v).to(dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------
def lstm_layer(x, h0, c0, w_ih, w_hh, b_ih, b_hh, lengths=None, reverse=False):
    """One LSTM layer, one direction; gate order i,f,g,o (torch nn.LSTM, onmt/Models.py:124-129).

    ``lengths`` reproduces pack_padded_sequence / pad_packed_sequence (onmt/Models.py:139-147):
    past a row's length the state is frozen and the output is zero."""
    T = x.shape[0]
    h, c = h0, c0
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = x[t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, gg, o = g.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        if lengths is not None:
            m = (t < lengths).to(x.dtype).unsqueeze(1)
            c = m * c_new + (1 - m) * c
            h = m * h_new + (1 - m) * h
            outs[t] = m * h_new
        else:
            c, h = c_new, h_new
            outs[t] = h_new
    return torch.stack(outs), h, c


def lstm_stack(x, p, prefix, layers, h0=None, c0=None, lengths=None, bidirectional=False,
               drop=None):
    """Stacked (optionally bidirectional) LSTM; ``drop(x, site)`` is applied to the input of every
    layer but the first (nn.LSTM ``dropout=``)."""
    N = x.shape[1]
    hs, cs = [], []
    for l in range(layers):
        if l > 0 and drop is not None:
            x = drop(x, f"{prefix}.l{l}")
        dirs = ("", "_reverse") if bidirectional else ("",)
        outs = []
        for d, sfx in enumerate(dirs):
            w_ih, w_hh = p[f"{prefix}.weight_ih_l{l}{sfx}"], p[f"{prefix}.weight_hh_l{l}{sfx}"]
            b_ih, b_hh = p[f"{prefix}.bias_ih_l{l}{sfx}"], p[f"{prefix}.bias_hh_l{l}{sfx}"]
            Hd = w_hh.shape[1]
            idx = l * len(dirs) + d
            hh = h0[idx] if h0 is not None else x.new_zeros(N, Hd)
            cc = c0[idx] if c0 is not None else x.new_zeros(N, Hd)
            o, h, c = lstm_layer(x, hh, cc, w_ih, w_hh, b_ih, b_hh, lengths, reverse=(d == 1))
            outs.append(o); hs.append(h); cs.append(c)
        x = outs[0] if len(outs) == 1 else torch.cat(outs, 2)
    return x, torch.stack(hs), torch.stack(cs)


def masked_mean(x, lengths):
    """GlobalInferenceNetwork.encode_seq (onmt/modules/NormalVariationalEncoder.py:65-84)."""
    T = x.shape[0]
    m = (torch.arange(T, device=x.device).unsqueeze(1) < lengths.unsqueeze(0)).to(x.dtype).unsqueeze(2)
    return (x * m).sum(0) / m.sum(0)


def mlp2(p, prefix, x, softplus=False):
    """LocationLayer / ScaleLayer (NormalVariationalEncoder.py:12-25, 29-43)."""
    h = F.relu(x @ p[prefix + ".fc1.weight"].t() + p[prefix + ".fc1.bias"])
    y = h @ p[prefix + ".fc2.weight"].t() + p[prefix + ".fc2.bias"]
    return F.softplus(y) if softplus else y


def global_attention(q, ctx, lengths, w_in, w_out):
    """GlobalAttention 'general' (onmt/modules/GlobalAttention.py:108-113,169-190,204-205).
    q [T,B,H], ctx [S,B,H] time-major -> attn_h [T,B,H], align [T,B,S]."""
    S = ctx.shape[0]
    qb, cb = q.transpose(0, 1), ctx.transpose(0, 1)
    scores = (qb @ w_in.t()) @ cb.transpose(1, 2)                   # [B,T,S]
    if lengths is not None:
        mask = torch.arange(S, device=q.device).unsqueeze(0) < lengths.unsqueeze(1)  # [B,S]
        scores = scores.masked_fill(~mask.unsqueeze(1), float("-inf"))
    align = scores.softmax(-1)
    c = align @ cb
    out = torch.tanh(torch.cat([c, qb], 2) @ w_out.t())
    return out.transpose(0, 1).contiguous(), align.transpose(0, 1).contiguous()


def image_head(p, z):
    """ImageGlobalInferenceNetwork.forward, use_source_encodings=False
    (NormalVariationalEncoder.py:286-304); the scale branch is computed lazily by callers."""
    g = torch.sigmoid(z @ p["inf_net_image.gate_affine_transform.weight"].t()
                      + p["inf_net_image.gate_affine_transform.bias"])
    return mlp2(p, "inf_net_image.location", z * g), g


def kl_normal(mu_q, sd_q, mu_p, sd_p):
    """VILoss.py:439-460: sum over Z, mean over B."""
    t = 0.5 / sd_p ** 2 * ((mu_q - mu_p) ** 2 + sd_q ** 2 - sd_p ** 2) + sd_p.log() - sd_q.log()
    return t.sum(1).mean()


def image_terms(loc, v, legacy_grad=True):
    """Cosine (reported) and the aliased log-prob (hazard H3; VILoss.py:22-56,289-296,317-332).

    Value: IMG = sum_b mean_d [ -(p^-v^)^2/2 - log(2 pi)/2 ] with p^, v^ L2-normalised rows.
    Gradient: ``legacy_grad`` passes dIMG/dp^ straight to ``loc`` (torch-0.3.1 aliasing);
    otherwise the true Jacobian of the normalisation is used."""
    p_hat = loc / loc.norm(dim=1, keepdim=True)
    v_hat = v / v.norm(dim=1, keepdim=True)
    cos = (p_hat * v_hat).sum(1).mean().detach()
    p_in = loc + (p_hat - loc).detach() if legacy_grad else p_hat
    lp = (-0.5 * (p_in - v_hat) ** 2 - 0.5 * math.log(2 * math.pi)).sum(0).mean()
    return lp, cos


# ---------------------------------------------------------------------------------------------
# the model
# ---------------------------------------------------------------------------------------------
def forward(p, cfg, batch, training=False, eps=None, drop=None, z_override=None):
    """NMTVIModel.forward (onmt/Models.py:850-1011) + StdRNNVIModel1Decoder._run_forward_pass
    (onmt/VI_Model1.py:51-135).  ``batch`` fields are torch tensors (src [S,B], tgt [Tf,B], ...)."""
    src, lengths, tgt, tgt_lengths, v = batch["src"], batch["src_lengths"], batch["tgt"], \
        batch["tgt_lengths"], batch["img_feats"]
    L = cfg.layers
    # 1 encoder (Models.py:131-149)
    x = p["encoder.embeddings.make_embedding.emb_luts.0.weight"][src]
    brnn = getattr(cfg, "brnn", False)
    ctx, h_enc, c_enc = lstm_stack(x, p, "encoder.rnn", L, lengths=lengths, bidirectional=brnn, drop=drop)
    if brnn:
        # RNNVIDecoderBase._fix_enc_hidden (Models.py:1158-1165): [layers*2, B, H/2] -> [layers, B, H], fwd ; bwd per layer
        h_enc = torch.cat([h_enc[0::2], h_enc[1::2]], 2)
        c_enc = torch.cat([c_enc[0::2], c_enc[1::2]], 2)
    out = {"context": ctx, "enc_h": h_enc, "enc_c": c_enc}
    hx = masked_mean(ctx, lengths)
    if cfg.conditional:
        # 2 prior p(z|x) on the live context (Models.py:889)
        mu_p = mlp2(p, "gen_net_global.location", hx)
        sd_p = mlp2(p, "gen_net_global.scale", hx, softplus=True)
        # 3 posterior q(z|x,y,v): target encoder recurs over the BATCH axis (H1, Models.py:892-893)
        y = p["decoder.embeddings.make_embedding.emb_luts.0.weight"][tgt].transpose(0, 1)
        yctx, _, _ = lstm_stack(y, p, "encoder_tgt.rnn", L, bidirectional=True, drop=drop)
        yctx = yctx.transpose(0, 1)
        hy = masked_mean(yctx, tgt_lengths)
        hq = torch.cat([masked_mean(ctx.detach(), lengths), hy, v], 1)   # Models.py:911
        mu_q = mlp2(p, "inf_net_global.location", hq)
        sd_q = mlp2(p, "inf_net_global.scale", hq, softplus=True)
        out["tgt_context"] = yctx
    else:
        hq = masked_mean(ctx.detach(), lengths)                          # Models.py:930
        mu_q = mlp2(p, "inf_net_global.location", hq)
        sd_q = mlp2(p, "inf_net_global.scale", hq, softplus=True)
        mu_p, sd_p = torch.zeros_like(mu_q), torch.ones_like(mu_q)       # Models.py:936-939
    # 4 sample: torch.normal has no pathwise gradient (H2; Dists.py:21-26)
    if z_override is not None:
        z = z_override
    elif training:
        z = (mu_q + sd_q * eps).detach()
    else:
        z = (mu_p if cfg.conditional else mu_q).detach()
    # 5 decoder (VI_Model1.py:94-106), initial state = encoder final state (Models.py:1167-1174)
    e = p["decoder.embeddings.make_embedding.emb_luts.0.weight"][tgt[:-1]]
    u = torch.cat([e, z.unsqueeze(0).expand(e.shape[0], -1, -1)], 2)
    q, h_dec, c_dec = lstm_stack(u, p, "decoder.rnn", L, h0=h_enc, c0=c_enc, drop=drop)
    # 6 attention + output dropout (VI_Model1.py:117-132)
    attn_h, align = global_attention(q, ctx, lengths, p["decoder.attn.linear_in.weight"],
                                     p["decoder.attn.linear_out.weight"])
    if drop is not None:
        attn_h = drop(attn_h, "decoder.out")
    # 7 image head (Models.py:986)
    loc_v, gate = image_head(p, z)
    out.update(dict(out=attn_h, attn=align, rnn_out=q, dec_h=h_dec, dec_c=c_dec, z=z,
                    mu_q=mu_q, sd_q=sd_q, mu_p=mu_p, sd_p=sd_p, img_loc=loc_v, img_gate=gate))
    return out


def log_probs(p, x):
    """Generator = Linear + LogSoftmax (ModelConstructor.py:582-585)."""
    return F.log_softmax(x @ p["generator.0.weight"].t() + p["generator.0.bias"], dim=-1)


def compute_loss(p, cfg, fwd, batch, shard_size=None, kl_weight=1.0, legacy_image_grad=True):
    """NMTVIModel1LossCompute._compute_loss (onmt/VILoss.py:217-513).

    ``shard_size`` = 32 reproduces the training path's truncation to the first shard (H4,
    onmt/Loss.py:226-273); None = monolithic (validation)."""
    out, tgt = fwd["out"], batch["tgt"][1:]
    if shard_size is not None:
        out, tgt = out[:shard_size], tgt[:shard_size]
    lp = log_probs(p, out.reshape(-1, out.shape[2]))
    tg = tgt.reshape(-1)
    nz = tg.ne(PAD)
    nll = -(lp.gather(1, tg.unsqueeze(1)).squeeze(1) * nz.to(lp.dtype)).sum()
    kl = kl_normal(fwd["mu_q"], fwd["sd_q"], fwd["mu_p"], fwd["sd_p"])
    img_lp, cos = image_terms(fwd["img_loc"], batch["img_feats"], legacy_grad=legacy_image_grad)
    loss = nll - img_lp + kl * kl_weight
    pred = lp.argmax(1)
    stats = dict(nmt=float(nll.detach()), td_kl_before=float(kl.detach()), td_kl_after=float((kl * kl_weight).detach()),
                 img_feats_loss=float(img_lp.detach()), img_feats_cos=float(cos), elbo=float(loss.detach()),
                 n_words=int(nz.sum()), n_correct=int((pred.eq(tg) & nz).sum()))
    return loss, stats, lp


def train_step_grads(params, cfg, batch, normalization=None, dtype=torch.float32, shard_size=32,
                     drop=None, legacy_image_grad=True):
    """forward + sharded loss + backward (Loss.py:88-132): returns (grads, stats, fwd)."""
    p = _t(params, dtype, requires_grad=True)
    b = to_torch_batch(batch, dtype)
    fwd = forward(p, cfg, b, training=True, eps=b["eps"], drop=drop)
    loss, stats, _ = compute_loss(p, cfg, fwd, b, shard_size=shard_size,
                                  legacy_image_grad=legacy_image_grad)
    norm = normalization if normalization is not None else b["src"].shape[1]   # H5: 'sents'
    (loss / norm).backward()
    grads = {k: (v.grad.detach() if v.grad is not None else None) for k, v in p.items()}
    # nn.Embedding(padding_idx=1): the pad row receives no gradient (Embeddings.py:118)
    for k in grads:
        if "emb_luts" in k and grads[k] is not None:
            grads[k][PAD] = 0
    return grads, stats, fwd


def clip_and_adam(params, grads, state, lr=0.002, max_norm=5.0, betas=(0.9, 0.999), eps=1e-9):
    """onmt/Optim.py:69-70,78-96: global-norm clip then Adam(eps=1e-9); tensors without a
    gradient are skipped (H6).  ``state`` = {"step": int, "m": {...}, "v": {...}} updated in place."""
    keys = [k for k in params if grads.get(k) is not None]
    total = math.sqrt(sum(float((grads[k].double() ** 2).sum()) for k in keys))
    coef = max_norm / (total + 1e-6)
    coef = min(coef, 1.0)
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    b1, b2 = betas
    new = dict(params)
    for k in keys:
        g = grads[k] * coef
        m = state.setdefault("m", {}).get(k, torch.zeros_like(g))
        v = state.setdefault("v", {}).get(k, torch.zeros_like(g))
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        state["m"][k], state["v"][k] = m, v
        denom = v.sqrt() / math.sqrt(1 - b2 ** t) + eps
        new[k] = torch.as_tensor(params[k]) - (lr / (1 - b1 ** t)) * m / denom
    return new, total


def to_torch_batch(batch, dtype=torch.float32):
    return dict(src=torch.as_tensor(batch.src), src_lengths=torch.as_tensor(batch.src_lengths),
                tgt=torch.as_tensor(batch.tgt), tgt_lengths=torch.as_tensor(batch.tgt_lengths),
                img_feats=torch.as_tensor(batch.img_feats).to(dtype),
                eps=torch.as_tensor(batch.eps).to(dtype))


def eval_step(params, cfg, batch, dtype=torch.float32):
    """validation path (TrainerMultimodal.py:409-485): eval forward + monolithic loss."""
    with torch.no_grad():
        p = _t(params, dtype)
        b = to_torch_batch(batch, dtype)
        fwd = forward(p, cfg, b, training=False)
        _, stats, lp = compute_loss(p, cfg, fwd, b, shard_size=None)
    return fwd, stats, lp
