"""Shared helpers for the -m gpu parity tests: build the CUDA model from the oracle's deterministic
weights, move a synthetic batch to the device, compare against oracle outputs."""
import numpy as np
import torch

from oracle import synth
from oracle import vi_model1_ref as R


def build_cuda_model(cfg, params, dropout=0.0):
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import synthetic
    opt = synthetic.make_opt(emb=cfg.emb, hidden=cfg.hidden, z_dim=cfg.z_dim, layers=cfg.layers,
                             conditional=cfg.conditional, dropout=dropout,
                             encoder_type="brnn" if getattr(cfg, "brnn", False) else "rnn")
    fields = synthetic.make_fields(cfg.v_src, cfg.v_tgt)
    model = vm.make_vi_model_mmt(opt, fields, gpu=True)
    sd = model.state_dict()
    new = {}
    for k in sd:
        src = k.replace("encoder_tgt.embeddings", "decoder.embeddings")
        new[k] = torch.as_tensor(params[src])
    model.load_state_dict(new)
    return model, fields


class DevBatch:
    pass


def to_device(batch, dev="cuda"):
    b = DevBatch()
    b.src = torch.as_tensor(batch.src).unsqueeze(2).to(dev)
    b.src_lengths = torch.as_tensor(batch.src_lengths).to(dev)
    b.tgt_ids = torch.as_tensor(batch.tgt).to(dev)
    b.tgt_in = b.tgt_ids.unsqueeze(2)
    b.tgt = b.tgt_ids                      # what the loss reads (TrainerMultimodal.py:668-677)
    b.tgt_lengths = torch.as_tensor(batch.tgt_lengths).to(dev)
    b.img_feats = torch.as_tensor(batch.img_feats).to(dev)
    b.eps = torch.as_tensor(batch.eps).to(dev)
    b.batch_size = batch.batch_size
    return b


def relerr(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def maxabs(a, b):
    return float(np.abs(np.asarray(a, np.float64).reshape(-1) - np.asarray(b, np.float64).reshape(-1)).max())


def named_grads(model):
    out = {}
    for k, p in model.named_parameters():
        out[k] = None if p.grad is None else p.grad.detach().cpu().numpy().copy()
    return out


def attn_max_rel(a, ref, src_lengths):
    """north_star's attention criterion: max over UNMASKED entries of |a - ref| / ref (a, ref: [T, B, S])."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    live = np.arange(a.shape[2])[None, None, :] < np.asarray(src_lengths)[None, :, None]
    live = np.broadcast_to(live, a.shape)
    assert not np.any(a[~live]), "attention mass on padding"
    return float((np.abs(a - ref)[live] / ref[live]).max())
