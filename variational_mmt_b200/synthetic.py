"""Synthetic Multi30k-shaped workloads and the option / vocabulary objects the constructor expects.

The surfdrive Multi30k tarballs are not available offline; throughput is measured on batches of the
same shape: BPE-sized vocabularies, ~14-token sentences (or the all-30-token
variant), 2048-d non-negative pooled image features.  Token ids follow the reference's special
symbols (onmt/io/DatasetBase.py:7-11): <unk>=0 <blank>=1 <s>=2 </s>=3 on the target side.
"""
from types import SimpleNamespace

import numpy as np
import torch

PAD, BOS, EOS = 1, 2, 3


class Vocab(object):
    """Stand-in for torchtext.vocab.Vocab: ``stoi``, ``itos``, ``len()``."""

    def __init__(self, n, specials):
        self.itos = list(specials) + ["w%d" % i for i in range(n - len(specials))]
        self.stoi = {w: i for i, w in enumerate(self.itos)}

    def __len__(self):
        return len(self.itos)


class Field(object):
    def __init__(self, vocab):
        self.vocab = vocab


def make_fields(v_src, v_tgt):
    return {"src": Field(Vocab(v_src, ["<unk>", "<blank>"])),
            "tgt": Field(Vocab(v_tgt, ["<unk>", "<blank>", "<s>", "</s>"]))}


def make_opt(emb=500, hidden=500, z_dim=500, layers=2, conditional=True, dropout=0.5, encoder_type="rnn",
             param_init=0.1):
    """The model options of run_translated_m30k_only.sh:53-81 + opts.py defaults that the constructor reads."""
    return SimpleNamespace(
        model_type="text", multimodal_model_type="vi-model1", src_word_vec_size=emb, tgt_word_vec_size=emb,
        rnn_type="LSTM", rnn_size=hidden, enc_layers=layers, dec_layers=layers, encoder_type=encoder_type,
        brnn=(encoder_type == "brnn"), dropout=dropout, word_dropout=0.0, global_attention="general",
        coverage_attn=False, context_gate=None, copy_attn=False, reuse_copy_attn=False,
        z_latent_dim=z_dim, conditional=conditional, use_global_image_features=True,
        use_posterior_image_features=False, use_local_image_features=False, image_loss="logprob",
        path_to_train_img_feats="resnet50.hdf5", param_init=param_init, share_embeddings=False,
        share_decoder_embeddings=False)


class Batch(object):
    """The attributes the trainer / loss read."""

    def __init__(self, src, src_lengths, tgt, tgt_lengths, img_feats, indices=None):
        self.src = (src, src_lengths)
        self.tgt = tgt                      # the trainer overwrites batch.tgt with the id tensor
        self.tgt_lengths = tgt_lengths
        self.img_feats = img_feats
        self.batch_size = src.shape[1]
        self.indices = indices


def random_batch(v_src=10000, v_tgt=10000, batch_size=40, img_dim=2048, seed=0, full_length=None,
                 src_max=50, tgt_max=50, pinned=False):
    """Host (CPU) tensors: src [S,B], src_lengths [B] (descending), tgt [Tf,B], tgt_lengths [B]
    (incl. BOS/EOS), img_feats [B,D]."""
    rng = np.random.RandomState(seed)
    B = batch_size
    if full_length is not None:
        sl = np.full(B, full_length[0], np.int64)
        tw = np.full(B, full_length[1], np.int64)
    else:
        sl = np.clip(np.rint(rng.normal(14, 5, B)), 3, src_max).astype(np.int64)
        tw = np.clip(np.rint(rng.normal(14, 5, B)), 3, tgt_max - 2).astype(np.int64)
        order = np.argsort(-sl, kind="stable")
        sl, tw = sl[order], tw[order]
    tl = tw + 2
    S, Tf = int(sl.max()), int(tl.max())
    src = np.full((S, B), PAD, np.int64)
    tgt = np.full((Tf, B), PAD, np.int64)
    for b in range(B):
        src[:sl[b], b] = rng.randint(4, v_src, sl[b])
        tgt[0, b] = BOS
        tgt[1:tl[b] - 1, b] = rng.randint(4, v_tgt, tw[b])
        tgt[tl[b] - 1, b] = EOS
    img = (np.abs(rng.normal(0, 1, (B, img_dim))) * 0.5).astype(np.float32)
    out = [torch.from_numpy(a) for a in (src, sl, tgt, tl, img)]
    if pinned and torch.cuda.is_available():
        out = [t.pin_memory() for t in out]
    return tuple(out)
