"""Deterministic, library-version-independent synthetic weights and Multi30k-shaped batches.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Nothing here depends on a torch / numpy RNG stream:
every value is an integer hash (splitmix64 finaliser) of (seed, tensor tag, element index), so the
golden fixtures in tests/golden/ can be regenerated bit-for-bit anywhere, and the fixtures only need
to store *outputs* (the 2048x2048 image-head weights alone would be 16 MB).

Shapes follow the reference constructor ``onmt/ModelConstructor.py:328-620`` (state_dict key names are
the drop-in contract); the batch layout follows SURVEY.md appendix B
(``onmt/io/TextDataset.py:190-205``, ``onmt/TrainerMultimodal.py:632-677``).
"""
from collections import OrderedDict
from dataclasses import dataclass, asdict
import zlib

import numpy as np

PAD, BOS, EOS, UNK = 1, 2, 3, 0   # onmt/io/DatasetBase.py:7-11, IO.py:221-226

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x):
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def uniform01(n, seed, tag):
    """n doubles in [0,1): hash of (seed, crc32(tag), index)."""
    key = (int(seed) * 0x9E3779B97F4A7C15 + zlib.crc32(tag.encode()) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(key)
    h = _mix(idx)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def normal01(n, seed, tag):
    """n standard normals by Box-Muller over two hash streams."""
    u1 = uniform01(n, seed, tag + "/u1")
    u2 = uniform01(n, seed, tag + "/u2")
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


@dataclass
class ModelConfig:
    """Hyper-parameters on the hot path (opts.py:14-17,67-70,477-538; run_translated_m30k_only.sh:53-81)."""
    v_src: int = 10000
    v_tgt: int = 10000
    emb: int = 500          # -src/tgt_word_vec_size
    hidden: int = 500       # -rnn_size
    z_dim: int = 500        # --z_latent_dim
    img_dim: int = 2048     # ModelConstructor.py:350-354
    layers: int = 2
    conditional: bool = True
    dropout: float = 0.0
    brnn: bool = False      # -encoder_type brnn (opts.py:54-58): bidirectional source encoder, H/2 per direction

    def to_dict(self):
        return asdict(self)


CFG1 = ModelConfig()                                           # BASELINE.json configs[0]/[1]
CFG_FIXED = ModelConfig(z_dim=50, conditional=False)           # run_additional_data.sh:47
CFG5 = ModelConfig(v_src=32000, v_tgt=32000, emb=1024, hidden=1024, z_dim=1024)
TINY = ModelConfig(v_src=120, v_tgt=150, emb=32, hidden=64, z_dim=24)
TINY_FIXED = ModelConfig(v_src=120, v_tgt=150, emb=32, hidden=64, z_dim=24, conditional=False)
TINY_BRNN = ModelConfig(v_src=120, v_tgt=150, emb=32, hidden=64, z_dim=24, brnn=True)


def param_shapes(cfg):
    """Ordered {state_dict key: shape}; ``encoder_tgt.embeddings`` aliases ``decoder.embeddings``
    (ModelConstructor.py:456-457) and is therefore not listed separately."""
    E, H, Z, D, L = cfg.emb, cfg.hidden, cfg.z_dim, cfg.img_dim, cfg.layers
    s = OrderedDict()
    s["encoder.embeddings.make_embedding.emb_luts.0.weight"] = (cfg.v_src, E)
    He = H // 2 if cfg.brnn else H                 # Models.py:107-109: hidden_size // num_directions
    for l in range(L):
        i = E if l == 0 else H
        for sfx in (("", "_reverse") if cfg.brnn else ("",)):
            s[f"encoder.rnn.weight_ih_l{l}{sfx}"] = (4 * He, i)
            s[f"encoder.rnn.weight_hh_l{l}{sfx}"] = (4 * He, He)
            s[f"encoder.rnn.bias_ih_l{l}{sfx}"] = (4 * He,)
            s[f"encoder.rnn.bias_hh_l{l}{sfx}"] = (4 * He,)
    s["decoder.embeddings.make_embedding.emb_luts.0.weight"] = (cfg.v_tgt, E)
    for l in range(L):
        i = E + Z if l == 0 else H
        s[f"decoder.rnn.weight_ih_l{l}"] = (4 * H, i)
        s[f"decoder.rnn.weight_hh_l{l}"] = (4 * H, H)
        s[f"decoder.rnn.bias_ih_l{l}"] = (4 * H,)
        s[f"decoder.rnn.bias_hh_l{l}"] = (4 * H,)
    s["decoder.attn.linear_in.weight"] = (H, H)
    s["decoder.attn.linear_out.weight"] = (H, 2 * H)
    if cfg.conditional:
        Hd = H // 2
        for l in range(L):
            i = E if l == 0 else H
            for sfx in ("", "_reverse"):
                s[f"encoder_tgt.rnn.weight_ih_l{l}{sfx}"] = (4 * Hd, i)
                s[f"encoder_tgt.rnn.weight_hh_l{l}{sfx}"] = (4 * Hd, Hd)
                s[f"encoder_tgt.rnn.bias_ih_l{l}{sfx}"] = (4 * Hd,)
                s[f"encoder_tgt.rnn.bias_hh_l{l}{sfx}"] = (4 * Hd,)
    inf_in = 2 * H + D if cfg.conditional else H
    nets = [("inf_net_global", inf_in, Z, Z)]
    if cfg.conditional:
        nets.append(("gen_net_global", H, Z, Z))
    nets.append(("inf_net_image", Z, D, D))
    for name, i, h, o in nets:
        for br in ("location", "scale"):
            s[f"{name}.{br}.fc1.weight"] = (h, i)
            s[f"{name}.{br}.fc1.bias"] = (h,)
            s[f"{name}.{br}.fc2.weight"] = (o, h)
            s[f"{name}.{br}.fc2.bias"] = (o,)
    s["inf_net_image.gate_affine_transform.weight"] = (1, Z)
    s["inf_net_image.gate_affine_transform.bias"] = (1,)
    s["generator.0.weight"] = (cfg.v_tgt, H)
    s["generator.0.bias"] = (cfg.v_tgt,)
    return s


def make_params(cfg, seed=3435, scale=0.1, dtype=np.float32):
    """uniform(-scale, scale) for every tensor (ModelConstructor.py:598-603, opts.py:249)."""
    out = OrderedDict()
    for name, shp in param_shapes(cfg).items():
        n = int(np.prod(shp))
        out[name] = ((uniform01(n, seed, name) * 2.0 - 1.0) * scale).astype(dtype).reshape(shp)
    return out


def num_params(cfg):
    return int(sum(int(np.prod(s)) for s in param_shapes(cfg).values()))


@dataclass
class Batch:
    """What TrainerMultimodal hands to the model (appendix B): ids time-major, int64."""
    src: np.ndarray          # [S, B]
    src_lengths: np.ndarray  # [B]  sorted descending
    tgt: np.ndarray          # [Tf, B]  <s> ... </s> pad
    tgt_lengths: np.ndarray  # [B]  counts BOS and EOS
    img_feats: np.ndarray    # [B, D] fp32, non-negative (post-ReLU pool5)
    eps: np.ndarray          # [B, Z] injected N(0,1) noise for z = mu + sigma*eps

    @property
    def batch_size(self):
        return int(self.src.shape[1])

    @property
    def n_tgt_tokens(self):
        return int((self.tgt[1:] != PAD).sum())


def make_batch(cfg, batch_size=40, seed=0, full_length=None, src_max=50, tgt_max=50,
               s_force=None, t_force=None):
    """Multi30k-shaped batch.

    ``full_length=(S, n_tgt_words)`` gives the all-equal-length variant (cfg1: (30, 30) -> tgt_len 32);
    otherwise lengths ~ clip(round(N(14,5)), 3, max).  ``s_force`` / ``t_force`` set the longest
    source / one target (incl. BOS, EOS) to that length: the reference needs the padded extent to
    equal the longest row (sequence_mask, onmt/Utils.py:23-33)."""
    B = batch_size
    if full_length is not None:
        sl = np.full(B, int(full_length[0]), np.int64)
        tl_words = np.full(B, int(full_length[1]), np.int64)
    else:
        sl = np.clip(np.rint(14 + 5 * normal01(B, seed, "srclen")), 3, src_max).astype(np.int64)
        tl_words = np.clip(np.rint(14 + 5 * normal01(B, seed, "tgtlen")), 3, tgt_max - 2).astype(np.int64)
        if s_force is not None:
            sl = np.minimum(sl, s_force); sl[0] = s_force
        if t_force is not None:
            tl_words = np.minimum(tl_words, t_force - 2); tl_words[B // 2] = t_force - 2
        order = np.argsort(-sl, kind="stable")          # rows sorted by src length, descending
        sl, tl_words = sl[order], tl_words[order]
    tl = tl_words + 2
    S, Tf = int(sl.max()), int(tl.max())
    src = np.full((S, B), PAD, np.int64)
    tgt = np.full((Tf, B), PAD, np.int64)
    us = uniform01(S * B, seed, "srctok").reshape(S, B)
    ut = uniform01(Tf * B, seed, "tgttok").reshape(Tf, B)
    src_tok = 4 + np.floor(us * (cfg.v_src - 4)).astype(np.int64)
    tgt_tok = 4 + np.floor(ut * (cfg.v_tgt - 4)).astype(np.int64)
    for b in range(B):
        src[: sl[b], b] = src_tok[: sl[b], b]
        tgt[0, b] = BOS
        tgt[1: tl[b] - 1, b] = tgt_tok[1: tl[b] - 1, b]
        tgt[tl[b] - 1, b] = EOS
    img = (np.abs(normal01(B * cfg.img_dim, seed, "img")) * 0.5).astype(np.float32).reshape(B, cfg.img_dim)
    eps = normal01(B * cfg.z_dim, seed, "eps").astype(np.float32).reshape(B, cfg.z_dim)
    return Batch(src, sl, tgt, tl, img, eps)
