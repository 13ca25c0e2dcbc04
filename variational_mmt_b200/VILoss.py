"""Loss computation for VI model 1 (reference: onmt/VILoss.py:59-531, onmt/Loss.py:68-132,218-273,
onmt/TrainerMultimodal.py:32-228 for VIStatistics).

Same constructor, attributes (``padding_idx``, ``kl_annealing_current``, ``n_model_updates``) and
methods.  ``_compute_loss`` evaluates generator + log-softmax + NLL, the image log-prob / cosine and
the analytic KL in libvmmt kernels (one autograd node, ``ops.VILossFn``); the M x V score matrix is
never handed back to Python, the statistics come out of the same kernels.

Reference behaviours kept on purpose (hazards H3-H5):
  * training (``sharded_compute_loss``) scores only decoder positions [0, shard_size) -- the zip in
    onmt/Loss.py:shards() truncates to the first shard because the latent tensors have one row;
  * total = NLL_sum + (-image log-prob) + kl_weight * KL, divided by ``normalization``;
  * the image log-prob is taken on L2-normalised prediction / observation with unit scale, and its
    gradient reaches the prediction head by the legacy pass-through (switch: ``image_grad``).
"""
import math
import sys
import time

import torch
import torch.nn as nn

from . import ops

PAD_WORD = "<blank>"


class VIStatistics(object):
    """Accumulator for the loss statistics (onmt/TrainerMultimodal.py:32-228).  Values live in one
    device vector and are only copied to the host when read."""
    _KEYS = ("nmt_loss", "n_words", "n_correct", "td_kl_before", "image_feats_loss", "image_feats_cos",
             "td_kl_after", "elbo_loss")

    def __init__(self, multimodal_model_type="vi-model1", loss_data=None, n_words=0, n_correct=0):
        self.multimodal_model_type = multimodal_model_type
        self.progress_state_train, self.progress_state_valid = [], []
        self._vec = None if loss_data is None else loss_data["_vec"]
        self.two_step_image_prediction = False
        self.image_loss_type = loss_data["image_loss_type"] if loss_data else "logprob"
        self.td_kl_multiplier = loss_data["td_kl_multiplier"] if loss_data else 1.0
        self.image_pixels_loss = self.image_pixels_acc = self.img_pixels_acc = 0.0
        self.te_kl = self.td_kl = 0.0
        self.n_src_words = 0
        self.n_updates = 0
        self.start_time = time.time()

    def _get(self, i):
        return 0.0 if self._vec is None else float(self._vec[i])

    nmt_loss = property(lambda s: s._get(0))
    n_words = property(lambda s: int(round(s._get(1))))
    n_correct = property(lambda s: int(round(s._get(2))))
    td_kl_before = property(lambda s: s._get(3))
    image_feats_loss = property(lambda s: s._get(4))
    image_feats_cos = property(lambda s: s._get(5))
    td_kl_after = property(lambda s: s._get(6))
    elbo_loss = property(lambda s: s._get(7))

    def update(self, stat):
        if stat._vec is not None:
            self._vec = stat._vec.clone() if self._vec is None else self._vec.add_(stat._vec)
        self.td_kl_multiplier = stat.td_kl_multiplier
        self.n_updates += 1

    def accuracy(self):
        return 100 * (self.n_correct / max(self.n_words, 1))

    def ppl(self):
        return math.exp(min(self.nmt_loss / max(self.n_words, 1), 100))

    def elapsed_time(self):
        return time.time() - self.start_time

    def output(self, epoch, batch, n_batches, start):
        t = self.elapsed_time()
        n = max(self.n_updates, 1)
        print(("Epoch %2d, %5d/%5d; acc: %6.2f; ppl: %6.2f; td-kl-before (avg.): %6.2f; "
               "td-kl-after (avg.): %6.2f; td-kl-multiplier: %2.2f;img-feats-loss (avg.): %6.2f; "
               "img-feats-cos (avg.): %6.2f; elbo (avg.): %6.2f; %3.0f src tok/s; %3.0f tgt tok/s; "
               "%6.0f s elapsed") %
              (epoch, batch, n_batches, self.accuracy(), self.ppl(), self.td_kl_before / n,
               self.td_kl_after / n, self.td_kl_multiplier, self.image_feats_loss / n,
               self.image_feats_cos / n, self.elbo_loss / n, float(self.n_src_words) / (t + 1e-5),
               self.n_words / (t + 1e-5), time.time() - start))
        self.n_updates = 0
        sys.stdout.flush()

    def save_progress(self, lr, model_updates, epoch, split):
        assert split in ("train", "valid")
        n = max(self.n_updates, 1)
        t = self.elapsed_time()
        progress = {"epoch": epoch, "model_updates": model_updates, "elapsed_time": t, "ppl": self.ppl(),
                    "acc": self.accuracy(), "image_feats_loss": self.image_feats_loss / n,
                    "image_feats_cos": self.image_feats_cos / n, "img_pixels_loss": 0.0,
                    "img_pixels_acc": 0.0, "td_kl_before": self.td_kl_before / n,
                    "td_kl_after": self.td_kl_after / n, "td_kl": 0.0,
                    "td_kl_multiplier": float(self.td_kl_multiplier), "elbo": self.elbo_loss / n,
                    "tgt_per": self.n_words / max(t, 1e-9), "lr": lr}
        (self.progress_state_train if split == "train" else self.progress_state_valid).append(progress)


class NMTVIModel1LossCompute(nn.Module):
    def __init__(self, generator, tgt_vocab, normalization="sents", label_smoothing=0.0,
                 use_kl_annealing=False, use_kl_freebits=False, kl_freebits_margin=0.0,
                 kl_annealing_current=0.0, kl_annealing_increment=0.0001, kl_annealing_warmup_steps=1000,
                 image_loss_type="logprob", use_local_image_features=False,
                 two_step_image_prediction=False, image_grad="legacy_passthrough"):
        super().__init__()
        assert label_smoothing == 0.0, "label smoothing is 0 in the published configurations"
        assert image_loss_type == "logprob" and not use_local_image_features and not two_step_image_prediction
        assert not use_kl_freebits, "free bits are off in the published configurations (opts.py:498-509)"
        assert image_grad in ("legacy_passthrough", "true_jacobian")
        self.multimodal_model_type = "vi-model1"
        self.generator = generator
        self.tgt_vocab = tgt_vocab
        self.padding_idx = tgt_vocab.stoi[PAD_WORD]
        self.confidence = 1.0
        self.n_model_updates = 0
        self.use_kl_annealing = use_kl_annealing
        if use_kl_annealing:
            self.kl_annealing_current = kl_annealing_current
            self.kl_annealing_increment = kl_annealing_increment
            self.kl_annealing_warmup_steps = kl_annealing_warmup_steps
        else:
            self.kl_annealing_current, self.kl_annealing_increment, self.kl_annealing_warmup_steps = 1.0, 0.0, 0
        self.use_kl_freebits, self.kl_freebits_margin = False, 0.0
        self.image_loss_type = image_loss_type
        self.use_local_image_features = False
        self.two_step_image_prediction = False
        self.image_grad = image_grad
        self._statistics = VIStatistics

    # -- onmt/VILoss.py:121-215
    def _make_shard_state(self, batch, output, range_, attns):
        q, p = attns["z_latent"][0], attns["p_latent"][0]
        loc, scale = q.params()
        standard_prior = getattr(p, "is_standard", False)
        p_loc, p_scale = (None, None) if standard_prior else p.params()
        pv = attns["p_global_image_features"][0]
        return {"output": output, "target": batch.tgt[range_[0] + 1: range_[1]],
                "qz_location": loc.unsqueeze(0), "qz_scale": scale.unsqueeze(0),
                "pz_location": None if p_loc is None else p_loc.unsqueeze(0),
                "pz_scale": None if p_scale is None else p_scale.unsqueeze(0),
                "p_global_image_features_location": pv.mean().unsqueeze(0),
                "ground_truth_global_image_features": attns["ground_truth_global_image_features"][0].unsqueeze(0)}

    # -- onmt/VILoss.py:217-513
    def _compute_loss(self, batch, output, target, qz_location, qz_scale, pz_location, pz_scale,
                      p_global_image_features_location=None, p_global_image_features_scale=None,
                      ground_truth_global_image_features=None, p_image_pixels_location=None,
                      ground_truth_image_pixels=None, p_image_pixels_scale=None):
        lin = self.generator[0]
        kw = self.kl_annealing_current if self.use_kl_annealing else 1.0
        cfg = {"pad_idx": self.padding_idx, "kl_weight": kw,
               "legacy_image_grad": self.image_grad == "legacy_passthrough"}
        loss, stats = ops.vi_loss(
            output.reshape(-1, output.size(2)), target.reshape(-1), lin.weight, lin.bias,
            qz_location.squeeze(0), qz_scale.squeeze(0),
            None if pz_location is None else pz_location.squeeze(0),
            None if pz_scale is None else pz_scale.squeeze(0),
            p_global_image_features_location.squeeze(0), ground_truth_global_image_features.squeeze(0), cfg)
        vec = stats                                    # [6] = td_kl_after, [7] = elbo filled by vmmt_loss_finalize
        loss_data = {"_vec": vec, "two_step_image_prediction": False, "image_loss_type": self.image_loss_type,
                     "td_kl_multiplier": self.kl_annealing_current}
        batch_stats = VIStatistics(self.multimodal_model_type, loss_data)
        # annealing schedule (onmt/VILoss.py:501-511)
        if self.kl_annealing_current > 1.0:
            self.kl_annealing_current = 1.0
        if self.kl_annealing_current < 1.0 and self.n_model_updates >= self.kl_annealing_warmup_steps:
            self.kl_annealing_current += self.kl_annealing_increment
        self.n_model_updates += 1
        return loss, batch_stats

    # -- onmt/Loss.py:68-86
    def monolithic_compute_loss(self, batch, output, attns):
        range_ = (0, batch.tgt.size(0))
        with torch.no_grad():
            _, stats = self._compute_loss(batch, **self._make_shard_state(batch, output, range_, attns))
        return stats

    # -- onmt/Loss.py:88-132 + shards() 226-273
    def sharded_compute_loss(self, batch, output, attns, cur_trunc, trunc_size, shard_size, normalization):
        """Forward + backward.  Every shard-state tensor is split by ``shard_size`` on dim 0 and the
        shards are zipped; the latent / image tensors have a single row, so only the first shard of
        ``output`` / ``target`` is ever scored (hazard H4) -- reproduced here by slicing."""
        batch_stats = VIStatistics(self.multimodal_model_type)
        range_ = (cur_trunc, cur_trunc + trunc_size)
        state = self._make_shard_state(batch, output, range_, attns)
        if state["output"].size(0) > shard_size:           # (a no-op slice would still cost a zero-fill + copy in backward)
            state["output"] = state["output"][:shard_size]
            state["target"] = state["target"][:shard_size]
        loss, stats = self._compute_loss(batch, **state)
        ops.stamp(5)
        hook = getattr(self, "before_backward", None)      # GraphedTrainStep: join the gradient memset issued beside the loss
        if hook is not None:
            hook()
            self.before_backward = None
        loss.div(normalization).backward()
        ops.stamp(15)                   # backward's main chain done (measurement aid)
        ops.join_side()                 # weight-gradient GEMMs issued on the side stream are complete from here on
        ops.stamp(16)
        batch_stats.update(stats)
        return batch_stats
