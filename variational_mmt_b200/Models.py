"""Encoder, decoder base, decoder state and the VI model (reference: onmt/Models.py:90-149,
576-638, 737-1011, 1014-1174).  Same class names, signatures, attributes and state_dict keys; the
arithmetic runs in libvmmt kernels through ``modules`` / ``ops``.
"""
import os

import torch
import torch.nn as nn

from . import ops
from .flat import FlatParamsMixin
from .modules import (LSTM, GlobalAttention, Normal, GlobalInferenceNetwork,
                      GlobalFullInferenceNetwork)

MODEL_TYPES = ["vi-model1"]          # onmt/Utils.py (MODEL_TYPES); the path built here


class RNNEncoder(nn.Module):
    """LSTM encoder (onmt/Models.py:90-149)."""

    def __init__(self, rnn_type, bidirectional, num_layers, hidden_size, dropout=0.0, embeddings=None):
        super().__init__()
        assert embeddings is not None
        assert rnn_type == "LSTM", "published configurations use -rnn_type LSTM (run_translated_m30k_only.sh:54)"
        ndir = 2 if bidirectional else 1
        assert hidden_size % ndir == 0
        self.embeddings = embeddings
        self.no_pack_padded_seq = False
        self.rnn = LSTM(embeddings.embedding_size, hidden_size // ndir, num_layers=num_layers,
                        dropout=dropout, bidirectional=bidirectional)

    def forward(self, input, lengths=None, hidden=None):
        """input [len, batch, 1] -> (hidden_t = (h_n, c_n), outputs [len, batch, hidden])"""
        s_len, n_batch, _ = input.size()
        if lengths is not None:
            assert lengths.numel() == n_batch
        emb = self.embeddings(input)
        use_len = lengths if (lengths is not None and not self.no_pack_padded_seq) else None
        outputs, hidden_t = self.rnn(emb, hidden, lengths=use_len)
        return hidden_t, outputs


class DecoderState(object):
    def detach(self):
        for h in self._all:
            if h is not None:
                h.detach_()

    def beam_update(self, idx, positions, beam_size):
        for e in self._all:
            a, br, d = e.size()
            sent_states = e.view(a, beam_size, br // beam_size, d)[:, :, idx]
            sent_states.data.copy_(sent_states.data.index_select(1, positions))


class RNNDecoderState(DecoderState):
    """onmt/Models.py:597-638"""

    def __init__(self, context, hidden_size, rnnstate):
        self.hidden = rnnstate if isinstance(rnnstate, tuple) else (rnnstate,)
        self.coverage = None
        batch_size = context.size(1)
        self.input_feed = context.new_zeros(1, batch_size, hidden_size)

    @property
    def _all(self):
        return self.hidden + (self.input_feed,)

    def update_state(self, rnnstate, input_feed, coverage):
        self.hidden = rnnstate if isinstance(rnnstate, tuple) else (rnnstate,)
        self.input_feed = input_feed
        self.coverage = coverage

    def repeat_beam_size_times(self, beam_size):
        vars_ = [e.detach().repeat(1, beam_size, 1) for e in self._all]
        self.hidden = tuple(vars_[:-1])
        self.input_feed = vars_[-1]


class RNNVIDecoderBase(nn.Module):
    """onmt/Models.py:1014-1174"""

    def __init__(self, rnn_type, bidirectional_encoder, num_layers, hidden_size, attn_type="general",
                 coverage_attn=False, context_gate=None, copy_attn=False, dropout=0.0, word_dropout=0.0,
                 embeddings=None, latent_dim=None, reuse_copy_attn=False):
        super().__init__()
        assert rnn_type == "LSTM"
        assert not coverage_attn and not copy_attn and context_gate is None, \
            "coverage / copy attention / context gates are not reachable with the published flags"
        assert word_dropout == 0.0, "word dropout is 0 in every published run (opts.py default)"
        self.decoder_type = "rnn"
        self.bidirectional_encoder = bidirectional_encoder
        self.num_layers, self.hidden_size = num_layers, hidden_size
        self.embeddings = embeddings
        self.dropout_p = float(dropout)
        self.latent_dim = latent_dim
        self.rnn = self._build_rnn(rnn_type, self._input_size, hidden_size, num_layers, dropout)
        self.context_gate = None
        self._coverage, self._copy = False, False
        self.attn = GlobalAttention(hidden_size, coverage=False, attn_type=attn_type)

    def forward(self, input, context, state, context_lengths=None, **kwargs):
        """input [tgt_len, batch, 1], context [src_len, batch, hidden] ->
        (outputs [tgt_len, batch, hidden], state, attns {"std": [tgt_len, batch, src_len]})"""
        assert isinstance(state, RNNDecoderState)
        assert input.size(1) == context.size(1)
        hidden, outputs, attns, coverage = self._run_forward_pass(
            input, context, state, context_lengths=context_lengths, **kwargs)
        state.update_state(hidden, outputs[-1].unsqueeze(0), None)
        return outputs, state, attns

    def _fix_enc_hidden(self, h):
        if self.bidirectional_encoder:
            h = torch.cat([h[0:h.size(0):2], h[1:h.size(0):2]], 2)
        return h

    def init_decoder_state(self, src, context, enc_hidden):
        if isinstance(enc_hidden, tuple):
            return RNNDecoderState(context, self.hidden_size,
                                   tuple(self._fix_enc_hidden(h) for h in enc_hidden))
        return RNNDecoderState(context, self.hidden_size, self._fix_enc_hidden(enc_hidden))


class NMTVIModel(FlatParamsMixin, nn.Module):
    """onmt/Models.py:737-1011 -- conditional / fixed-prior VI model 1."""

    def __init__(self, encoder, decoder, multigpu=False, **kwargs):
        super().__init__()
        self.multigpu = multigpu
        self.multimodal_model_type = kwargs["multimodal_model_type"]
        assert self.multimodal_model_type in MODEL_TYPES
        self.image_loss_type = kwargs["image_loss_type"]
        self.conditional = kwargs["conditional"]
        self.image_features_type = kwargs.get("image_features_type", "global")
        assert kwargs.get("encoder_inference") is None and not kwargs.get("two_step_image_prediction", False)
        self.encoder_inference = None
        self.image_features_projector = None
        self.two_step_image_prediction = False
        self.encoder, self.decoder = encoder, decoder
        self.encoder.rnn.fires_early_exchange = True
        self.encoder_tgt = kwargs["encoder_tgt"] if self.conditional else None
        if self.encoder_tgt is not None:
            self.encoder_tgt.no_pack_padded_seq = True
            # concurrent source / target encoders (training): 5 clusters of 16 CTAs (8 batch rows each: the one-row-group
            # kernel) beside 8 clusters of 8 CTAs = 144 of the 148 SMs
            self.encoder_tgt.rnn.fires_early_exchange = True
            self.encoder.rnn.cluster_budget = int(os.environ.get("VMMT_ENC_BUDGET", "5")) or None
            self.encoder_tgt.rnn.cluster_budget = int(os.environ.get("VMMT_TGT_BUDGET", "8")) or None
            self.encoder.rnn.cluster_budget_bwd = int(os.environ.get("VMMT_ENC_BUDGET_BWD", "0")) or None
            self.encoder_tgt.rnn.cluster_budget_bwd = int(os.environ.get("VMMT_TGT_BUDGET_BWD", "0")) or None
        self.inf_net_global = kwargs["inf_net_global"]
        self.gen_net_global = kwargs["gen_net_global"]
        self.inf_net_image = kwargs["inf_net_image"]
        assert self.inf_net_global is not None and self.inf_net_image is not None

    def forward(self, src, tgt, lengths, tgt_lengths, img_feats, img_vecs=None, dec_state=None,
                padding_token=None):
        orig_tgt = tgt
        tgt = tgt[:-1]
        dec_gx = None
        ops.stamp(0)                                             # (measurement aid: no-ops unless VMMT_STAMPS=1)
        if self.conditional:
            # target encoder over the transposed ids: recurrence along the batch axis (hazard H1).  It does not depend
            # on the source encoder: the two stacks run side by side on two streams, each on its share of the SMs.  It is
            # issued FIRST: its chain is the longer one (two input projections per layer, 40 steps), and with the source
            # encoder issued first the target encoder finished 84 us later (measured with tools/phase_stamps.py)
            with ops.branch():
                _, tgt_context = self.encoder_tgt(orig_tgt.transpose(0, 1), lengths=None)
                tgt_context = ops.stamped(tgt_context, 2, 12).transpose(0, 1)
        enc_hidden, context = self.encoder(src, lengths)
        context = ops.stamped(context, 1, 11)
        hook = getattr(self, "early_exchange_hook", None)
        if hook is not None and self.training and torch.is_grad_enabled():
            # data parallel (Optim.enable_early_exchange): the first encoder-stack backward node to run fires the early
            # reduce-scatter.  Autograd runs ready nodes in reverse creation order and the encoders are created first, so
            # by then every module that hangs off the loss rather than off the encoder recurrences (generator, image
            # head, prior / posterior networks, decoder, attention) has issued its weight gradients.
            ops.arm_early_exchange(hook)
        if self.conditional:
            ops.join_branch(tgt_context)
        ready = getattr(self, "params_ready_hook", None)
        if ready is not None:
            # data parallel (Optim.enable_early_exchange): the previous update's all-gather of the buffer's tail -- the
            # latent / image networks and the generator, first read below -- runs under this step's encoder phase; wait
            # for it here (an external-event wait node when the step is a captured graph)
            ready()
        if self.training and dec_state is None and hasattr(self.decoder, "input_projection"):
            # decoder input projection emb(tgt) W_ih[:, :E]^T: needs neither the encoders nor z.  It is issued HERE, beside
            # the latent networks (batch-row MLPs that leave most SMs idle), and not beside the encoder recurrences: a
            # GEMM that fills the SMs keeps the recurrences' thread-block clusters from being placed (measured: the
            # target encoder's first layer went from 80 to 183 us with the projection issued at the top of the step)
            with ops.branch(lane=2):
                dec_gx = self.decoder.input_projection(tgt)
        if self.conditional:
            assert isinstance(self.inf_net_global, GlobalFullInferenceNetwork)
            # p(z|x): only the KL needs it in training -> the low-priority lane (its row-block kernels, forward and backward,
            # must not queue ahead of the critical chain's short kernels in the block scheduler)
            with ops.branch(lane=ops.LOW_LANE if self.training else 0):
                pz0, _ = self.gen_net_global(context, lengths)
            if not self.training:
                ops.join_branch(pz0.mean())
            z0, h = self.inf_net_global(context.detach(), lengths, tgt_context, tgt_lengths, img_feats)
            z0_sample = z0.sample() if self.training else pz0.mean().detach()
        else:
            assert isinstance(self.inf_net_global, GlobalInferenceNetwork)
            z0, h = self.inf_net_global(context.detach(), lengths)               # q(z|x)
            z0_sample = z0.sample() if self.training else z0.mean().detach()
            pz0 = Normal(torch.zeros_like(z0.params()[0]), torch.ones_like(z0.params()[0]))
            pz0.is_standard = True
        with ops.branch(lane=ops.LOW_LANE):                                      # p(v|z) beside the decoder: only the loss needs it
            p_v, _ = self.inf_net_image(z0_sample, context, lengths)
        ops.stamp(3)                                             # latent block done on the main chain
        enc_state = self.decoder.init_decoder_state(src, context, enc_hidden)
        extra = {}
        if dec_gx is not None:
            ops.join_branch(dec_gx, lane=2)
            extra["input_projection"] = dec_gx
        out, dec_state, attns = self.decoder(tgt, context, enc_state if dec_state is None else dec_state,
                                             lengths, image_features=None, z_sample=z0_sample, **extra)
        out = ops.stamped(out, 4, 14)
        ops.join_branch(p_v.mean(), *([] if getattr(pz0, "is_standard", False) else pz0.params()), lane=ops.LOW_LANE)
        attns["p_global_image_features"] = [p_v]
        attns["ground_truth_global_image_features"] = [img_feats]
        attns["p_latent"] = [pz0]
        attns["z_latent"] = [z0]
        attns["z0_sample"] = [z0_sample]
        attns["zz"] = [None]
        attns["logdet"] = [None]
        if self.multigpu:
            dec_state, attns = None, None
        return out, attns, dec_state
