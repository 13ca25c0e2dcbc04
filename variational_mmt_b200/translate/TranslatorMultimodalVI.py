"""Batched prior-only beam search (reference: onmt/translate/TranslatorMultimodalVI.py:9-243).

Same constructor and ``translate_batch(batch, data, sent_idx)`` contract as the reference.  Differences in
execution, not in results:
  * the reference decodes ONE sentence per call (translate_mm_vi.py:80-82 forces batch_size 1; batch > 1
    crashes on ``s0.repeat(beam,1,1).squeeze(1)``); here all B sentences x K beams of a batch advance
    together, rows beam-major (row = k*B + b, exactly the reference's ``repeat(1, beam, 1)`` tiling);
  * ``Beam.advance`` / ``beam_update`` (Beam.py:64-123, Models.py:589-594) -- top-K over K*V candidates,
    EOS bookkeeping, back pointers, state reorder -- run in two kernels per step for the whole batch; the
    host only polls a "sentences still active" counter every few steps;
  * the image features the reference loads (TranslatorMultimodalVI.py:75-84) are never used by its decoder
    and are not touched here; z = mean of p(z|x) (conditional) or of q(z|x) (fixed prior) (:129-134).
The decoder step itself (1-step LSTM x 2 layers + attention + generator) calls the libvmmt C ABI directly
(no autograd bookkeeping): GEMMs for the gate pre-activations (M = K*B rows), the fused cell kernel, the
attention core, the log-softmax generator.
"""
import numpy as np
import torch

from .. import _lib as L
from .. import ops
from .._lib import ACT_NONE, ACT_TANH, fptr, ptr, stream

PAD_WORD, BOS_WORD, EOS_WORD = "<blank>", "<s>", "</s>"


class TranslatorMultimodalVI(object):
    def __init__(self, model, fields, beam_size, n_best=1, max_length=100, global_scorer=None, copy_attn=False,
                 cuda=False, beam_trace=False, min_length=0, test_img_feats=None, multimodal_model_type=None):
        assert test_img_feats is not None, "Please provide file with test image features."
        assert multimodal_model_type is not None, "Please provide the multimodal model type name."
        assert multimodal_model_type == "vi-model1", "Multi-modal model not implemented: %s" % multimodal_model_type
        assert not copy_attn, "copy attention is not part of VI model 1"
        assert n_best == 1, "n_best > 1 is not used by the published scripts (device beam keeps the best finished hypothesis)"
        assert min_length == 0, "min_length > 0 is not used by the published scripts"
        assert 1 <= beam_size <= 8, "device beam supports beam sizes 1..8"
        if global_scorer is not None and hasattr(global_scorer, "alpha"):
            assert float(global_scorer.alpha) == 0.0 and float(global_scorer.beta) == 0.0, \
                "only alpha = beta = 0 (summed log-prob, the published decode flags) is implemented"
        self.model, self.fields = model, fields
        self.n_best, self.max_length, self.beam_size = n_best, max_length, beam_size
        self.global_scorer, self.copy_attn, self.cuda, self.min_length = global_scorer, copy_attn, cuda, min_length
        self.test_img_feats, self.multimodal_model_type = test_img_feats, multimodal_model_type
        self.beam_accum = None
        self.poll_every = 4            # steps between host polls of the active-sentence counter
        vocab = fields["tgt"].vocab
        self.pad, self.bos, self.eos = vocab.stoi[PAD_WORD], vocab.stoi[BOS_WORD], vocab.stoi[EOS_WORD]

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def translate_batch(self, batch, data=None, sent_idx=None):
        model = self.model
        dec = model.decoder
        assert not model.training, "call model.eval() before translating"
        src, src_lengths = batch.src
        if src.dim() == 2:
            src = src.unsqueeze(2)
        dev = next(model.parameters()).device
        src, src_lengths = src.to(dev), src_lengths.to(dev)
        S, B = src.size(0), src.size(1)
        K = self.beam_size
        R = K * B
        H = dec.hidden_size
        nl = dec.num_layers
        # (1) encoder + prior mean (TranslatorMultimodalVI.py:125-138)
        enc_states, context = model.encoder(src, src_lengths)
        net = model.gen_net_global if model.conditional else model.inf_net_global
        q0, _ = net(context, src_lengths)
        z = q0.mean()                                                     # [B, Z]
        # (2) beam-major tiling (:146-158)
        ctx_r = context.repeat(1, K, 1).contiguous()                      # [S, K*B, H]
        len_r = src_lengths.repeat(K).contiguous()
        h = dec._fix_enc_hidden(enc_states[0]).repeat(1, K, 1).contiguous()   # [L, K*B, H]
        c = dec._fix_enc_hidden(enc_states[1]).repeat(1, K, 1).contiguous()
        h2, c2 = torch.empty_like(h), torch.empty_like(c)
        E = dec.embeddings.embedding_size
        w = dec.rnn
        zb = torch.empty(B, 4 * H, device=dev)
        ops.gemm(z.contiguous(), w.weight_ih_l0[:, E:], zb, B, 4 * H, z.size(1))   # z W_ih[:, E:]^T once per batch
        zb = zb.repeat(K, 1).contiguous()
        Lmax = self.max_length
        f32 = dict(device=dev, dtype=torch.float32)
        scores = torch.zeros(B, K, **f32)
        next_ys = torch.full((Lmax + 1, K, B), self.pad, device=dev, dtype=torch.int64)
        next_ys[0, 0] = self.bos
        prev_ks = torch.zeros(Lmax, K, B, device=dev, dtype=torch.int32)
        fin_score = torch.zeros(B, **f32)
        fin_t = torch.zeros(B, device=dev, dtype=torch.int32)
        fin_k = torch.zeros(B, device=dev, dtype=torch.int32)
        n_fin = torch.zeros(B, device=dev, dtype=torch.int32)
        done = torch.zeros(B, device=dev, dtype=torch.int32)
        n_active = torch.full((1,), B, device=dev, dtype=torch.int32)
        attn_hist = torch.zeros(Lmax, R, S, **f32)
        emb_w = dec.embeddings.word_lut.weight
        gen = model.generator[0]
        V = gen.weight.size(0)
        emb = torch.empty(R, E, **f32)
        gpre = torch.empty(R, 4 * H, **f32)
        qp = torch.empty(R, H, **f32)
        cvec = torch.empty(R, H, **f32)
        out = torch.empty(R, H, **f32)
        logp = torch.empty(R, V, **f32)
        lse = torch.empty(R, **f32)
        w_in = dec.attn.linear_in.weight if dec.attn.attn_type == "general" else None
        w_out = dec.attn.linear_out.weight
        st = stream()
        steps = 0
        # (3) the step loop (:163-218)
        for i in range(Lmax):
            if i % self.poll_every == 0 and i > 0 and int(n_active.item()) == 0:
                break
            tok = next_ys[i].view(R)
            L.call("vmmt_embedding_fwd", ptr(tok), R, fptr(emb_w), E, fptr(emb), st)
            x = emb
            for l in range(nl):
                w_ih = getattr(w, "weight_ih_l%d" % l)
                w_hh = getattr(w, "weight_hh_l%d" % l)
                in_dim = E if l == 0 else H
                ops.gemm(x, w_ih[:, :in_dim], gpre, R, 4 * H, in_dim)
                ops.gemm(h[l], w_hh, gpre, R, 4 * H, H, accumulate=1)
                L.call("vmmt_lstm_cell_fwd", fptr(gpre), fptr(getattr(w, "bias_ih_l%d" % l)),
                       fptr(getattr(w, "bias_hh_l%d" % l)), fptr(zb) if l == 0 else None, fptr(c[l]),
                       fptr(h2[l]), fptr(c2[l]), R, H, st)
                x = h2[l]
            # attention, one step (GlobalAttention.py:147-151,169-190)
            if w_in is not None:
                ops.gemm(x, w_in, qp, R, H, H)
                q = qp
            else:
                q = x
            align = attn_hist[i]
            L.call("vmmt_attention_fwd", fptr(q), fptr(ctx_r), ptr(len_r), fptr(align), fptr(cvec), 1, R, S, H, st)
            ops.gemm(cvec, w_out[:, :H], out, R, H, H)
            ops.gemm(x, w_out[:, H:], out, R, H, H, act=ACT_TANH, accumulate=2)
            L.call("vmmt_generator_logprobs", fptr(out), fptr(gen.weight), fptr(gen.bias), R, H, V, fptr(logp),
                   fptr(lse), st)
            # Beam.advance for every sentence + DecoderState.beam_update (Beam.py:64-123, Models.py:589-594)
            L.call("vmmt_beam_advance", fptr(logp), B, K, V, i, self.eos, fptr(scores), ptr(next_ys), ptr(prev_ks),
                   fptr(fin_score), ptr(fin_t), ptr(fin_k), ptr(n_fin), ptr(done), ptr(n_active), st)
            L.call("vmmt_beam_reorder", fptr(h2), fptr(h), ptr(prev_ks[i]), ptr(done), nl, K, B, H, st)
            L.call("vmmt_beam_reorder", fptr(c2), fptr(c), ptr(prev_ks[i]), ptr(done), nl, K, B, H, st)
            steps = i + 1
        # (4) hypotheses (Beam.sort_finished / get_hyp, Beam.py:128-153; _from_beam :228-243)
        ys = next_ys[: steps + 1].cpu().numpy()
        pk = prev_ks[:steps].cpu().numpy()
        fs, ft, fk, nf = fin_score.cpu().numpy(), fin_t.cpu().numpy(), fin_k.cpu().numpy(), n_fin.cpu().numpy()
        sc = scores.cpu().numpy()
        dn = done.cpu().numpy()
        lens = src_lengths.cpu().numpy()
        att = attn_hist[:steps].view(steps, K, B, S).cpu()
        # a sentence that was frozen (done) stopped advancing at its own last step
        ret = {"predictions": [], "scores": [], "attention": []}
        for b in range(B):
            if nf[b] >= 1:
                score, t, k = float(fs[b]), int(ft[b]), int(fk[b])
            else:                                        # sort_finished(minimum=n_best): top of the beam as it stands
                score, t, k = float(sc[b, 0]), steps, 0
            hyp, rows = [], []
            for j in range(t - 1, -1, -1):
                hyp.append(int(ys[j + 1, k, b]))
                k = int(pk[j, k, b])                     # attn[j] was re-ordered by prev_ks[j] (Beam.py:107)
                rows.append(att[j, k, b, : int(lens[b])])
            ret["predictions"].append([hyp[::-1]])
            ret["scores"].append([score])
            ret["attention"].append([torch.stack(rows[::-1]) if rows else torch.zeros(0, int(lens[b]))])
        ret["gold_score"] = [0] * B
        ret["batch"] = batch
        ret["steps"] = steps
        return ret
