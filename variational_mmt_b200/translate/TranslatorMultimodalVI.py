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
import collections

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from .._lib import ACT_NONE, ACT_TANH, fptr, ptr, stream

PAD_WORD, BOS_WORD, EOS_WORD = "<blank>", "<s>", "</s>"


class _State(object):
    """Device buffers of one decode bucket (all at fixed addresses: the step graph refers to them)."""


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
        self.use_graph = True          # replay the decode step from a CUDA graph (one per (sentences, src_len) bucket)
        self.return_attention = True   # copy the per-step attention history back for ret["attention"]
        self.max_buckets = 8
        self._buckets = collections.OrderedDict()
        vocab = fields["tgt"].vocab
        self.pad, self.bos, self.eos = vocab.stoi[PAD_WORD], vocab.stoi[BOS_WORD], vocab.stoi[EOS_WORD]

    # ------------------------------------------------------------------------------------------
    def _buffers(self, B, S, dev):
        """Static device buffers (and the captured step graph) of one (sentences, src_len) bucket."""
        # the captured step graph bakes in raw parameter addresses: they are part of the key (Optim.set_parameters re-homes
        # the flat buffers into the NVLink peer segment, model.to() re-flattens them -- a stale bucket must not be replayed)
        key = (B, S, str(dev), ops.flags(), next(self.model.parameters()).data_ptr())
        st = self._buckets.get(key)
        if st is not None:
            self._buckets.move_to_end(key)
            return st
        while len(self._buckets) >= self.max_buckets:
            self._buckets.popitem(last=False)
        model, dec = self.model, self.model.decoder
        K, H, nl, Lmax = self.beam_size, dec.hidden_size, dec.num_layers, self.max_length
        R = K * B
        E = dec.embeddings.embedding_size
        V = model.generator[0].weight.size(0)
        f32 = dict(device=dev, dtype=torch.float32)
        i32 = dict(device=dev, dtype=torch.int32)
        st = _State()
        st.B, st.S, st.R = B, S, R
        st.ctx = torch.zeros(S, B, H, **f32)            # NOT tiled over the beam: the K hypotheses of a sentence share it
        st.len_b = torch.zeros(B, device=dev, dtype=torch.int64)
        st.zb = torch.zeros(R, 4 * H, **f32)
        st.h, st.c = torch.zeros(nl, R, H, **f32), torch.zeros(nl, R, H, **f32)
        st.h2, st.c2 = torch.zeros_like(st.h), torch.zeros_like(st.c)
        st.scores = torch.zeros(B, K, **f32)
        st.next_ys = torch.zeros(Lmax + 1, K, B, device=dev, dtype=torch.int64)
        st.prev_ks = torch.zeros(Lmax, K, B, **i32)
        st.tok_cur = torch.zeros(K, B, device=dev, dtype=torch.int64)
        st.prev_cur = torch.zeros(K, B, **i32)
        st.fin_score = torch.zeros(B, **f32)
        st.fin_t, st.fin_k, st.n_fin, st.done = (torch.zeros(B, **i32) for _ in range(4))
        st.n_active = torch.zeros(1, **i32)
        st.step = torch.zeros(1, device=dev, dtype=torch.int64)
        st.attn_hist = torch.zeros(Lmax, R, S, **f32)
        st.align = torch.zeros(R, S, **f32)
        st.emb, st.gpre = torch.zeros(R, E, **f32), torch.zeros(R, 4 * H, **f32)
        st.qp, st.cvec, st.out = (torch.zeros(R, H, **f32) for _ in range(3))
        # generator: with the tensor-core GEMM the epilogue keeps per-tile {max, sum exp, top-K} only; the [R,V] log-prob
        # matrix exists only on the exact-fp32 (SIMT) parity path
        st.fused_gen = bool(L.lib.vmmt_generator_topk_supported(fptr(st.out), fptr(model.generator[0].weight), R, H, V,
                                                              ops.flags()))
        if st.fused_gen:
            st.gen_ws_bytes = int(L.lib.vmmt_generator_topk_workspace_bytes(R, V, K))
            st.gen_ws = torch.zeros(st.gen_ws_bytes // 4, **f32)
            st.logp = st.lse = None
        else:
            st.logp, st.lse = torch.zeros(R, V, **f32), torch.zeros(R, **f32)
        st.graph = None
        st.att_host = None
        self._buckets[key] = st
        return st

    def _reset(self, st):
        st.scores.zero_()
        st.next_ys.fill_(self.pad)
        st.next_ys[0, 0] = self.bos
        st.tok_cur.copy_(st.next_ys[0])
        st.prev_ks.zero_()
        st.prev_cur.zero_()
        st.fin_score.zero_()
        for t in (st.fin_t, st.fin_k, st.n_fin, st.done, st.step):
            t.zero_()
        st.n_active.fill_(st.B)

    def _issue_step(self, st):
        """One decode position for all K*B rows; every argument is a fixed device address (the step index is
        device-resident), so the same launch sequence is valid for every step and can be replayed from a graph."""
        model, dec = self.model, self.model.decoder
        K, B, R, S = self.beam_size, st.B, st.R, st.S
        H, nl = dec.hidden_size, dec.num_layers
        E = dec.embeddings.embedding_size
        w, gen = dec.rnn, model.generator[0]
        V = gen.weight.size(0)
        sm = stream()
        L.call("vmmt_embedding_fwd", ptr(st.tok_cur), R, fptr(dec.embeddings.word_lut.weight),
               dec.embeddings.word_lut.weight.shape[0], E, fptr(st.emb), sm)
        x = st.emb
        for l in range(nl):
            w_ih, w_hh = getattr(w, "weight_ih_l%d" % l), getattr(w, "weight_hh_l%d" % l)
            in_dim = E if l == 0 else H
            ops.gemm_dual(x, w_ih[:, :in_dim], st.h[l], w_hh, st.gpre, R, 4 * H, in_dim, H)   # x W_ih^T + h W_hh^T
            L.call("vmmt_lstm_cell_fwd", fptr(st.gpre), fptr(getattr(w, "bias_ih_l%d" % l)),
                   fptr(getattr(w, "bias_hh_l%d" % l)), fptr(st.zb) if l == 0 else None, fptr(st.c[l]),
                   fptr(st.h2[l]), fptr(st.c2[l]), R, H, sm)
            x = st.h2[l]
        # attention, one step (GlobalAttention.py:147-151,169-190)
        if dec.attn.attn_type == "general":
            ops.gemm(x, dec.attn.linear_in.weight, st.qp, R, H, H)
            q = st.qp
        else:
            q = x
        # rows are beam-major (row = k*B + b), i.e. q is [K, B, H]: the K hypotheses of sentence b are K "time steps" over
        # the SAME context column, so the tiled sequence kernel reads ctx[:, b, :] once per sentence, not once per row
        L.call("vmmt_attention_fwd", fptr(q), fptr(st.ctx), ptr(st.len_b), fptr(st.align), fptr(st.cvec), K, B, S, H, sm)
        L.call("vmmt_beam_record", fptr(st.align), fptr(st.attn_hist), ptr(st.step), R * S, sm)
        w_out = dec.attn.linear_out.weight
        ops.gemm_dual(st.cvec, w_out[:, :H], x, w_out[:, H:], st.out, R, H, H, H, act=ACT_TANH)   # tanh(linear_out([c ; q]))
        # generator + Beam.advance for every sentence + DecoderState.beam_update (Beam.py:64-123, Models.py:589-594)
        if st.fused_gen:
            L.call("vmmt_generator_topk", fptr(st.out), fptr(gen.weight), fptr(gen.bias), R, H, V, K, fptr(st.gen_ws),
                   st.gen_ws_bytes, ops.flags(), sm)
            L.call("vmmt_beam_advance_topk", fptr(st.gen_ws), B, K, V, 0, ptr(st.step), ptr(st.tok_cur),
                   ptr(st.prev_cur), self.eos, fptr(st.scores), ptr(st.next_ys), ptr(st.prev_ks), fptr(st.fin_score),
                   ptr(st.fin_t), ptr(st.fin_k), ptr(st.n_fin), ptr(st.done), ptr(st.n_active), sm)
        else:
            L.call("vmmt_generator_logprobs", fptr(st.out), fptr(gen.weight), fptr(gen.bias), R, H, V, fptr(st.logp),
                   fptr(st.lse), ops.flags(), sm)
            L.call("vmmt_beam_advance", fptr(st.logp), B, K, V, 0, ptr(st.step), ptr(st.tok_cur), ptr(st.prev_cur),
                   self.eos, fptr(st.scores), ptr(st.next_ys), ptr(st.prev_ks), fptr(st.fin_score), ptr(st.fin_t),
                   ptr(st.fin_k), ptr(st.n_fin), ptr(st.done), ptr(st.n_active), sm)
        L.call("vmmt_beam_reorder", fptr(st.h2), fptr(st.h), ptr(st.prev_cur), ptr(st.done), nl, K, B, H, sm)
        L.call("vmmt_beam_reorder", fptr(st.c2), fptr(st.c), ptr(st.prev_cur), ptr(st.done), nl, K, B, H, sm)
        L.call("vmmt_counter_add", ptr(st.step), 1, sm)

    def _capture(self, st):
        dev = st.ctx.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                  # lazy one-time work must not happen inside the capture
            self._issue_step(st)
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._issue_step(st)
        st.graph = g

    @torch.no_grad()
    def translate_batch(self, batch, data=None, sent_idx=None):
        # batch-invariant arithmetic: a sentence decoded alone = the same sentence decoded inside a batch (the reference
        # decodes one sentence at a time, translate_mm_vi.py:80-82; batching must not change its output)
        with ops.batch_invariant():
            return self._translate_batch(batch, data, sent_idx)

    def _translate_batch(self, batch, data=None, sent_idx=None):
        model = self.model
        dec = model.decoder
        assert not model.training, "call model.eval() before translating"
        src, src_lengths = batch.src
        if src.dim() == 2:
            src = src.unsqueeze(2)
        dev = next(model.parameters()).device
        src, src_lengths = src.to(dev, non_blocking=True), src_lengths.to(dev, non_blocking=True)
        S, B = src.size(0), src.size(1)
        K = self.beam_size
        H = dec.hidden_size
        st = self._buffers(B, S, dev)
        R = st.R
        if self.use_graph and st.graph is None:       # capture on the zero-filled bucket, before real data goes in
            self._reset(st)
            self._capture(st)
        # (1) encoder + prior mean (TranslatorMultimodalVI.py:125-138)
        enc_states, context = model.encoder(src, src_lengths)
        net = model.gen_net_global if model.conditional else model.inf_net_global
        q0, _ = net(context, src_lengths)
        z = q0.mean()                                                     # [B, Z]
        # (2) beam-major tiling (:146-158) into the bucket's static buffers
        st.ctx.copy_(context)
        st.len_b.copy_(src_lengths)
        h0, c0 = dec._fix_enc_hidden(enc_states[0]), dec._fix_enc_hidden(enc_states[1])
        nl = h0.size(0)
        st.h.view(nl, K, B, H).copy_(h0.unsqueeze(1).expand(nl, K, B, H))
        st.c.view(nl, K, B, H).copy_(c0.unsqueeze(1).expand(nl, K, B, H))
        E = dec.embeddings.embedding_size
        zb = torch.empty(B, 4 * H, device=dev)
        # z W_ih[:, E:]^T once per batch, exact fp32 per row (as the training path: VI_Model1.StdRNNVIModel1Decoder)
        ops.rowlin([ops._rl_prob([z.contiguous()], dec.rnn.weight_ih_l0[:, E:], None, zb)], B, 4 * H, z.size(1))
        st.zb.view(K, B, 4 * H).copy_(zb.unsqueeze(0).expand(K, B, 4 * H))
        self._reset(st)
        Lmax = self.max_length
        steps = 0
        # (3) the step loop (:163-218)
        for i in range(Lmax):
            if i % self.poll_every == 0 and i > 0 and int(st.n_active.item()) == 0:
                break
            if st.graph is not None:
                st.graph.replay()
            else:
                self._issue_step(st)
            steps = i + 1
        # (4) hypotheses (Beam.sort_finished / get_hyp, Beam.py:128-153; _from_beam :228-243): back-pointer walk for
        # all sentences at once (one vectorised pass per step instead of one Python iteration per token)
        ys = st.next_ys[: steps + 1].cpu().numpy()
        pk = st.prev_ks[:steps].cpu().numpy()
        fs, ft, fk, nf = st.fin_score.cpu().numpy(), st.fin_t.cpu().numpy(), st.fin_k.cpu().numpy(), st.n_fin.cpu().numpy()
        sc = st.scores.cpu().numpy()
        lens = src_lengths.cpu().numpy()
        att = None
        if self.return_attention:
            if st.att_host is None:
                st.att_host = torch.empty(st.attn_hist.shape, dtype=torch.float32, pin_memory=True)
            st.att_host[:steps].copy_(st.attn_hist[:steps], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            att = st.att_host[:steps].view(steps, K, B, S).numpy()
        fin = nf >= 1                                    # else sort_finished(minimum=n_best): top of the beam as it stands
        t_end = np.where(fin, ft, steps).astype(np.int64)
        k = np.where(fin, fk, 0).astype(np.int64)
        score = np.where(fin, fs, sc[:, 0])
        rows = np.arange(B)
        toks = np.full((B, max(steps, 1)), self.pad, np.int64)
        best_att = np.zeros((B, max(steps, 1), S), np.float32) if att is not None else None
        for j in range(steps - 1, -1, -1):
            live = j < t_end
            toks[live, j] = ys[j + 1, k[live], rows[live]]
            k = np.where(live, pk[j, k, rows], k)        # attn[j] was re-ordered by prev_ks[j] (Beam.py:107)
            if best_att is not None:
                best_att[live, j] = att[j, k[live], rows[live]]
        ret = {"predictions": [], "scores": [], "attention": []}
        for b in range(B):
            n = int(t_end[b])
            ret["predictions"].append([toks[b, :n].tolist()])
            ret["scores"].append([float(score[b])])
            ret["attention"].append([torch.from_numpy(best_att[b, :n, : int(lens[b])].copy()) if best_att is not None
                                     else torch.zeros(0, int(lens[b]))])
        ret["gold_score"] = [0] * B
        ret["batch"] = batch
        ret["steps"] = steps
        return ret
