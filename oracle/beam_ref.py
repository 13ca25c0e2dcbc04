"""CPU restatement of prior-only beam decode (one sentence at a time, as the reference does).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows onmt/translate/TranslatorMultimodalVI.py:59-243 (translate_batch/_from_beam) and
onmt/translate/Beam.py:5-183 with the run-script defaults alpha = beta = 0 (opts.py:425-429), for
which GNMTGlobalScorer.score reduces to the raw summed log-prob (hazard H9: the -0.0 * log(0) NaN
corner is not reproduced; it cannot occur while every accumulated attention entry is > 0).
"""
import torch

from . import vi_model1_ref as R

PAD, BOS, EOS = 1, 2, 3


def decoder_step(p, cfg, tok, z, ctx, lengths, h, c):
    """One decoder position for ``n`` hypotheses: tok [n], z [n,Z], ctx [S,n,H], h/c [L,n,H]
    (VI_Model1.py:94-132 with input length 1)."""
    e = p["decoder.embeddings.make_embedding.emb_luts.0.weight"][tok].unsqueeze(0)
    u = torch.cat([e, z.unsqueeze(0)], 2)
    q, h2, c2 = R.lstm_stack(u, p, "decoder.rnn", cfg.layers, h0=h, c0=c)
    out, align = R.global_attention(q, ctx, lengths, p["decoder.attn.linear_in.weight"],
                                    p["decoder.attn.linear_out.weight"])
    return out[0], align[0], h2, c2


def encode_for_decode(p, cfg, src, length):
    """encoder + z = mean of p(z|x) (conditional) or of q(z|x) (fixed prior)
    (TranslatorMultimodalVI.py:125-138)."""
    x = p["encoder.embeddings.make_embedding.emb_luts.0.weight"][src]
    ctx, h, c = R.lstm_stack(x, p, "encoder.rnn", cfg.layers, lengths=length)
    hx = R.masked_mean(ctx, length)
    net = "gen_net_global" if cfg.conditional else "inf_net_global"
    z = R.mlp2(p, net + ".location", hx)
    return ctx, h, c, z


def beam_search_one(p, cfg, src, beam_size=5, max_length=100, n_best=1):
    """src: int64 [S] (one sentence, no padding).  Returns (tokens, score, attention [len,S])."""
    with torch.no_grad():
        S = src.shape[0]
        length = torch.tensor([S])
        ctx, h, c, z = encode_for_decode(p, cfg, src.view(S, 1), length)
        K = beam_size
        ctx = ctx.repeat(1, K, 1); h = h.repeat(1, K, 1); c = c.repeat(1, K, 1)
        z = z.repeat(K, 1); lens = length.repeat(K)
        scores = torch.zeros(K, dtype=ctx.dtype)
        next_ys = [torch.full((K,), PAD, dtype=torch.long)]
        next_ys[0][0] = BOS
        prev_ks, attns, finished = [], [], []
        eos_top = False
        for _ in range(max_length):
            if eos_top and len(finished) >= n_best:            # Beam.done()
                break
            out, align, h, c = decoder_step(p, cfg, next_ys[-1], z, ctx, lens, h, c)
            lp = R.log_probs(p, out)                             # [K,V]
            V = lp.shape[1]
            if prev_ks:                                          # Beam.advance, Beam.py:83-94
                bs = lp + scores.unsqueeze(1)
                bs[next_ys[-1].eq(EOS)] = -1e20
            else:
                bs = lp[0]
            best, idx = bs.reshape(-1).topk(K, 0, True, True)
            scores = best
            pk = idx // V
            prev_ks.append(pk)
            next_ys.append(idx - pk * V)
            attns.append(align.index_select(0, pk))
            for i in range(K):
                if int(next_ys[-1][i]) == EOS:
                    finished.append((float(scores[i]), len(next_ys) - 1, i))
            if int(next_ys[-1][0]) == EOS:
                eos_top = True
            h = h.index_select(1, pk); c = c.index_select(1, pk)   # DecoderState.beam_update
        if len(finished) < n_best:                               # sort_finished(minimum=n_best)
            finished.append((float(scores[0]), len(next_ys) - 1, 0))
        finished.sort(key=lambda a: -a[0])
        sc, t, k = finished[0]
        hyp, att = [], []
        for j in range(t - 1, -1, -1):                           # Beam.get_hyp
            hyp.append(int(next_ys[j + 1][k]))
            att.append(attns[j][k])
            k = int(prev_ks[j][k])
        return hyp[::-1], sc, torch.stack(att[::-1])
