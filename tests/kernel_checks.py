"""Per-kernel numerical checks (libvmmt C ABI vs plain torch fp32/fp64 math on the same device).
Each check returns a list of (label, error, tolerance).  Used by tests/test_gpu_kernels.py (asserts)
and tools/gpu_diag.py (prints everything, never stops at the first failure)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

DEV = "cuda"


def _r(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_gemm(mode_tol=1e-5):
    from variational_mmt_b200 import ops
    out = []
    for (M, N, K) in [(1240, 2000, 500), (40, 2048, 2048), (77, 130, 52), (1240, 10000, 500), (5, 7, 3),
                      (2000, 500, 1240), (256, 512, 1024)]:
        for ak, bk in [(True, True), (True, False), (False, False), (False, True)]:
            a = _r(M, K, seed=1) if ak else _r(K, M, seed=1)
            b = _r(N, K, seed=2) if bk else _r(K, N, seed=2)
            bias = _r(N, seed=3)
            A = a if ak else a.t()
            B = b.t() if bk else b
            ref = A.double() @ B.double() + bias.double()
            c = torch.empty(M, N, device=DEV)
            ops.gemm(a, b, c, M, N, K, a_kmajor=ak, b_kmajor=bk, bias=bias)
            out.append((f"gemm {M}x{N}x{K} a_k={int(ak)} b_k={int(bk)}", _rel(c, ref), mode_tol))
    # epilogues + accumulate modes + strided operands
    M, N, K = 130, 96, 200
    a, b, bias = _r(M, K, seed=4), _r(N, 2 * K, seed=5), _r(N, seed=6)
    bv = b[:, K:]
    for act, fn in [(1, F.relu), (2, torch.tanh), (3, F.softplus), (4, torch.sigmoid)]:
        c = torch.empty(M, N, device=DEV)
        ops.gemm(a, bv, c, M, N, K, bias=bias, act=act)
        out.append((f"gemm act={act} strided-B", _rel(c, fn(a.double() @ bv.double().t() + bias.double())), mode_tol))
    c0 = _r(M, N, seed=7)
    c = c0.clone()
    ops.gemm(a, bv, c, M, N, K, accumulate=1)
    out.append(("gemm accumulate=1", _rel(c, c0.double() + a.double() @ bv.double().t()), mode_tol))
    c = c0.clone()
    ops.gemm(a, bv, c, M, N, K, act=2, accumulate=2)
    out.append(("gemm accumulate=2 tanh", _rel(c, torch.tanh(c0.double() + a.double() @ bv.double().t())), mode_tol))
    # two operand pairs into one accumulator (x W_ih^T + h W_hh^T of a decode step; linear_out([c ; q])):
    # ragged K tails of both pairs, strided weight slices, skinny and tiny shapes (the tiny one takes the fallback)
    for (M, N, K1, K2, act, fn) in [(1250, 2000, 500, 500, 0, lambda v: v), (1250, 500, 500, 500, 2, torch.tanh),
                                    (77, 130, 100, 68, 1, F.relu), (5, 256, 64, 64, 0, lambda v: v),
                                    (3, 7, 5, 9, 2, torch.tanh)]:
        a1, a2 = _r(M, K1, seed=8), _r(M, K2, seed=9)
        w = _r(N, K1 + K2, seed=10)                         # one [N, K1+K2] weight, used as two column slices
        bias = _r(N, seed=11)
        c = torch.empty(M, N, device=DEV)
        ops.gemm_dual(a1, w[:, :K1], a2, w[:, K1:], c, M, N, K1, K2, bias=bias, act=act)
        ref = fn(torch.cat([a1, a2], 1).double() @ w.double().t() + bias.double())
        out.append((f"gemm_dual {M}x{N}x({K1}+{K2}) act={act}", _rel(c, ref), 2 * mode_tol))   # K1+K2 terms of TF32 rounding
    return out


def _torch_lstm_ref(x, h0, c0, w_ih, w_hh, b_ih, b_hh, lengths, reverse, rowbias):
    from oracle import vi_model1_ref as R
    xx = x.double().cpu()
    if rowbias is not None:
        # fold the per-example bias in as an extra input block
        pass
    T, N, _ = xx.shape
    H = w_hh.shape[1]
    h = h0.double().cpu() if h0 is not None else torch.zeros(N, H, dtype=torch.float64)
    c = c0.double().cpu() if c0 is not None else torch.zeros(N, H, dtype=torch.float64)
    bb = b_ih.double().cpu() + (rowbias.double().cpu() if rowbias is not None else 0)
    outs = [None] * T
    L = lengths.cpu() if lengths is not None else None
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        g = xx[t] @ w_ih.double().cpu().t() + bb + h @ w_hh.double().cpu().t() + b_hh.double().cpu()
        i, f, gg, o = g.chunk(4, 1)
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        if L is not None:
            m = (t < L).double().unsqueeze(1)
            c, h, outs[t] = m * c2 + (1 - m) * c, m * h2 + (1 - m) * h, m * h2
        else:
            c, h, outs[t] = c2, h2, h2
    return torch.stack(outs), h, c


def check_lstm(tol=2e-5):
    from variational_mmt_b200 import ops
    out = []
    cases = [("enc-masked", 30, 40, 500, 500, True, False, False),
             ("dec-h0-rowbias", 31, 40, 500, 500, False, True, True),
             ("tiny-odd", 7, 5, 33, 50, True, True, False),
             ("tgtenc-bi", 40, 32, 500, 250, False, False, False),
             # H > 512: beyond the cluster kernel (16 CTAs x 32 units); the REAL dispatch (no environment override) takes
             # the step-wise GEMM + cell path in tensor-core mode (csrc/lstm_step.cu), masked, with h0/c0, both directions
             ("wide-stepwise", 6, 9, 96, 600, True, True, False),
             ("wide-stepwise-bi", 5, 4, 64, 520, False, False, False),
             # two row groups per cluster (9..16 batch rows: MMA N = 16, two cells per thread)
             ("two-row-groups", 12, 16, 64, 96, True, True, True)]
    for name, T, N, In, H, masked, with_h0, with_rb in cases:
        ndir = 2 if name.endswith("-bi") else 1
        x = _r(T, N, In, scale=0.5, seed=11).requires_grad_(True)
        ws = []
        for d in range(ndir):
            ws += [_r(4 * H, In, scale=0.1, seed=12 + d).requires_grad_(True), _r(4 * H, H, scale=0.1, seed=14 + d).requires_grad_(True),
                   _r(4 * H, scale=0.1, seed=16 + d).requires_grad_(True), _r(4 * H, scale=0.1, seed=18 + d).requires_grad_(True)]
        h0 = _r(ndir, N, H, scale=0.5, seed=20).requires_grad_(True) if with_h0 else None
        c0 = _r(ndir, N, H, scale=0.5, seed=21).requires_grad_(True) if with_h0 else None
        rb = _r(N, 4 * H, scale=0.2, seed=22).requires_grad_(True) if with_rb else None
        lengths = None
        if masked:
            lengths = torch.randint(1, T + 1, (N,), generator=torch.Generator().manual_seed(3)).sort(descending=True)[0]
            lengths[0] = T
            lengths = lengths.to(DEV)
        o, hT, cT = ops.lstm_layer(x, h0, c0, rb, lengths, {"save": True}, ws)
        wo, wh, wc = _r(T, N, ndir * H, seed=30), _r(ndir, N, H, seed=31), _r(ndir, N, H, seed=32)
        for w in ws:
            w.grad = None
        (o * wo).sum().add((hT * wh).sum()).add((cT * wc).sum()).backward()
        got = dict(out=o, hT=hT, cT=cT, dx=x.grad, **{f"dw{i}": w.grad for i, w in enumerate(ws)})
        if with_h0:
            got.update(dh0=h0.grad, dc0=c0.grad)
        if with_rb:
            got.update(drb=rb.grad)
        # reference (fp64 CPU autograd)
        xr = x.detach().double().cpu().requires_grad_(True)
        wr = [w.detach().double().cpu().requires_grad_(True) for w in ws]
        h0r = h0.detach().double().cpu().requires_grad_(True) if with_h0 else None
        c0r = c0.detach().double().cpu().requires_grad_(True) if with_h0 else None
        rbr = rb.detach().double().cpu().requires_grad_(True) if with_rb else None
        outs, hs, cs = [], [], []
        for d in range(ndir):
            oo, hh, cc = _torch_lstm_ref(xr, None if h0r is None else h0r[d], None if c0r is None else c0r[d],
                                         wr[4 * d], wr[4 * d + 1], wr[4 * d + 2], wr[4 * d + 3],
                                         lengths, d == 1, rbr)
            outs.append(oo); hs.append(hh); cs.append(cc)
        oref, href, cref = torch.cat(outs, 2), torch.stack(hs), torch.stack(cs)
        ((oref * wo.double().cpu()).sum() + (href * wh.double().cpu()).sum() + (cref * wc.double().cpu()).sum()).backward()
        ref = dict(out=oref, hT=href, cT=cref, dx=xr.grad, **{f"dw{i}": w.grad for i, w in enumerate(wr)})
        if with_h0:
            ref.update(dh0=h0r.grad, dc0=c0r.grad)
        if with_rb:
            ref.update(drb=rbr.grad)
        for k in got:
            out.append((f"lstm[{name}] {k}", _rel(got[k].detach().cpu(), ref[k].detach()), tol))
    return out


def check_attention(tol=2e-5):
    from variational_mmt_b200 import ops
    from oracle import vi_model1_ref as R
    out = []
    # (.., 80, 1024): 3 source blocks x 2 contraction chunks of the v2 kernel; (17, 4, 50, 512): 2 position groups x 2 query
    # shares in the context gradient; (33, 2, 128, 96): the maximum source length; (5, 3, 33, 130): H % 4 != 0 -> v1 kernels
    for (T, B, S, H) in [(31, 40, 30, 500), (9, 5, 11, 64), (1, 200, 17, 500), (3, 2, 70, 36), (20, 3, 80, 1024),
                         (17, 4, 50, 512), (33, 2, 128, 96), (5, 3, 33, 130), (16, 2, 32, 512), (2, 1, 1, 4)]:
        qp = _r(T, B, H, scale=0.3, seed=41).requires_grad_(True)
        ctx = _r(S, B, H, scale=0.3, seed=42).requires_grad_(True)
        lengths = torch.randint(1, S + 1, (B,), generator=torch.Generator().manual_seed(5))
        lengths[0] = S
        lengths = lengths.to(DEV)
        cvec, align = ops.AttentionCoreFn.apply(qp, ctx, lengths)
        w = _r(T, B, H, seed=43)
        (cvec * w).sum().backward()
        q2, c2 = qp.detach().double().requires_grad_(True), ctx.detach().double().requires_grad_(True)
        sc = torch.einsum("tbh,sbh->tbs", q2, c2)
        mask = torch.arange(S, device=DEV).unsqueeze(0) < lengths.unsqueeze(1)
        sc = sc.masked_fill(~mask.unsqueeze(0), float("-inf"))
        a = sc.softmax(-1)
        cv = torch.einsum("tbs,sbh->tbh", a, c2)
        (cv * w.double()).sum().backward()
        tag = f"attn T{T} B{B} S{S} H{H}"
        out += [(tag + " align", float((align.double() - a).abs().max()), tol),
                (tag + " cvec", _rel(cvec, cv), tol), (tag + " dqp", _rel(qp.grad, q2.grad), tol),
                (tag + " dctx", _rel(ctx.grad, c2.grad), tol)]
    return out


def check_small_ops(tol=2e-5):
    from variational_mmt_b200 import ops, _lib as L
    out = []
    # embedding
    V, E = 300, 52
    w = torch.nn.Parameter(_r(V, E, seed=50))
    idx = torch.randint(0, V, (9, 7), generator=torch.Generator().manual_seed(1)).to(DEV)
    idx[0, 0] = 1
    e = ops.embedding(idx, w, 1)
    g = _r(9, 7, E, seed=51)
    w.grad = None
    (e * g).sum().backward()
    wr = w.detach().clone().requires_grad_(True)
    er = F.embedding(idx, wr, padding_idx=1)
    (er * g).sum().backward()
    out += [("embedding fwd", _rel(e, er), 1e-7), ("embedding bwd", _rel(w.grad, wr.grad), tol)]
    # masked mean
    x = _r(11, 6, 40, seed=52).requires_grad_(True)
    ln = torch.tensor([11, 9, 9, 4, 2, 1], device=DEV)
    m = ops.MaskedMeanFn.apply(x, ln)
    gg = _r(6, 40, seed=53)
    (m * gg).sum().backward()
    xr = x.detach().double().requires_grad_(True)
    mask = (torch.arange(11, device=DEV).unsqueeze(1) < ln.unsqueeze(0)).double().unsqueeze(2)
    mr = (xr * mask).sum(0) / mask.sum(0)
    (mr * gg.double()).sum().backward()
    out += [("masked_mean fwd", _rel(m, mr), tol), ("masked_mean bwd", _rel(x.grad, xr.grad), tol)]
    # linear (+act, col slice, addend)
    W = torch.nn.Parameter(_r(70, 90, scale=0.2, seed=54)); bb = torch.nn.Parameter(_r(70, seed=55))
    x = _r(13, 40, seed=56).requires_grad_(True); add = _r(13, 70, seed=57).requires_grad_(True)
    y = ops.linear(x, W, bb, ops.ACT_SOFTPLUS, cols=(50, 90), addend=add)
    gy = _r(13, 70, seed=58)
    W.grad = None; bb.grad = None
    (y * gy).sum().backward()
    Wr, br, xr, ar = (t.detach().double().requires_grad_(True) for t in (W, bb, x, add))
    yr = F.softplus(xr @ Wr[:, 50:90].t() + br + ar)
    (yr * gy.double()).sum().backward()
    out += [("linear fwd", _rel(y, yr), tol), ("linear dW", _rel(W.grad, Wr.grad), tol),
            ("linear db", _rel(bb.grad, br.grad), tol), ("linear dx", _rel(x.grad, xr.grad), tol),
            ("linear dadd", _rel(add.grad, ar.grad), tol)]
    # gate
    z = _r(8, 30, seed=59); gw = torch.nn.Parameter(_r(1, 30, scale=0.3, seed=60)); gb = torch.nn.Parameter(_r(1, seed=61))
    gated = ops.gate(z, gw, gb)
    gg = _r(8, 30, seed=62)
    gw.grad = None; gb.grad = None
    (gated * gg).sum().backward()
    gwr, gbr = gw.detach().double().requires_grad_(True), gb.detach().double().requires_grad_(True)
    gr = z.double() * torch.sigmoid(z.double() @ gwr.t() + gbr)
    (gr * gg.double()).sum().backward()
    out += [("gate fwd", _rel(gated, gr), tol), ("gate dw", _rel(gw.grad, gwr.grad), tol),
            ("gate db", _rel(gb.grad, gbr.grad), tol)]
    # dropout: keep-rate, scaling, determinism of the regenerated mask
    xx = torch.ones(1 << 20, device=DEV, requires_grad=True)
    ops.manual_seed(7)
    yy = ops.DropoutFn.apply(xx, 0.5)
    yy.sum().backward()
    keep = float((yy > 0).float().mean())
    out += [("dropout keep-rate", abs(keep - 0.5), 5e-3), ("dropout scale", float((yy.max() - 2.0).abs()), 1e-6),
            ("dropout bwd mask", float((xx.grad - yy).abs().max()), 0.0)]
    # sampling statistics
    mu, sd = torch.zeros(1 << 20, device=DEV), torch.ones(1 << 20, device=DEV)
    zz = ops.normal_sample(mu, sd)
    out += [("philox normal mean", abs(float(zz.mean())), 5e-3), ("philox normal var", abs(float(zz.var()) - 1), 1e-2)]
    return out


def check_loss(tol=2e-5):
    from variational_mmt_b200 import ops
    from oracle import vi_model1_ref as R
    out = []
    for (M, H, V, B, Z, D, prior) in [(1240, 500, 10000, 40, 500, 2048, True), (35, 64, 150, 5, 24, 2048, False)]:
        x = _r(M, H, scale=0.5, seed=70).requires_grad_(True)
        W = torch.nn.Parameter(_r(V, H, scale=0.1, seed=71)); b = torch.nn.Parameter(_r(V, scale=0.1, seed=72))
        tg = torch.randint(0, V, (M,), generator=torch.Generator().manual_seed(2)).to(DEV)
        tg[::7] = 1
        mq = _r(B, Z, seed=73).requires_grad_(True); sq = (F.softplus(_r(B, Z, seed=74)) + 0.1).requires_grad_(True)
        mp = _r(B, Z, seed=75).requires_grad_(True) if prior else None
        sp = (F.softplus(_r(B, Z, seed=76)) + 0.1).requires_grad_(True) if prior else None
        loc = _r(B, D, seed=77).requires_grad_(True); v = _r(B, D, seed=78).abs()
        for legacy in (True, False):
            for t in (x, mq, sq, loc) + ((mp, sp) if prior else ()):
                t.grad = None
            W.grad = None; b.grad = None
            cfg = {"pad_idx": 1, "kl_weight": 0.7, "legacy_image_grad": legacy}
            loss, stats = ops.vi_loss(x, tg, W, b, mq, sq, mp, sp, loc, v, cfg)
            (loss / 40.0).backward()
            xr, Wr, br, mqr, sqr, locr = (t.detach().double().requires_grad_(True) for t in (x, W, b, mq, sq, loc))
            mpr = mp.detach().double().requires_grad_(True) if prior else torch.zeros(B, Z, device=DEV, dtype=torch.float64)
            spr = sp.detach().double().requires_grad_(True) if prior else torch.ones(B, Z, device=DEV, dtype=torch.float64)
            lp = F.log_softmax(xr @ Wr.t() + br, 1)
            nz = tg.ne(1)
            nll = -(lp.gather(1, tg.unsqueeze(1)).squeeze(1) * nz.double()).sum()
            kl = R.kl_normal(mqr, sqr, mpr, spr)
            img, cos = R.image_terms(locr, v.double(), legacy_grad=legacy)
            lr = nll - img + 0.7 * kl
            (lr / 40.0).backward()
            tag = f"loss M{M} V{V} legacy={int(legacy)}"
            out += [(tag + " value", abs(float(loss) - float(lr)) / abs(float(lr)), tol),
                    (tag + " nll", abs(float(stats[0]) - float(nll)) / float(nll), tol),
                    (tag + " n_words", abs(float(stats[1]) - float(nz.sum())), 0.0),
                    (tag + " n_correct", abs(float(stats[2]) - float((lp.argmax(1).eq(tg) & nz).sum())), 0.0),
                    (tag + " kl", abs(float(stats[3]) - float(kl)) / float(kl), tol),
                    (tag + " img", abs(float(stats[4]) - float(img)) / abs(float(img)), tol),
                    (tag + " cos", abs(float(stats[5]) - float(cos)), tol),
                    (tag + " dx", _rel(x.grad, xr.grad), tol), (tag + " dW", _rel(W.grad, Wr.grad), tol),
                    (tag + " db", _rel(b.grad, br.grad), tol), (tag + " dmq", _rel(mq.grad, mqr.grad), tol),
                    (tag + " dsq", _rel(sq.grad, sqr.grad), tol), (tag + " dloc", _rel(loc.grad, locr.grad), tol)]
            if prior:
                out += [(tag + " dmp", _rel(mp.grad, mpr.grad), tol), (tag + " dsp", _rel(sp.grad, spr.grad), tol)]
    return out


def check_optim(tol=1e-5):
    import variational_mmt_b200 as vm
    out = []
    n = 100003
    p = torch.nn.Parameter(_r(n, seed=90)); q = torch.nn.Parameter(_r(17, 5, seed=91))
    opt = vm.Optim("adam", 0.002, 5)
    opt.set_parameters([p, q])
    pr, qr = p.detach().clone().requires_grad_(True), q.detach().clone().requires_grad_(True)
    topt = torch.optim.Adam([pr, qr], lr=0.002, betas=(0.9, 0.999), eps=1e-9)
    for step in range(3):
        g1, g2 = _r(n, scale=0.05 * (step + 1), seed=92 + step), _r(17, 5, seed=95 + step)
        p.grad.copy_(g1); q.grad.copy_(g2)
        pr.grad, qr.grad = g1.clone(), g2.clone()
        tn = torch.nn.utils.clip_grad_norm_([pr, qr], 5.0)
        out.append((f"optim grad_norm step{step}", abs(float(opt.grad_norm()) - float(tn)) / float(tn), tol))
        opt.step(); topt.step()
        out.append((f"optim adam p step{step}", _rel(p.detach() - _r(n, seed=90), pr.detach() - _r(n, seed=90)), 1e-4))
        out.append((f"optim adam q step{step}", _rel(q.detach(), qr.detach()), tol))
    return out


def check_rowmlp(tol=1e-5):
    """ops.row_mlp (csrc/rowlin.cu, exact fp32, cluster split-K) against fp64 torch: the prior / posterior / image-head
    shapes of config 1, a segmented input with gradients to a subset of the segments, ragged dims, M > one row block."""
    from variational_mmt_b200 import ops
    out = []
    exact = 5e-7          # fp32 FMA chains of <= 3048 terms: far below TF32's 1e-3 in either GEMM mode
    cases = [("posterior", 40, [500, 500, 2048], 500, 500, 2, [False, True, False]),
             ("prior", 40, [500], 500, 500, 2, [True]),
             ("image-head", 40, [500], 2048, 2048, 1, [True]),
             ("ragged", 7, [33, 50, 21], 45, 19, 2, [True, False, True]),
             ("tiny", 5, [64], 24, 24, 2, [False]),
             ("many-rows", 250, [500], 500, 500, 2, [False])]
    for name, M, ks, N1, N2, nh, need in cases:
        K = sum(ks)
        xs = [(_r(M, k, scale=0.5, seed=100 + i)).requires_grad_(nd) for i, (k, nd) in enumerate(zip(ks, need))]
        heads = []
        for hd in range(nh):
            heads.append([torch.nn.Parameter(_r(N1, K, scale=0.05, seed=110 + hd)), torch.nn.Parameter(_r(N1, scale=0.1, seed=112 + hd)),
                          torch.nn.Parameter(_r(N2, N1, scale=0.05, seed=114 + hd)), torch.nn.Parameter(_r(N2, scale=0.1, seed=116 + hd))])
        acts = (ops.ACT_NONE, ops.ACT_SOFTPLUS)[:nh]
        ys = ops.row_mlp(xs, heads, acts)
        wy = [_r(M, N2, seed=120 + hd) for hd in range(nh)]
        sum((y * w).sum() for y, w in zip(ys, wy)).backward()
        ops.join_side()
        torch.cuda.synchronize()
        xr = [x.detach().double().requires_grad_(nd) for x, nd in zip(xs, need)]
        hr = [[t.detach().double().requires_grad_(True) for t in hd] for hd in heads]
        xc = torch.cat(xr, 1)
        yr = []
        for hd, (W1, b1, W2, b2) in enumerate(hr):
            y = F.relu(xc @ W1.t() + b1) @ W2.t() + b2
            yr.append(F.softplus(y) if hd == 1 else y)
        sum((y * w.double()).sum() for y, w in zip(yr, wy)).backward()
        for hd in range(nh):
            out.append((f"rowmlp {name} y{hd}", _rel(ys[hd], yr[hd]), exact))
            for j, nm in enumerate(("dW1", "db1", "dW2", "db2")):
                out.append((f"rowmlp {name} head{hd} {nm}", _rel(heads[hd][j].grad, hr[hd][j].grad), tol))
        for i, nd in enumerate(need):
            if nd:
                out.append((f"rowmlp {name} dx{i}", _rel(xs[i].grad, xr[i].grad), exact))
            else:
                out.append((f"rowmlp {name} dx{i} absent", 0.0 if xs[i].grad is None else 1.0, 0.0))
    # batch invariance: a row computed alone equals the same row computed inside a batch, bit for bit
    x = _r(250, 500, scale=0.5, seed=130)
    heads = [[_r(500, 500, scale=0.05, seed=131), _r(500, seed=132), _r(500, 500, scale=0.05, seed=133), _r(500, seed=134)]]
    with torch.no_grad():
        (yb,) = ops.row_mlp([x], heads, (ops.ACT_NONE,))
        (y1,) = ops.row_mlp([x[77:78].contiguous()], heads, (ops.ACT_NONE,))
    out.append(("rowmlp batch invariance (row alone == row in batch)", float((yb[77] - y1[0]).abs().max()), 0.0))
    return out


def check_gemm_bf16(tol=6e-3):
    """bf16 variant (ops.set_gemm_mode(2)): vmmt_cast_bf16 + vmmt_gemm_bf16 (tcgen05.mma.kind::f16 on bf16 operands, fp32
    accumulate) in all four operand-major combinations (forward, dgrad, wgrad forms), epilogues and accumulate modes,
    against fp64 products of the bf16-ROUNDED operands (tight: only accumulation order differs) and of the unrounded
    ones (the bf16 operand-rounding budget)."""
    from variational_mmt_b200 import ops
    out = []
    assert ops.get_gemm_mode() == 2
    rb = lambda t: t.to(torch.bfloat16).double()
    for (M, N, K) in [(1240, 2000, 500), (256, 512, 1024), (77, 130, 52), (2000, 500, 1240), (1240, 10000, 500)]:
        for ak, bk in [(True, True), (True, False), (False, False), (False, True)]:
            a = _r(M, K, seed=1) if ak else _r(K, M, seed=1)
            b = _r(N, K, seed=2) if bk else _r(K, N, seed=2)
            bias = _r(N, seed=3)
            A = a if ak else a.t()
            B = b.t() if bk else b
            c = torch.empty(M, N, device=DEV)
            ops.gemm(a, b, c, M, N, K, a_kmajor=ak, b_kmajor=bk, bias=bias)
            out.append((f"gemm_bf16 {M}x{N}x{K} a_k={int(ak)} b_k={int(bk)} vs rounded operands",
                        _rel(c, rb(A) @ rb(B) + bias.double()), 2e-5))
            out.append((f"gemm_bf16 {M}x{N}x{K} a_k={int(ak)} b_k={int(bk)} vs fp64", _rel(c, A.double() @ B.double() + bias.double()), tol))
    M, N, K = 130, 96, 200
    a, b, bias = _r(M, K, seed=4), _r(N, 2 * K, seed=5), _r(N, seed=6)
    bv = b[:, K:]
    c0 = _r(M, N, seed=7)
    c = c0.clone()
    ops.gemm(a, bv, c, M, N, K, act=2, accumulate=2)
    out.append(("gemm_bf16 accumulate=2 tanh strided-B", _rel(c, torch.tanh(c0.double() + rb(a) @ rb(bv).t())), 2e-5))
    c = c0.clone()
    ops.gemm(a, bv, c, M, N, K, accumulate=1)
    out.append(("gemm_bf16 accumulate=1", _rel(c, c0.double() + rb(a) @ rb(bv).t()), 2e-5))
    a1, a2, w = _r(1250, 500, seed=8), _r(1250, 500, seed=9), _r(500, 1000, seed=10)
    c = torch.empty(1250, 500, device=DEV)
    ops.gemm_dual(a1, w[:, :500], a2, w[:, 500:], c, 1250, 500, 500, 500, act=2)
    out.append(("gemm_dual bf16 tanh", _rel(c, torch.tanh(rb(torch.cat([a1, a2], 1)) @ rb(w).t())), 2e-5))
    return out


ALL = [("gemm", check_gemm), ("lstm", check_lstm), ("attention", check_attention), ("small_ops", check_small_ops),
       ("loss", check_loss), ("optim", check_optim), ("rowmlp", check_rowmlp)]
