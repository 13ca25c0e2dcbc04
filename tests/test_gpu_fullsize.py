"""-m gpu: size-independent properties at BASELINE.json's FULL sizes (the CPU oracle cannot finish these in seconds).

* generator + NLL at the cfg5 stress shape (M = 512 x 79 rows, H = 1024, V = 32000): the fused tensor-core path (logits
  never leave TMEM / registers) against torch's fp32 log_softmax on the same device, loss / n_words / n_correct, and the
  backward's dX against autograd -- tolerance 1e-3 relative (north_star; TF32 operands).
* beam decode at the cfg1 shape with the test-2016-sized batch (250 sentences x beam 5 = 1250 rows, V = 10000): the
  batched decode must return, for a sample of sentences, exactly the tokens of the one-sentence-at-a-time decode the
  reference performs (translate_mm_vi.py:80-82), with scores within 1e-3.
* ragged last batch / single-sentence batch / sentence of length 1 go through the same path.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_generator_nll_cfg5_shape_matches_torch(cuda_device):
    from variational_mmt_b200 import _lib as L
    from variational_mmt_b200._lib import fptr, ptr, stream
    dev = cuda_device
    M, H, V, pad = 512 * 79, 1024, 32000, 1
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, H, device=dev, generator=g) * 0.5
    W = (torch.rand(V, H, device=dev, generator=g) - 0.5) * 0.2
    b = (torch.rand(V, device=dev, generator=g) - 0.5) * 0.2
    tgt = torch.randint(4, V, (M,), device=dev, generator=g)
    tgt[::7] = pad                                                    # ignored positions
    lse = torch.empty(M, device=dev)
    stats = torch.zeros(3, device=dev)
    wsb = int(L.lib.vmmt_generator_workspace_bytes(M, H, V))
    ws = torch.empty(wsb // 4, device=dev)
    L.call("vmmt_generator_nll_fwd", fptr(x), fptr(W), fptr(b), ptr(tgt), pad, M, H, V, fptr(lse), fptr(stats), fptr(ws),
           wsb, 0, stream())
    dx = torch.empty(M, H, device=dev)
    gs = torch.ones(1, device=dev)
    L.call("vmmt_generator_nll_bwd", fptr(x), fptr(W), fptr(b), ptr(tgt), pad, fptr(lse), fptr(gs), 1.0, M, H, V,
           fptr(dx), None, None, fptr(ws), wsb, 0, stream())
    torch.cuda.synchronize()
    # reference in chunks of rows (the [M,V] fp32 log-prob matrix is 5.2 GB; keep the test's footprint small)
    nll, correct, words = 0.0, 0, 0
    lse_ref = torch.empty(M, device=dev)
    dx_ref = torch.empty(M, H, device=dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for r0 in range(0, M, 4096):
            xs = x[r0:r0 + 4096].clone().requires_grad_(True)
            t = tgt[r0:r0 + 4096]
            logits = xs @ W.t() + b
            lp = torch.log_softmax(logits, dim=1)
            lse_ref[r0:r0 + 4096] = torch.logsumexp(logits.detach(), dim=1)
            on = t != pad
            loss = -(lp.gather(1, t.unsqueeze(1)).squeeze(1) * on).sum()
            loss.backward()
            dx_ref[r0:r0 + 4096] = xs.grad
            nll += float(loss)
            words += int(on.sum())
            correct += int(((lp.argmax(1) == t) & on).sum())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert float(stats[1]) == words
    assert abs(float(stats[0]) - nll) <= 1e-3 * abs(nll)
    assert abs(float(stats[2]) - correct) <= max(2, 0.02 * max(correct, 1))        # near-tie argmax may flip under TF32
    assert float((lse - lse_ref).abs().max()) <= 1e-3 * float(lse_ref.abs().max())
    rel = float((dx - dx_ref).norm() / dx_ref.norm())
    assert rel < 2e-3, rel


def test_batched_decode_cfg1_shape_equals_sentence_by_sentence(cuda_device):
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import synthetic
    opt = synthetic.make_opt(conditional=True, dropout=0.5)
    fields = synthetic.make_fields(10000, 10000)
    torch.manual_seed(3435)
    model = vm.make_vi_model_mmt(opt, fields, gpu=True)
    model.eval()

    def translator():
        return vm.TranslatorMultimodalVI(model, fields, beam_size=5, n_best=1, max_length=30,
                                         global_scorer=vm.GNMTGlobalScorer(0., -0.), cuda=True,
                                         test_img_feats=np.zeros((1, 2048), np.float32), multimodal_model_type="vi-model1")

    class B:
        pass
    src, sl, _t, _tl, _img = synthetic.random_batch(10000, 10000, 250, 8, seed=77)
    src[:, -1] = synthetic.PAD                       # a sentence of length 1 (lengths are sorted: the last is the shortest)
    sl[-1] = 1
    src[0, -1] = 17
    b = B(); b.batch_size = 250; b.src = (src.cuda(), sl.cuda())
    tr = translator()
    tr.return_attention = False
    full = tr.translate_batch(b, None, None)
    one = translator()
    one.return_attention = False
    for i in [0, 1, 57, 123, 248, 249]:
        n = int(sl[i])
        s = B(); s.batch_size = 1; s.src = (src[:n, i:i + 1].cuda(), sl[i:i + 1].cuda())
        r = one.translate_batch(s, None, i)
        assert r["predictions"][0][0] == full["predictions"][i][0], f"sentence {i}"
        assert abs(r["scores"][0][0] - full["scores"][i][0]) <= 1e-3 * max(1.0, abs(r["scores"][0][0]))
    # ragged tail: a batch of 3 sentences reuses nothing from the 250-sentence bucket
    t = B(); t.batch_size = 3; t.src = (src[:int(sl[100]), 100:103].cuda(), sl[100:103].cuda())
    r3 = tr.translate_batch(t, None, None)
    for j in range(3):
        assert r3["predictions"][j][0] == full["predictions"][100 + j][0]


def test_cfg5_dimension_train_step_matches_oracle(cuda_device):
    """One whole training step at the HIDDEN SIZES of BASELINE configs[4] (E = H = Z = 1024: the step-wise recurrence
    path for H > 512, 128-wide row-MLP tiles, K = 4096 posterior) on a reduced batch and vocabulary so that the CPU
    oracle finishes in seconds; benchmarked arithmetic (TF32 tensor cores), tolerances of test_gpu_parity.py."""
    import variational_mmt_b200 as vm
    from gpu_helpers import build_cuda_model, to_device, named_grads, relerr, attn_max_rel
    from oracle import synth
    from oracle import vi_model1_ref as R
    cfg = synth.ModelConfig(v_src=2000, v_tgt=2000, emb=1024, hidden=1024, z_dim=1024)
    params = synth.make_params(cfg, 3435, 0.05)
    batch = synth.make_batch(cfg, batch_size=6, seed=9, t_force=14, s_force=12)
    model, fields = build_cuda_model(cfg, params)
    model.train()
    b = to_device(batch)
    loss = vm.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
    model.zero_grad()
    with vm.Normal.inject_noise(b.eps):
        out, attns, _ = model(b.src, b.tgt_in, b.src_lengths, b.tgt_lengths, b.img_feats)
    st = loss.sharded_compute_loss(b, out, attns, 0, b.tgt.size(0), 32, b.batch_size)
    torch.cuda.synchronize()
    ograds, ostats, ofwd = R.train_step_grads(params, cfg, batch)
    assert st.n_words == ostats["n_words"]
    assert st.nmt_loss == pytest.approx(ostats["nmt"], rel=1e-3)
    assert st.td_kl_before == pytest.approx(ostats["td_kl_before"], rel=1e-3)
    assert st.image_feats_loss == pytest.approx(ostats["img_feats_loss"], rel=1e-3)
    assert attn_max_rel(attns["std"].detach().cpu().numpy(), ofwd["attn"].detach().numpy(), batch.src_lengths) <= 1e-3
    grads = named_grads(model)
    for k, og in ograds.items():
        if og is None or not np.any(og.numpy()):
            continue
        floor = 1e-4 * float(np.abs(og.numpy()).max())
        err = np.linalg.norm(grads[k].astype(np.float64) - og.numpy()) / max(np.linalg.norm(og.numpy()), floor)
        assert err <= 3e-2, f"grad {k}: rel err {err:.3e}"
