"""-m gpu: batched device-side beam search against the fixtures of the EXECUTED reference
(tests/golden/*beam5*, *greedy*: TranslatorMultimodalVI.translate_batch run sentence by sentence) and
against the CPU oracle (oracle/beam_ref.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_helpers import build_cuda_model
from oracle import synth

pytestmark = pytest.mark.gpu


class _B:
    pass


def _translator(name, gemm_mode):
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import _lib, ops
    ops.set_gemm_mode(gemm_mode)
    meta, arr = load_golden(name)
    cfg = synth.ModelConfig(**meta["cfg"])
    params = synth.make_params(cfg, meta["param_seed"], meta["param_scale"])
    model, fields = build_cuda_model(cfg, params)
    model.eval()
    ex = meta["extra"]
    tr = vm.TranslatorMultimodalVI(model, fields, beam_size=ex["beam"], n_best=1, max_length=ex["max_length"],
                                   global_scorer=vm.GNMTGlobalScorer(0., -0.), copy_attn=False, cuda=True,
                                   test_img_feats=np.zeros((ex["n_sent"], cfg.img_dim), np.float32),
                                   multimodal_model_type="vi-model1")
    batch = synth.make_batch(cfg, **meta["batch"])
    return tr, batch, arr, ex


def _check(ret, j, arr, i, tol):
    assert ret["predictions"][j][0] == arr[f"tokens/{i}"].tolist(), f"sentence {i}"
    assert ret["scores"][j][0] == pytest.approx(float(arr[f"score/{i}"]), rel=tol, abs=tol)
    a = ret["attention"][j][0].numpy()
    ref = arr[f"attn/{i}"]
    assert a.shape == ref.shape
    # RELATIVE error over every entry (a single sentence's attention has no masked entries).  The decode fixtures use
    # weights of 5x the reference's init scale (param_scale 0.5, so that beam / greedy tokens have real margins): gate
    # pre-activations are 5x larger and the TF32 operand rounding grows along the decoded sentence to 1.4e-3 .. 5.1e-3
    # here (measured, the recurrent state is saturated at this scale; 1e-3 is held at the reference's own init scale by
    # tests/test_gpu_parity.py).  Tokens and scores above are exact / 2e-3.
    assert (np.abs(a - ref) / ref).max() <= (2e-5 if tol <= 1e-4 else 1e-2)


@pytest.mark.parametrize("mode", [1, 0], ids=["fp32_simt", "tf32_tc"])
@pytest.mark.parametrize("name", ["tiny_cond_beam5", "tiny_cond_greedy", "tiny_fixed_beam5"])
def test_sentence_by_sentence_matches_reference(name, mode, cuda_device):
    """batch_size 1, as translate_mm_vi.py runs it."""
    from variational_mmt_b200 import _lib, ops
    try:
        tr, batch, arr, ex = _translator(name, mode)
        for i in range(ex["n_sent"]):
            n = int(batch.src_lengths[i])
            b = _B()
            b.batch_size = 1
            b.src = (torch.as_tensor(batch.src[:n, i:i + 1]), torch.as_tensor([n]))
            ret = tr.translate_batch(b, None, i)
            _check(ret, 0, arr, i, 1e-4 if mode == 1 else 2e-3)
    finally:
        ops.set_gemm_mode(0)


@pytest.mark.parametrize("mode", [1, 0], ids=["fp32_simt", "tf32_tc"])
@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("name", ["tiny_cond_beam5", "tiny_cond_greedy", "tiny_fixed_beam5"])
def test_batched_decode_equals_sentence_by_sentence(name, graph, mode, cuda_device):
    """All sentences of the (padded, length-sorted) batch advance together: new behaviour whose oracle is
    the reference run one sentence at a time."""
    from variational_mmt_b200 import _lib, ops
    try:
        tr, batch, arr, ex = _translator(name, mode)       # mode 0: fused top-K generator epilogue + tiled beam attention
        tr.poll_every = 1
        tr.use_graph = graph
        b = _B()
        b.batch_size = ex["n_sent"]
        b.src = (torch.as_tensor(batch.src), torch.as_tensor(batch.src_lengths))
        for rep in range(2):                      # second call reuses the bucket's buffers / graph
            ret = tr.translate_batch(b, None, list(range(ex["n_sent"])))
            for i in range(ex["n_sent"]):
                _check(ret, i, arr, i, 1e-4 if mode == 1 else 2e-3)
    finally:
        ops.set_gemm_mode(0)


def test_in_process_validation_translate_equals_sentence_by_sentence(cuda_device, tmp_path):
    """translate_dataset (batched, in place on the live training model) returns, in corpus order, what
    the per-sentence decode of translate_mm_vi.py returns, and restores train mode."""
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import _lib, ops, io
    ops.set_gemm_mode(1)
    try:
        cfg = synth.TINY
        params = synth.make_params(cfg, 3435, 0.1)
        model, fields = build_cuda_model(cfg, params)
        model.train()
        ds = io.TripletDataset.synthetic(23, v_src=cfg.v_src, v_tgt=cfg.v_tgt, seed=4, src_max=12, tgt_max=10)
        out = tmp_path / "hyp.txt"
        hyps = vm.translate.translate_dataset(model, fields, ds, batch_size=7, beam_size=1, max_length=20, output=str(out))
        assert model.training
        model.eval()
        tr = vm.TranslatorMultimodalVI(model, fields, beam_size=1, n_best=1, max_length=20,
                                       global_scorer=vm.GNMTGlobalScorer(0., -0.), cuda=True,
                                       test_img_feats=np.zeros((1, 1), np.float32), multimodal_model_type="vi-model1")
        itos = fields["tgt"].vocab.itos
        for i in range(len(ds)):
            b = _B()
            b.batch_size = 1
            s = ds.src_flat[ds.src_off[i]: ds.src_off[i + 1]]
            b.src = (torch.as_tensor(s).view(-1, 1), torch.as_tensor([len(s)]))
            words = [itos[t] for t in tr.translate_batch(b, None, i)["predictions"][0][0]]
            if words and words[-1] == "</s>":
                words = words[:-1]
            assert hyps[i] == words, f"sentence {i}"
        assert out.read_text().splitlines() == [" ".join(h) for h in hyps]
    finally:
        ops.set_gemm_mode(0)


@pytest.mark.parametrize("step,V", [(0, 1000), (3, 1000), (3, 130), (7, 10000)])
def test_fused_topk_generator_equals_materialised_logprobs(step, V, cuda_device):
    """vmmt_generator_topk + vmmt_beam_advance_topk (per-tile {max, sum exp, top-K} kept by the GEMM epilogue) select the
    same hypotheses as vmmt_generator_logprobs + vmmt_beam_advance on the materialised [K*B, V] log-probs -- including
    finished (EOS) beams, the first step (beam row 0 only) and a ragged last tile."""
    from variational_mmt_b200 import _lib as L, ops
    from variational_mmt_b200._lib import fptr, ptr, stream
    dev = cuda_device
    g = torch.Generator(device="cuda").manual_seed(step * 31 + V)
    B, K, H, Lmax, eos = 37, 5, 128, 12, 3
    R = K * B
    x = torch.randn(R, H, device=dev, generator=g)
    W = (torch.rand(V, H, device=dev, generator=g) - 0.5) * 0.4
    bias = (torch.rand(V, device=dev, generator=g) - 0.5) * 0.4
    assert L.lib.vmmt_generator_topk_supported(fptr(x), fptr(W), R, H, V, 0)

    def state():
        st = {}
        st["scores"] = -torch.rand(B, K, device=dev, generator=torch.Generator(device="cuda").manual_seed(5)) * 3
        ys = torch.randint(4, V, (Lmax + 1, K, B), device=dev, generator=torch.Generator(device="cuda").manual_seed(6))
        if step > 0:
            ys[step, 1, ::3] = eos                       # some finished beams
            ys[step, :, 5] = eos                         # a sentence whose beams are all finished
        st["next_ys"] = ys
        st["prev_ks"] = torch.zeros(Lmax, K, B, device=dev, dtype=torch.int32)
        st["tok_cur"] = torch.zeros(K, B, device=dev, dtype=torch.int64)
        st["prev_cur"] = torch.zeros(K, B, device=dev, dtype=torch.int32)
        st["fin_score"] = torch.zeros(B, device=dev)
        for k in ("fin_t", "fin_k", "n_fin", "done"):
            st[k] = torch.zeros(B, device=dev, dtype=torch.int32)
        st["done"][7] = 1                                # a frozen sentence
        st["n_active"] = torch.full((1,), B - 1, device=dev, dtype=torch.int32)
        return st

    def tail(st):
        return (step, None, ptr(st["tok_cur"]), ptr(st["prev_cur"]), eos, fptr(st["scores"]), ptr(st["next_ys"]),
                ptr(st["prev_ks"]), fptr(st["fin_score"]), ptr(st["fin_t"]), ptr(st["fin_k"]), ptr(st["n_fin"]),
                ptr(st["done"]), ptr(st["n_active"]), stream())
    a, f = state(), state()
    logp, lse = torch.empty(R, V, device=dev), torch.empty(R, device=dev)
    L.call("vmmt_generator_logprobs", fptr(x), fptr(W), fptr(bias), R, H, V, fptr(logp), fptr(lse), 0, stream())
    L.call("vmmt_beam_advance", fptr(logp), B, K, V, *tail(a))
    wsb = int(L.lib.vmmt_generator_topk_workspace_bytes(R, V, K))
    ws = torch.zeros(wsb // 4, device=dev)
    L.call("vmmt_generator_topk", fptr(x), fptr(W), fptr(bias), R, H, V, K, fptr(ws), wsb, 0, stream())
    L.call("vmmt_beam_advance_topk", fptr(ws), B, K, V, *tail(f))
    torch.cuda.synchronize()
    for k in ("next_ys", "prev_ks", "tok_cur", "prev_cur", "fin_t", "fin_k", "n_fin", "done", "n_active"):
        assert torch.equal(a[k], f[k]), k
    assert torch.allclose(a["scores"], f["scores"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(a["fin_score"], f["fin_score"], rtol=1e-5, atol=1e-5)
    # and the materialised path is what torch computes
    ref = torch.log_softmax(x.double() @ W.double().t() + bias.double(), dim=1)
    assert float((logp.double() - ref).abs().max()) < 5e-3          # TF32 operands


def test_encoder_and_prior_are_batch_invariant_bitwise(cuda_device):
    """A sentence alone and the same sentence inside a 250-sentence batch give BITWISE identical encoder context, final
    state and prior mean (ops.batch_invariant: no split-K; the recurrence pins its rounding order so that the one- and
    two-row-group instantiations agree; the row-MLP kernel's arithmetic per row does not depend on the row count).
    The reference decodes one sentence at a time (translate_mm_vi.py:80-82): batching must not change its output."""
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import synthetic, ops
    opt = synthetic.make_opt(conditional=True, dropout=0.5)
    fields = synthetic.make_fields(10000, 10000)
    torch.manual_seed(3435)
    model = vm.make_vi_model_mmt(opt, fields, gpu=True)
    model.eval()
    src, sl, _t, _tl, _img = synthetic.random_batch(10000, 10000, 250, 8, seed=77)
    src, sl = src.cuda(), sl.cuda()
    with torch.no_grad(), ops.batch_invariant():
        (hb, cb), ctx_b = model.encoder(src.unsqueeze(2), sl)
        zb = model.gen_net_global(ctx_b, sl)[0].mean()
        for i in [0, 1, 57, 123, 249]:
            n = int(sl[i])
            s1 = src[:n, i:i + 1].contiguous()
            (h1, c1), ctx_1 = model.encoder(s1.unsqueeze(2), sl[i:i + 1].contiguous())
            z1 = model.gen_net_global(ctx_1, sl[i:i + 1].contiguous())[0].mean()
            assert torch.equal(ctx_b[:n, i], ctx_1[:, 0]), f"context of sentence {i}"
            assert torch.equal(hb[:, i], h1[:, 0]) and torch.equal(cb[:, i], c1[:, 0]), f"final state of sentence {i}"
            assert torch.equal(zb[i], z1[0]), f"prior mean of sentence {i}"
