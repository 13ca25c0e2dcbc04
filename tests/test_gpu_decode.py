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
    from variational_mmt_b200 import _lib
    _lib.lib.vmmt_set_gemm_mode(gemm_mode)
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
    assert a.shape == arr[f"attn/{i}"].shape
    assert np.abs(a - arr[f"attn/{i}"]).max() <= tol


@pytest.mark.parametrize("mode", [1, 0], ids=["fp32_simt", "tf32_tc"])
@pytest.mark.parametrize("name", ["tiny_cond_beam5", "tiny_cond_greedy", "tiny_fixed_beam5"])
def test_sentence_by_sentence_matches_reference(name, mode, cuda_device):
    """batch_size 1, as translate_mm_vi.py runs it."""
    from variational_mmt_b200 import _lib
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
        _lib.lib.vmmt_set_gemm_mode(0)


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("name", ["tiny_cond_beam5", "tiny_cond_greedy", "tiny_fixed_beam5"])
def test_batched_decode_equals_sentence_by_sentence(name, graph, cuda_device):
    """All sentences of the (padded, length-sorted) batch advance together: new behaviour whose oracle is
    the reference run one sentence at a time."""
    from variational_mmt_b200 import _lib
    try:
        tr, batch, arr, ex = _translator(name, 1)
        tr.poll_every = 1
        tr.use_graph = graph
        b = _B()
        b.batch_size = ex["n_sent"]
        b.src = (torch.as_tensor(batch.src), torch.as_tensor(batch.src_lengths))
        for rep in range(2):                      # second call reuses the bucket's buffers / graph
            ret = tr.translate_batch(b, None, list(range(ex["n_sent"])))
            for i in range(ex["n_sent"]):
                _check(ret, i, arr, i, 1e-4)
    finally:
        _lib.lib.vmmt_set_gemm_mode(0)


def test_in_process_validation_translate_equals_sentence_by_sentence(cuda_device, tmp_path):
    """translate_dataset (batched, in place on the live training model) returns, in corpus order, what
    the per-sentence decode of translate_mm_vi.py returns, and restores train mode."""
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import _lib, io
    _lib.lib.vmmt_set_gemm_mode(1)
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
        _lib.lib.vmmt_set_gemm_mode(0)
