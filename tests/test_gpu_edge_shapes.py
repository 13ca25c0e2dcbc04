"""-m gpu: edge shapes of the training step against the CPU oracle (oracle/vi_model1_ref.py), exact-fp32 and TF32 modes.

The reference has no tests of its own; these are the shapes its data pipeline can produce at the
extremes: a single-sentence batch, source sentences of one token, targets of one word (<s> w </s>), ragged lengths,
a target longer than the 32-position shard the training loss scores (hazard H4), and a batch size that is not a
multiple of any tile (prime B)."""
import numpy as np
import pytest
import torch

from gpu_helpers import build_cuda_model, to_device, named_grads, relerr, attn_max_rel
from oracle import synth
from oracle import vi_model1_ref as R

pytestmark = pytest.mark.gpu


def _custom_batch(cfg, src_lens, tgt_words, seed):
    """Batch with the given per-row source lengths (descending) and target word counts."""
    B = len(src_lens)
    b = synth.make_batch(cfg, batch_size=B, seed=seed, full_length=(int(max(src_lens)), int(max(tgt_words))))
    sl = np.asarray(src_lens, np.int64)
    tl = np.asarray(tgt_words, np.int64) + 2
    assert np.all(np.diff(sl) <= 0), "rows must be sorted by decreasing source length"
    src, tgt = b.src.copy(), b.tgt.copy()
    for i in range(B):
        src[sl[i]:, i] = synth.PAD
        tgt[tl[i] - 1, i] = synth.EOS
        tgt[tl[i]:, i] = synth.PAD
    return synth.Batch(src, sl, tgt, tl, b.img_feats, b.eps)


CASES = {
    "single_sentence": ([7], [5]),
    "one_token_sources": ([1, 1, 1], [4, 2, 6]),
    "one_word_targets": ([5, 4, 2], [1, 1, 1]),
    "ragged_prime_batch": ([12, 11, 9, 9, 6, 3, 1], [3, 14, 1, 8, 20, 2, 5]),
    "target_longer_than_shard": ([6, 5], [40, 35]),          # 42 target positions: only the first 32 are scored (H4)
}


@pytest.mark.parametrize("mode", [1, 0], ids=["fp32_simt", "tf32_tc"])
@pytest.mark.parametrize("conditional", [True, False], ids=["conditional", "fixed_prior"])
@pytest.mark.parametrize("case", list(CASES))
def test_edge_shape_train_step_matches_oracle(case, conditional, mode, cuda_device):
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import _lib, ops
    ltol, gtol = (2e-5, 2e-4) if mode == 1 else (1e-3, 3e-2)
    ops.set_gemm_mode(mode)
    try:
        cfg = synth.ModelConfig(**{**synth.TINY.to_dict(), "conditional": conditional})
        params = synth.make_params(cfg, 3435, 0.1)
        batch = _custom_batch(cfg, *CASES[case], seed=11)
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
        assert st.nmt_loss == pytest.approx(ostats["nmt"], rel=ltol, abs=ltol)
        assert st.td_kl_before == pytest.approx(ostats["td_kl_before"], rel=ltol, abs=ltol)
        assert st.image_feats_loss == pytest.approx(ostats["img_feats_loss"], rel=ltol, abs=ltol)
        a = attns["std"].detach().cpu().numpy()
        assert attn_max_rel(a, ofwd["attn"].detach().numpy(), batch.src_lengths) <= (2e-5 if mode == 1 else 1e-3)
        assert np.allclose(a.sum(-1), 1.0, atol=1e-5)                       # every row of the masked softmax sums to 1
        for i, n in enumerate(batch.src_lengths):                           # and puts no mass on padding
            assert not np.any(a[:, i, int(n):])
        grads = named_grads(model)
        for k, og in ograds.items():
            if og is None or not np.any(og.numpy()):
                assert grads[k] is None or not np.any(grads[k]), k
                continue
            floor = 1e-7 if mode == 1 else 1e-4 * float(np.abs(og.numpy()).max())
            err = np.linalg.norm(grads[k].astype(np.float64) - og.numpy()) / max(np.linalg.norm(og.numpy()), floor)
            assert err <= gtol, f"grad {k}: rel err {err:.3e}"
    finally:
        ops.set_gemm_mode(0)
