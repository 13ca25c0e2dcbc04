"""CPU: the oracle restatement (oracle/vi_model1_ref.py, oracle/beam_ref.py) against the fixtures
produced by the EXECUTED reference (oracle/make_golden.py -> tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_sample
from oracle import synth, beam_ref
from oracle import vi_model1_ref as R

torch.set_num_threads(8)


def _setup(name):
    meta, arr = load_golden(name)
    cfg = synth.ModelConfig(**meta["cfg"])
    params = synth.make_params(cfg, meta["param_seed"], meta["param_scale"])
    return meta, arr, cfg, params


def _close(a, b, rtol, atol, what):
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    err = np.abs(a - b).max()
    assert np.allclose(a, b, rtol=rtol, atol=atol), f"{what}: max abs err {err:.3e}"


STAT_KEYS = ("nmt", "td_kl_before", "td_kl_after", "img_feats_loss", "img_feats_cos", "elbo")


def _check_stats(stats, ref, rtol=2e-5):
    for k in STAT_KEYS:
        assert stats[k] == pytest.approx(ref[k], rel=rtol, abs=1e-5), k
    assert stats["n_words"] == ref["n_words"]
    assert stats["n_correct"] == ref["n_correct"]


@pytest.mark.parametrize("name", ["tiny_cond_eval", "tiny_fixed_eval", "tiny_brnn_eval", "cfg1_eval"])
def test_eval_forward_and_loss(name):
    meta, arr, cfg, params = _setup(name)
    batch = synth.make_batch(cfg, **meta["batch"])
    fwd, stats, _ = R.eval_step(params, cfg, batch)
    for k in ("out", "attn", "mu_q", "sd_q", "mu_p", "sd_p", "z", "img_loc"):
        _close(golden_sample(fwd[k].numpy()), arr[k], 1e-4, 2e-6, k)
    _check_stats(stats, meta["stats"])


@pytest.mark.parametrize("name", ["tiny_cond_train", "tiny_fixed_train", "tiny_brnn_train", "cfg1_train",
                                  "cfg1_fixed_train"])
def test_train_step_grads_and_update(name):
    meta, arr, cfg, params = _setup(name)
    batch = synth.make_batch(cfg, **meta["batch"])
    grads, stats, fwd = R.train_step_grads(params, cfg, batch)
    _check_stats(stats, meta["stats"])
    for k in ("out", "attn", "mu_q", "sd_q", "z"):
        _close(golden_sample(fwd[k].detach().numpy()), arr[k], 1e-4, 2e-6, k)
    gn = meta["extra"]["grad_norm"]
    assert sorted(k for k, g in grads.items() if g is None) == sorted(meta["extra"]["no_grad"])
    for k, ref_norm in gn.items():
        g = grads[k].numpy()
        assert float(np.linalg.norm(g.astype(np.float64))) == pytest.approx(ref_norm, rel=2e-4, abs=1e-9), k
        _close(golden_sample(g), arr["grad/" + k], 2e-3, 1e-7 + 2e-5 * ref_norm / np.sqrt(g.size), "grad " + k)
    # clip(5) + Adam(eps=1e-9) update (Optim.py:78-96)
    new, total = R.clip_and_adam({k: torch.as_tensor(v) for k, v in params.items()}, grads, {},
                                 lr=meta["extra"]["lr"])
    assert total == pytest.approx(meta["extra"]["total_grad_norm"], rel=2e-4)
    for k in gn:
        delta = new[k].double().numpy() - params[k].astype(np.float64)
        # first Adam step moves every touched weight by ~lr*sign(g); tiny |g| are ill-conditioned
        mask = np.abs(golden_sample(grads[k].numpy())) > 1e-7
        _close(golden_sample(delta)[mask], arr["delta/" + k][mask], 2e-2, 2e-5, "delta " + k)


@pytest.mark.parametrize("name", ["tiny_cond_beam5", "tiny_cond_greedy", "tiny_fixed_beam5"])
def test_beam_decode(name):
    meta, arr, cfg, params = _setup(name)
    ex = meta["extra"]
    batch = synth.make_batch(cfg, **meta["batch"])
    p = R._t(params)
    for i in range(ex["n_sent"]):
        L = int(batch.src_lengths[i])
        toks, score, attn = beam_ref.beam_search_one(
            p, cfg, torch.as_tensor(batch.src[:L, i]), beam_size=ex["beam"],
            max_length=ex["max_length"])
        assert toks == arr[f"tokens/{i}"].tolist(), f"sentence {i}"
        assert score == pytest.approx(float(arr[f"score/{i}"]), rel=1e-5, abs=1e-4)
        _close(attn.numpy(), arr[f"attn/{i}"], 1e-4, 1e-6, f"attn {i}")


def test_fp64_budget_tiny():
    """fp32 restatement vs fp64 restatement: the round-off floor the CUDA tolerances sit above."""
    meta, arr, cfg, params = _setup("tiny_cond_eval")
    batch = synth.make_batch(cfg, **meta["batch"])
    f32, s32, _ = R.eval_step(params, cfg, batch, torch.float32)
    f64, s64, _ = R.eval_step(params, cfg, batch, torch.float64)
    assert (f32["attn"].double() - f64["attn"]).abs().max() < 1e-6
    assert abs(s32["nmt"] - s64["nmt"]) / s64["nmt"] < 1e-6


def test_synth_is_deterministic():
    a = synth.uniform01(5, 3435, "x")
    assert np.allclose(a, synth.uniform01(5, 3435, "x"))
    # frozen values: the fixtures depend on this exact stream
    p = synth.make_params(synth.TINY, 3435, 0.1)
    assert synth.num_params(synth.CFG1) == 42781301
    w = p["generator.0.bias"]
    assert w.shape == (150,) and np.abs(w).max() <= 0.1


def test_cpu_baseline_fast_path_equals_restatement():
    """oracle/cpu_baseline.py (nn.LSTM + pack_padded_sequence, what bench.py times as the CPU arm)
    gives the same losses and gradients as the explicit-loop restatement."""
    import torch
    from oracle import cpu_baseline as cb
    for cfg in (synth.TINY, synth.TINY_FIXED):
        params = synth.make_params(cfg, 3435, 0.1)
        batch = synth.make_batch(cfg, batch_size=6, seed=2, t_force=20)
        m = cb.CpuStep(cfg, params, dropout=0.0)
        m.train()
        loss, stats, _ = m.forward_loss(R.to_torch_batch(batch), training=True)
        loss.backward()
        grads, ostats, _ = R.train_step_grads(params, cfg, batch)
        for k in ("nmt", "td_kl_before", "img_feats_loss", "elbo"):
            assert stats[k] == pytest.approx(ostats[k], rel=1e-5), k
        assert stats["n_words"] == ostats["n_words"]
        for k, kk in m._names.items():
            g = m.p[kk].grad
            if grads[k] is None:
                assert g is None or not bool(g.any()), k
                continue
            if "emb_luts" in k:
                g = g.clone()
                g[1] = 0
            assert float((g - grads[k]).norm() / grads[k].norm().clamp_min(1e-30)) < 1e-4, k
