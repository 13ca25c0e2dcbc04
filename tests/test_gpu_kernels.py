"""-m gpu: per-kernel checks of the libvmmt C ABI against plain torch math (tests/kernel_checks.py)."""
import pytest

import kernel_checks as kc

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fp32_simt", "tf32_tc"])
def gemm_mode(request, cuda_device):
    from variational_mmt_b200 import _lib, ops
    ops.set_gemm_mode(1 if request.param == "fp32_simt" else 0)
    yield 1.0 if request.param == "fp32_simt" else 100.0     # TF32 operands: ~1e-3 relative
    ops.set_gemm_mode(0)


@pytest.mark.parametrize("group", [n for n, _ in kc.ALL])
def test_kernel_group(group, gemm_mode):
    fn = dict(kc.ALL)[group]
    bad = [(l, e, t) for l, e, t in fn() if not (e <= t * gemm_mode or (t == 0.0 and e == 0.0))]
    assert not bad, "\n".join(f"{l}: err {e:.3e} > tol {t * gemm_mode:.1e}" for l, e, t in bad)


def test_lstm_stepwise_path(gemm_mode, monkeypatch):
    """The large-batch path (one GEMM + one cell kernel per step, csrc/lstm_step.cu) on the same cases."""
    monkeypatch.setenv("VMMT_LSTM_STEPWISE", "1")
    bad = [(l, e, t) for l, e, t in kc.check_lstm() if not (e <= t * gemm_mode)]
    assert not bad, "\n".join(f"{l}: err {e:.3e} > tol {t * gemm_mode:.1e}" for l, e, t in bad)


def test_image_feature_table_gather(cuda_device):
    """Device-resident feature table + row gather by batch.indices == the reference's host fancy index
    (onmt/TrainerMultimodal.py:632-639)."""
    import numpy as np
    import torch
    from variational_mmt_b200 import io
    rng = np.random.RandomState(0)
    feats = np.abs(rng.normal(0, 1, (777, 2048))).astype(np.float32)
    table = io.ImageFeatureTable(feats, cuda_device)
    for idx in ([0], [776, 0, 5, 5, 123], rng.randint(0, 777, 40).tolist()):
        got = table.gather(torch.as_tensor(idx, dtype=torch.int64)).cpu().numpy()
        assert np.array_equal(got, feats[np.asarray(idx)])            # bit-exact: a copy


def test_gemm_bf16_variant(cuda_device):
    """The bf16 operand variant of the tcgen05 GEMM (BASELINE configs[1] "fp32 and bf16"; kernel_checks.check_gemm_bf16)."""
    from variational_mmt_b200 import ops
    ops.set_gemm_mode(2)
    try:
        bad = [(l, e, t) for l, e, t in kc.check_gemm_bf16() if not e <= t]
    finally:
        ops.set_gemm_mode(0)
    assert not bad, "\n".join(f"{l}: err {e:.3e} > tol {t:.1e}" for l, e, t in bad)
