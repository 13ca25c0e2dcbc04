"""-m gpu: checkpoint / resume contract of Optim (reference: onmt/TrainerMultimodal.py:576-587 pickles the whole Optim
into the checkpoint; train_mm_vi_model1.py:433-452 unpickles it, calls optim.optimizer.load_state_dict(...) and then
optim.set_parameters(model.parameters()), which in the reference builds a FRESH torch.optim.Adam)."""
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _toy(dev, seed=7):
    import variational_mmt_b200  # noqa: F401
    from variational_mmt_b200.flat import FlatParamsMixin

    class Toy(FlatParamsMixin, torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(seed)
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(n, generator=g) * 0.1) for n in (1000, 37, 50003)])
    m = Toy().to(dev)
    m.flatten_parameters()
    return m


def _grad(optim, it):
    g = torch.Generator(device="cuda").manual_seed(100 + it)
    optim.gflat.copy_(torch.randn(optim.gflat.numel(), device="cuda", generator=g) * (0.5 if it % 2 else 2.0))


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_optim_pickles_and_resumes(cuda_device, exchange):
    import variational_mmt_b200 as vm
    m = _toy(cuda_device)
    o = vm.Optim("adam", 0.002, 5, lr_decay=0.5, start_decay_at=8, exchange=exchange)
    o.set_parameters(m.parameters())
    for it in range(3):
        _grad(o, it)
        o.step()
    o.update_learning_rate(10.0, 9)                       # start_decay_at reached: lr halves, bookkeeping changes
    sd = o.optimizer.state_dict()
    assert sd["step"] == 3 and sd["exp_avg"].numel() == o.flat.numel() and float(sd["exp_avg"].abs().sum()) > 0
    blob = pickle.dumps(o)                                # what drop_checkpoint does
    params_then = o.flat.clone()

    # (1) reference flow: unpickle -> optimizer.load_state_dict(own state) -> set_parameters: bookkeeping survives,
    #     Adam itself starts fresh (a new torch.optim.Adam in the reference)
    m2 = _toy(cuda_device)
    m2.ps[0].data.view(-1)                                # bound to its own flat buffer
    o2 = pickle.loads(blob)
    assert (o2._step, o2.lr, o2.start_decay, o2.last_ppl) == (3, 0.001, True, 10.0)
    o2.optimizer.load_state_dict(o2.optimizer.state_dict())
    o2.set_parameters(m2.parameters())
    assert o2._adam_t == 0 and float(o2.optimizer.state_dict()["exp_avg"].abs().sum()) == 0.0

    # (2) keep_state=True: the continued run equals the uninterrupted one, bit for bit
    m3 = _toy(cuda_device)
    o3 = pickle.loads(blob)
    o3.exchange = exchange
    o3.set_parameters(m3.parameters(), keep_state=True)
    o3.flat.copy_(params_then)
    assert o3._adam_t == 3
    for it in range(3, 5):
        _grad(o, it); o.step()
        _grad(o3, it); o3.step()
    torch.cuda.synchronize()
    assert torch.equal(o.flat, o3.flat)
    assert o3._step == o._step == 5


def test_sgd_with_clipping(cuda_device):
    """method='sgd' with max_grad_norm > 0 (the reference default optimiser, opts.py): p -= lr * g * min(1, c / ||g||)."""
    import variational_mmt_b200 as vm
    m = _toy(cuda_device)
    o = vm.Optim("sgd", 1.0, 5)
    o.set_parameters(m.parameters())
    before = o.flat.clone()
    _grad(o, 0)
    g = o.gflat.clone()
    o.grad_norm()
    o.step()
    torch.cuda.synchronize()
    norm = float(g.double().norm())
    ref = before - g * min(1.0, 5.0 / (norm + 1e-6))
    assert torch.allclose(o.flat, ref, rtol=1e-5, atol=1e-7)
