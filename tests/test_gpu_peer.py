"""-m gpu: the NVLink peer-memory optimiser step (csrc/peer.cu: reduce-scatter -> clip + Adam on the rank's slice ->
all-gather) against the replicated path (vmmt_sqnorm + vmmt_adam_clip_step, Optim.py:69-70,94-96 semantics).

Single GPU: world = 1 exercises the same kernels with the rank as its own only peer.  Two GPUs (skipped otherwise):
the peer step must leave both replicas identical and equal to the NCCL all-reduce step on the same gradients.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _toy_module(n_sizes, dev):
    import variational_mmt_b200  # noqa: F401
    from variational_mmt_b200.flat import FlatParamsMixin

    class Toy(FlatParamsMixin, torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(7)
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(n, generator=g) * 0.1) for n in n_sizes])
    m = Toy().to(dev)
    m.flatten_parameters()
    return m


def test_peer_step_world1_equals_replicated_step(cuda_device):
    import variational_mmt_b200 as vm
    sizes = [1000, 37, 4096 * 33 + 5, 3, 250000]        # ragged sizes: every tensor is padded to 4 floats
    ma, mb = _toy_module(sizes, cuda_device), _toy_module(sizes, cuda_device)
    oa = vm.Optim("adam", 0.002, 5, exchange="nccl")
    ob = vm.Optim("adam", 0.002, 5, exchange="peer")
    oa.set_parameters(ma.parameters())
    ob.set_parameters(mb.parameters())
    assert ob.peer is not None and ob.peer.world == 1 and oa.peer is None
    assert torch.equal(oa.flat, ob.flat)
    g = torch.Generator(device="cuda").manual_seed(11)
    for it in range(4):
        scale = 3.0 if it % 2 == 0 else 1e-3              # clipped and unclipped steps
        grad = torch.randn(oa.gflat.numel(), device=cuda_device, generator=g) * scale
        ma.zero_grad(); mb.zero_grad()
        oa.gflat.copy_(grad); ob.gflat.copy_(grad)
        oa.step(); ob.step()
        torch.cuda.synchronize()
        ref_sq = float((grad.double() ** 2).sum())
        assert abs(float(ob._sq) - ref_sq) <= 1e-5 * ref_sq
        assert abs(float(oa._sq) - ref_sq) <= 1e-5 * ref_sq
        # same arithmetic per element; only the clip coefficient may differ in its last bit
        assert torch.allclose(oa.flat, ob.flat, rtol=2e-6, atol=1e-8), float((oa.flat - ob.flat).abs().max())
    # parameters are still views of the (peer) flat buffer
    assert mb.ps[0].data_ptr() == ob.flat.data_ptr()
    assert mb.ps[0].grad.data_ptr() == ob.gflat.data_ptr()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    import variational_mmt_b200 as vm
    from variational_mmt_b200 import distributed as D
    D.init_from_env(backend="nccl", device=dev)
    sizes = [1000, 37, 4096 * 33 + 5, 3, 250000]
    res = {}
    for mode in ("nccl", "peer"):
        m = _toy_module(sizes, dev)
        o = vm.Optim("adam", 0.002, 5, exchange=mode)
        o.set_parameters(m.parameters())
        if mode == "peer":
            assert o.peer is not None and o.peer.world == world, "peer exchange did not come up"
        for it in range(5):
            g = torch.Generator(device="cuda").manual_seed(1000 * it + rank)     # a different gradient per rank
            grad = torch.randn(o.gflat.numel(), device=dev, generator=g) * (2.0 if it % 2 == 0 else 1e-3)
            m.zero_grad()
            o.gflat.copy_(grad)
            o.step()
        torch.cuda.synchronize()
        res[mode] = o.flat.cpu().numpy().copy()
        res[mode + "_sq"] = float(o._sq) if mode == "peer" else float(o.grad_norm() ** 2)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), nccl=res["nccl"], peer=res["peer"], peer_sq=res["peer_sq"],
             nccl_sq=res["nccl_sq"])
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_peer_step_equals_nccl_step(tmp_path, world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29650 + (os.getpid() % 100) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rs = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    for r in rs[1:]:
        assert np.array_equal(rs[0]["peer"], r["peer"])                    # replicas bit-identical
        assert np.array_equal(rs[0]["nccl"], r["nccl"])
    # vs NCCL: the same N-term sums in another order (and for N = 2 the very same sum); a near-cancelling sum may
    # flip the sign of Adam's first normalised step, hence a fraction instead of allclose
    bad = ~np.isclose(rs[0]["peer"], rs[0]["nccl"], rtol=2e-6, atol=1e-8)
    assert bad.mean() < 1e-5, (int(bad.sum()), float(np.abs(rs[0]["peer"] - rs[0]["nccl"]).max()))
    assert abs(float(rs[0]["peer_sq"]) - float(rs[0]["nccl_sq"])) <= 1e-5 * float(rs[0]["nccl_sq"])


def test_split_exchange_world1_equals_replicated_step(cuda_device):
    """Optim.enable_early_exchange on one GPU (world = 1: the rank is its own peer): the encoder backward fires the
    early reduce-scatter of the flat buffer's tail, step() exchanges the rest, and two training steps of the real
    (tiny) model end on the parameters the replicated clip + Adam produces."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import variational_mmt_b200 as vm
    from gpu_helpers import build_cuda_model, to_device
    from oracle import synth
    cfg = synth.TINY
    params = synth.make_params(cfg, 3435, 0.1)
    batch = synth.make_batch(cfg, batch_size=6, seed=3, t_force=20)
    res = {}
    for mode in ("nccl", "peer", "peer_early"):
        model, fields = build_cuda_model(cfg, params)
        model.train()
        loss = vm.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
        optim = vm.Optim("adam", 0.002, 5, exchange="nccl" if mode == "nccl" else "peer")
        optim.set_parameters(model.parameters())
        if mode == "peer_early":
            assert optim.enable_early_exchange(model)
            n, b0 = optim.flat.numel(), optim._early["begin"]
            names = [k for k, _ in model.named_parameters()]
            assert 0 < b0 < n and any(k.startswith("generator.") for k in names)
        d = to_device(batch)
        for _ in range(2):
            model.zero_grad()
            with vm.Normal.inject_noise(d.eps):
                out, attns, _ = model(d.src, d.tgt_in, d.src_lengths, d.tgt_lengths, d.img_feats)
            loss.sharded_compute_loss(d, out, attns, 0, d.tgt.size(0), 32, d.batch_size)
            if mode == "peer_early":
                assert optim._early_done, "the encoder backward did not fire the early reduce-scatter"
            optim.step()
        torch.cuda.synchronize()
        res[mode] = optim.flat.cpu().numpy()
        res[mode + "_sq"] = float(optim._sq)
    for mode in ("peer", "peer_early"):
        bad = ~np.isclose(res[mode], res["nccl"], rtol=1e-5, atol=2e-6)
        assert bad.mean() < 2e-4, (mode, int(bad.sum()), float(np.abs(res[mode] - res["nccl"]).max()))
        assert abs(res[mode + "_sq"] - res["nccl_sq"]) <= 1e-4 * res["nccl_sq"]
