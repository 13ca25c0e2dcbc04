"""CPU, world_size 2, gloo: the data-parallel contract of variational_mmt_b200.distributed.

No distributed reference exists; an N-rank step must equal the reference's own ``-accum_count N`` step over
the same N batches (TrainerMultimodal.py:342-346,625-718): normalization = total sentence count, gradients
summed, one optimiser update.  The arithmetic here is the CPU oracle (the product has no CPU path); what is
under test is the host logic: batch dealing, global normalization, SUM all-reduce of one flat buffer,
statistics reduction, and that every rank ends with identical parameters."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat(grads, keys):
    return torch.cat([grads[k].reshape(-1).float() for k in keys])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from variational_mmt_b200 import distributed as D
    from oracle import synth
    from oracle import vi_model1_ref as R
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and D.is_active()
    cfg = synth.TINY
    params = synth.make_params(cfg, 3435, 0.1)
    sizes = [5, 3, 4, 6]                                  # 4 batches -> 2 global steps of 2 batches
    mine = D.batches_of_rank(len(sizes))
    assert mine == [rank, rank + 2]
    keys = None
    state = {}
    for step, bi in enumerate(mine):
        batch = synth.make_batch(cfg, batch_size=sizes[bi], seed=100 + bi, t_force=20)
        norm = D.global_normalization(batch.batch_size)
        assert norm == sizes[2 * step] + sizes[2 * step + 1]
        grads, stats, _ = R.train_step_grads(params, cfg, batch, normalization=norm)
        if keys is None:
            keys = sorted(k for k, g in grads.items() if g is not None)
        flat = _flat(grads, keys)
        D.all_reduce_gradients(flat)                      # SUM, in place
        vec = D.reduce_statistics(torch.tensor([stats["nmt"], float(stats["n_words"])], dtype=torch.float64))
        # identical update on every rank from the reduced gradient
        off, red = 0, {}
        for k in keys:
            n = grads[k].numel()
            red[k] = flat[off: off + n].view_as(grads[k]); off += n
        params, _ = R.clip_and_adam(params, red, state)
        np.savez(os.path.join(out_dir, f"r{rank}_s{step}.npz"), flat=flat.numpy(), vec=vec.numpy(),
                 w=np.asarray(params["decoder.attn.linear_in.weight"]))
    s, e = D.sentences_of_rank(7)
    assert (s, e) == ((0, 4) if rank == 0 else (4, 7))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_accum_count_two(tmp_path):
    port = 29650 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    from oracle import synth
    from oracle import vi_model1_ref as R
    cfg = synth.TINY
    params = synth.make_params(cfg, 3435, 0.1)
    sizes = [5, 3, 4, 6]
    state = {}
    for step in range(2):
        bs = [synth.make_batch(cfg, batch_size=sizes[2 * step + j], seed=100 + 2 * step + j, t_force=20) for j in range(2)]
        norm = sum(b.batch_size for b in bs)              # accum_count = 2: normalization over both batches
        gs = [R.train_step_grads(params, cfg, b, normalization=norm) for b in bs]
        keys = sorted(k for k, g in gs[0][0].items() if g is not None)
        ref = _flat(gs[0][0], keys) + _flat(gs[1][0], keys)
        nmt = gs[0][1]["nmt"] + gs[1][1]["nmt"]
        words = gs[0][1]["n_words"] + gs[1][1]["n_words"]
        red = {k: gs[0][0][k] + gs[1][0][k] for k in keys}
        params, _ = R.clip_and_adam(params, red, state)
        got = [np.load(os.path.join(tmp_path, f"r{r}_s{step}.npz")) for r in range(2)]
        for g in got:
            assert np.allclose(g["flat"], ref.numpy(), rtol=1e-5, atol=1e-7)
            assert g["vec"][0] == pytest.approx(nmt, rel=1e-6) and int(g["vec"][1]) == words
            # Adam's first updates are ~lr*sign(g): elements whose gradient is at round-off level may flip with the
            # thread count of the matmuls, all others must agree
            dw = np.abs(g["w"] - np.asarray(params["decoder.attn.linear_in.weight"]))
            assert (dw > 1e-6).mean() < 0.01 and dw.max() <= 2.1 * 0.002 * (step + 1)
        assert np.array_equal(got[0]["w"], got[1]["w"])   # replicas stay bit-identical


def test_early_final_range_is_the_tail_of_the_flat_buffer():
    """distributed.early_final_begin: the latent / image networks and the generator (gradients final before the
    encoders' backward starts) form the tail of the flat parameter buffer, for both the conditional and fixed prior."""
    import pytest
    try:
        import variational_mmt_b200 as vm
        from variational_mmt_b200 import synthetic, distributed as D
    except ImportError as e:
        pytest.skip(str(e))
    early = ("inf_net_global.", "gen_net_global.", "inf_net_image.", "generator.")
    for cond in (True, False):
        opt = synthetic.make_opt(emb=32, hidden=64, z_dim=24, conditional=cond, dropout=0.0)
        model = vm.make_vi_model_mmt(opt, synthetic.make_fields(120, 150), gpu=False)
        begin = D.early_final_begin(model)
        off, seen, first = 0, set(), None
        for name, p in model.named_parameters():
            if id(p) in seen:
                continue
            seen.add(id(p))
            if name.startswith(early) and first is None:
                first = off
            assert name.startswith(early) == (first is not None), name        # nothing late-final after the first early one
            off += ((p.numel() + 3) // 4) * 4
        assert begin == first and 0 < begin < off and begin % 4 == 0
        flat = model._flat_params
        assert flat is not None and flat.numel() == off                       # same padding rule as flat.flatten_parameters

    class Odd(torch.nn.Module):                                               # an early module NOT at the tail: no split
        def __init__(self):
            super().__init__()
            self.generator = torch.nn.Linear(4, 4)
            self.encoder = torch.nn.Linear(4, 4)
    assert D.early_final_begin(Odd()) is None
