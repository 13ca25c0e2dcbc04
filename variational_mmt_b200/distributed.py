"""Batch-sharded data parallelism: one process per GPU, one SUM all-reduce per step.

The reference has no multi-GPU path (train_mm_vi_model1.py:73-75 exits when more than one GPU id is
given); what an N-rank step must equal is the reference's own gradient accumulation over N batches
(onmt/TrainerMultimodal.py:342-346, 625-718, ``-accum_count N``): ``normalization`` is the sentence
count of ALL the batches, every batch back-propagates ``(NLL + image + KL) / normalization`` and the
gradients ADD, then one ``optim.step()``.  So: each rank divides its local loss by the GLOBAL sentence
count and the collective is SUM (not mean) over the flat gradient buffer (4 B per parameter), followed
by the same global-norm clip + Adam on every rank.  Whole batches are the unit of sharding because the
target encoder couples the examples inside a batch (SURVEY.md hazard H1).  Decode needs no collective:
sentences are independent and are simply dealt out to the ranks.

Works on any torch.distributed backend: "nccl" on the GPUs (NVLink 5 / NVSwitch), "gloo" in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def is_active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def init_from_env(backend=None, device=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*) -> (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def rank_world():
    return (dist.get_rank(), dist.get_world_size()) if is_active() else (0, 1)


def batches_of_rank(n_batches, rank=None, world=None):
    """Indices of the (whole) batches rank r trains on within one global step sequence: r, r + N, r + 2N, ...
    Global step s consumes batches [s*N, (s+1)*N); a trailing partial group is dropped by every rank alike."""
    if rank is None:
        rank, world = rank_world()
    usable = (n_batches // world) * world
    return list(range(rank, usable, world))


def sentences_of_rank(n_sentences, rank=None, world=None):
    """Contiguous [start, stop) slice of a test set for replica decoding (no collective)."""
    if rank is None:
        rank, world = rank_world()
    per, rem = divmod(n_sentences, world)
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)


def global_normalization(local_sentences, device=None):
    """Sentence count of the global step = sum of the per-rank batch sizes (TrainerMultimodal.py:342-346)."""
    if not is_active():
        return int(local_sentences)
    t = torch.tensor([int(local_sentences)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def all_reduce_gradients(flat_grads):
    """SUM over ranks, in place, of the flat gradient buffer (one collective per step)."""
    if is_active():
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def reduce_statistics(vec):
    """SUM over ranks of a VIStatistics vector {nmt_loss, n_words, n_correct, kl, img, cos, kl_after, elbo}
    (reporting only; returns a new tensor)."""
    out = vec.clone()
    if is_active():
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out
