"""Host side of the data feed.

* ``OrderedIterator`` -- the reference's batch order (onmt/io/IO.py:382-393 ``OrderedIterator.create_batches`` over
  torchtext 0.2.3's ``pool`` / ``batch``; built by train_mm_vi_model1.py:127-205 with ``sort=False, train=is_train,
  sort_within_batch=True, repeat=False`` and ``sort_key = len(ex.src)``, onmt/io/TextDataset.py:87-89):
    train: shuffle the examples -> pools of 100 batches -> stable sort of each pool by source length -> batches of
           ``batch_size`` -> shuffle the batches of the pool;   eval: consecutive batches, each stably sorted;
    every mini-batch is then stably re-sorted by DECREASING source length (``sort_within_batch``, what
    pack_padded_sequence needs, train_mm_vi_model1.py:179-186).
  torchtext is not vendored in the reference (requirements.txt pins torchtext==0.2.3), so the published algorithm is
  restated; oracle/iterator_ref.py is the per-example Python-loop restatement the CPU tests compare against.  Here the
  whole epoch is planned with numpy (argsort per pool) and each batch is padded by one vectorised gather out of flat
  token arrays into pinned staging buffers -- cheap enough to feed 8 GPUs from one core per rank.
  ``bucket`` rounds the padded source / target lengths up to a multiple, so GraphedTrainStep sees a handful of (S, T)
  shapes instead of one per length pair.  The extra rows are pad (id 1): masked in the encoder / attention, ignored
  by the loss, independent rows of the target encoder (hazard H1 couples the batch axis, not the time axis).
* ``ImageFeatureTable`` -- ``train_img_feats[batch.indices]`` (onmt/TrainerMultimodal.py:632-639: numpy fancy index on
  the host + H2D copy of B x 2048 floats per step) becomes a table resident in HBM (238 MB for 29 K rows, 1.2 GB for
  145 K) and a row gather on the device by ``batch.indices`` (vmmt_embedding_fwd: same row-gather kernel).
"""
import numpy as np
import torch

from . import _lib as L
from ._lib import fptr, ptr, stream
from . import distributed

PAD, BOS, EOS = 1, 2, 3


class TripletDataset(object):
    """Token ids of a (source, target, image index) corpus in flat arrays: example i is
    ``src_flat[src_off[i]:src_off[i+1]]`` / ``tgt_flat[tgt_off[i]:tgt_off[i+1]]`` (target WITHOUT <s> / </s>: the
    iterator adds them, as torchtext's target Field does with init_token / eos_token)."""

    def __init__(self, src, tgt):
        self.src_flat, self.src_off = self._flatten(src)
        self.tgt_flat, self.tgt_off = self._flatten(tgt)
        assert len(self.src_off) == len(self.tgt_off)
        self.src_len = np.diff(self.src_off)
        self.tgt_len = np.diff(self.tgt_off)

    @staticmethod
    def _flatten(seqs):
        lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
        off = np.zeros(len(seqs) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        flat = np.concatenate([np.asarray(s, np.int64) for s in seqs]) if len(seqs) else np.zeros(0, np.int64)
        return flat, off

    def __len__(self):
        return len(self.src_len)

    @classmethod
    def synthetic(cls, n, v_src=10000, v_tgt=10000, seed=3435, src_max=50, tgt_max=48):
        """Multi30k-shaped corpus: lengths ~ clip(round(N(14,5)),3,50), ids uniform in [4,V)."""
        rng = np.random.RandomState(seed)
        sl = np.clip(np.rint(rng.normal(14, 5, n)), 3, src_max).astype(np.int64)
        tl = np.clip(np.rint(rng.normal(14, 5, n)), 3, tgt_max).astype(np.int64)
        ds = cls.__new__(cls)
        ds.src_len, ds.tgt_len = sl, tl
        ds.src_off = np.concatenate([[0], np.cumsum(sl)]).astype(np.int64)
        ds.tgt_off = np.concatenate([[0], np.cumsum(tl)]).astype(np.int64)
        ds.src_flat = rng.randint(4, v_src, int(sl.sum())).astype(np.int64)
        ds.tgt_flat = rng.randint(4, v_tgt, int(tl.sum())).astype(np.int64)
        return ds


class Batch(object):
    """What the trainer reads from a torchtext batch: ``src = (ids [S,B], lengths [B])``,
    ``tgt`` [T,B] with <s> ... </s>, ``tgt_lengths`` (counting both), ``indices`` [B] (corpus row of each example),
    ``batch_size``."""

    def __init__(self, src, src_lengths, tgt, tgt_lengths, indices):
        self.src = (src, src_lengths)
        self.tgt = tgt
        self.tgt_lengths = tgt_lengths
        self.indices = indices
        self.batch_size = int(src.shape[1])


def seeded_shuffler(seed):
    """``shuffler(n) -> permutation of range(n)``; successive calls continue one RandomState stream (torchtext's
    RandomShuffler keeps one `random` state across calls in the same way)."""
    rng = np.random.RandomState(seed)
    return lambda n: rng.permutation(n)


def plan_batches(src_len, batch_size, train, shuffler, pool_factor=100):
    """-> list of int64 index arrays, one per mini-batch, in the order the reference's iterator yields them."""
    n = len(src_len)
    out = []
    if train:
        order = np.asarray(shuffler(n), np.int64)
        pool = batch_size * pool_factor
        for p0 in range(0, n, pool):
            p = order[p0: p0 + pool]
            p = p[np.argsort(src_len[p], kind="stable")]                       # sorted(p, key=sort_key)
            chunks = [p[b0: b0 + batch_size] for b0 in range(0, len(p), batch_size)]
            for j in np.asarray(shuffler(len(chunks)), np.int64):              # random_shuffler(list(p_batch))
                out.append(chunks[int(j)])
    else:
        for b0 in range(0, n, batch_size):
            b = np.arange(b0, min(n, b0 + batch_size), dtype=np.int64)
            out.append(b[np.argsort(src_len[b], kind="stable")])               # sorted(b, key=sort_key)
    # sort_within_batch with sort=False: minibatch.sort(key=sort_key, reverse=True)  (stable, decreasing)
    return [b[np.argsort(-src_len[b], kind="stable")] for b in out]


def _round_up(x, m):
    return ((int(x) + m - 1) // m) * m


class OrderedIterator(object):
    """One epoch of mini-batches of a TripletDataset in the reference's order, padded on the host, optionally moved
    to ``device``.  With torch.distributed up, rank r yields batches r, r+N, ... of the global sequence (whole batches
    are the unit of sharding); a trailing partial group is dropped by every rank alike."""

    def __init__(self, dataset, batch_size, train=True, seed=3435, device=None, bucket=1, rank=None, world=None,
                 shuffler=None, pin=None):
        self.dataset, self.batch_size, self.train = dataset, batch_size, train
        self.device, self.bucket = device, max(1, int(bucket))
        if rank is None:
            rank, world = distributed.rank_world()
        self.rank, self.world = rank, world
        self.shuffler = shuffler if shuffler is not None else seeded_shuffler(seed)
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.batches = None

    def create_batches(self):
        plan = plan_batches(self.dataset.src_len, self.batch_size, self.train, self.shuffler)
        if self.world > 1:
            plan = [plan[i] for i in distributed.batches_of_rank(len(plan), self.rank, self.world)]
        self.batches = plan
        return plan

    def __len__(self):
        if self.batches is None:
            self.create_batches()
        return len(self.batches)

    def _pad(self, flat, off, lens, idx, width, bos_eos):
        """[width, B] int64: column b = tokens of example idx[b] (optionally <s> ... </s>), pad id elsewhere."""
        B = len(idx)
        ln = lens[idx]
        pos = np.arange(width, dtype=np.int64)[:, None]                         # [W,1]
        shift = 1 if bos_eos else 0
        src_pos = off[idx][None, :] + pos - shift                               # [W,B]
        valid = (pos >= shift) & (pos < ln[None, :] + shift)
        out = np.where(valid, flat[np.clip(src_pos, 0, max(len(flat) - 1, 0))], PAD) if len(flat) else \
            np.full((width, B), PAD, np.int64)
        if bos_eos:
            out[0, :] = BOS
            out[ln + 1, np.arange(B)] = EOS
        return np.ascontiguousarray(out, dtype=np.int64)

    def make_batch(self, idx):
        ds = self.dataset
        sl = ds.src_len[idx]
        tl = ds.tgt_len[idx] + 2
        S = _round_up(sl.max(), self.bucket)
        T = _round_up(tl.max(), self.bucket)
        arrs = (self._pad(ds.src_flat, ds.src_off, ds.src_len, idx, S, False), sl.astype(np.int64),
                self._pad(ds.tgt_flat, ds.tgt_off, ds.tgt_len, idx, T, True), tl.astype(np.int64),
                np.asarray(idx, np.int64))
        ts = [torch.from_numpy(a) for a in arrs]
        if self.device is not None and torch.device(self.device).type == "cuda":
            ts = [(t.pin_memory() if self.pin else t).to(self.device, non_blocking=True) for t in ts]
        return Batch(*ts)

    def __iter__(self):
        self.create_batches()
        for idx in self.batches:
            yield self.make_batch(idx)


class ImageFeatureTable(object):
    """Pooled image features [N, D] fp32 resident on the device; ``gather(batch.indices) -> [B, D]`` on the device.
    Replaces ``torch.from_numpy(train_img_feats[idxs]).cuda()`` (onmt/TrainerMultimodal.py:632-639)."""

    def __init__(self, feats, device):
        t = torch.as_tensor(np.ascontiguousarray(feats, dtype=np.float32) if isinstance(feats, np.ndarray) else feats)
        assert t.dim() == 2 and t.dtype == torch.float32
        self.device = torch.device(device)
        assert self.device.type == "cuda", "ImageFeatureTable lives in HBM (no CPU fallback)"
        self.table = t.to(self.device).contiguous()
        self.n, self.dim = self.table.shape

    def nbytes(self):
        return self.table.numel() * 4

    def gather(self, indices, out=None):
        idx = indices.to(self.device, dtype=torch.int64, non_blocking=True).contiguous()
        B = idx.numel()
        if out is None:
            out = torch.empty(B, self.dim, device=self.device, dtype=torch.float32)
        L.call("vmmt_embedding_fwd", ptr(idx), B, fptr(self.table), self.table.shape[0], self.dim, fptr(out), stream())
        return out

    __call__ = gather
