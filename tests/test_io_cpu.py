"""CPU: the vectorised batch planner / padder (variational_mmt_b200/io.py) against the per-example loop restatement
of the reference's iterator (oracle/iterator_ref.py: onmt/io/IO.py:382-393 over torchtext 0.2.3 pool/batch)."""
import numpy as np
import pytest

from oracle import iterator_ref


def _io():
    try:
        from variational_mmt_b200 import io
    except ImportError as e:                                   # libvmmt.so not built in this environment
        pytest.skip(str(e))
    return io


@pytest.mark.parametrize("n,bs,train", [(0, 4, True), (1, 4, True), (7, 4, False), (1000, 40, True), (4321, 40, True),
                                        (4321, 40, False), (400, 3, True)])
def test_batch_plan_equals_loop_restatement(n, bs, train):
    io = _io()
    ds = io.TripletDataset.synthetic(n, seed=5) if n else io.TripletDataset([], [])
    got = io.plan_batches(ds.src_len, bs, train, io.seeded_shuffler(9))
    want = iterator_ref.ordered_batches(ds.src_len.tolist(), bs, train, io.seeded_shuffler(9))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.tolist() == w
    if n:
        assert sorted(np.concatenate(got).tolist()) == list(range(n))          # every example exactly once
        for g in got:
            assert np.all(np.diff(ds.src_len[g]) <= 0)                         # decreasing source lengths


def test_padding_equals_loop_restatement_and_bucketing_only_adds_pad():
    io = _io()
    ds = io.TripletDataset.synthetic(500, v_src=50, v_tgt=60, seed=1)
    it = io.OrderedIterator(ds, 16, train=True, seed=3, device=None, rank=0, world=1)
    itb = io.OrderedIterator(ds, 16, train=True, seed=3, device=None, bucket=8, rank=0, world=1)
    n_seen = 0
    for b, bb in zip(it, itb):
        idx = b.indices.numpy()
        src_seqs = [ds.src_flat[ds.src_off[i]: ds.src_off[i + 1]] for i in idx]
        tgt_seqs = [ds.tgt_flat[ds.tgt_off[i]: ds.tgt_off[i + 1]] for i in idx]
        assert b.src[0].numpy().tolist() == iterator_ref.pad_batch(src_seqs, io.PAD)
        assert b.tgt.numpy().tolist() == iterator_ref.pad_batch(tgt_seqs, io.PAD, io.BOS, io.EOS)
        assert b.src[1].tolist() == [len(s) for s in src_seqs]
        assert b.tgt_lengths.tolist() == [len(s) + 2 for s in tgt_seqs]
        # bucketed variant: same content, padded up to a multiple of 8 with pad only
        S, T = b.src[0].shape[0], b.tgt.shape[0]
        assert bb.src[0].shape[0] % 8 == 0 and bb.tgt.shape[0] % 8 == 0
        assert np.array_equal(bb.src[0].numpy()[:S], b.src[0].numpy()) and np.all(bb.src[0].numpy()[S:] == io.PAD)
        assert np.array_equal(bb.tgt.numpy()[:T], b.tgt.numpy()) and np.all(bb.tgt.numpy()[T:] == io.PAD)
        n_seen += b.batch_size
    assert n_seen == 500


def test_ranks_deal_whole_batches():
    io = _io()
    ds = io.TripletDataset.synthetic(1000, seed=2)
    plans = [io.OrderedIterator(ds, 40, train=True, seed=7, rank=r, world=4).create_batches() for r in range(4)]
    full = io.plan_batches(ds.src_len, 40, True, io.seeded_shuffler(7))
    usable = (len(full) // 4) * 4
    for r in range(4):
        assert [p.tolist() for p in plans[r]] == [full[i].tolist() for i in range(r, usable, 4)]
