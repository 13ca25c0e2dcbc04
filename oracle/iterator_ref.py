"""TEST INFRASTRUCTURE (oracle): per-example Python-loop restatement of the reference's training batch order.

Follows onmt/io/IO.py:382-393 (``OrderedIterator.create_batches``) with the iterator flags of
train_mm_vi_model1.py:127-205 (sort=False, sort_within_batch=True, repeat=False) and sort_key = len(ex.src)
(onmt/io/TextDataset.py:87-89).  ``pool`` / ``batch`` / ``Iterator.__iter__`` live in torchtext 0.2.3, a pinned
third-party dependency that is NOT under /root/reference (requirements.txt: torchtext==0.2.3); its published algorithm:

    def batch(data, batch_size):            # consecutive chunks of batch_size examples
    def pool(data, batch_size, key, random_shuffler):
        for p in batch(data, batch_size * 100):
            p_batch = batch(sorted(p, key=key), batch_size)
            for b in random_shuffler(list(p_batch)):
                yield b
    Iterator.data(): shuffled examples when train (shuffle=True), dataset order otherwise
    Iterator.__iter__: if sort_within_batch: (sort=False) minibatch.sort(key=sort_key, reverse=True)

Parity is "unpinned" against torchtext's own RNG (its RandomShuffler draws from Python's `random`); the shuffler is
injected, so the comparison is exact given the same permutations.  Only tests/ may import this module.
"""


def _batch(data, batch_size):
    minibatch = []
    for ex in data:
        minibatch.append(ex)
        if len(minibatch) == batch_size:
            yield minibatch
            minibatch = []
    if minibatch:
        yield minibatch


def ordered_batches(src_lengths, batch_size, train, shuffler):
    """-> list of lists of example indices.  ``shuffler(n)`` returns a permutation of range(n)."""
    n = len(src_lengths)
    key = lambda i: src_lengths[i]                                     # noqa: E731
    out = []
    if train:
        data = [int(i) for i in shuffler(n)]
        for p in _batch(data, batch_size * 100):
            p_batch = list(_batch(sorted(p, key=key), batch_size))
            perm = shuffler(len(p_batch))
            for j in perm:
                out.append(list(p_batch[int(j)]))
    else:
        for b in _batch(range(n), batch_size):
            out.append(sorted(b, key=key))
    for mb in out:
        mb.sort(key=key, reverse=True)
    return out


def pad_batch(seqs, pad, bos=None, eos=None, width=None):
    """Column-major padding as torchtext's Field.pad + numericalize produce it: [width][B] nested lists."""
    rows = [([bos] if bos is not None else []) + list(s) + ([eos] if eos is not None else []) for s in seqs]
    w = max(len(r) for r in rows) if width is None else width
    return [[(r[t] if t < len(r) else pad) for r in rows] for t in range(w)]
