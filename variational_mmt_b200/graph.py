"""CUDA-graph capture of the forward + loss + backward of one training batch.

The reference's trainer (onmt/TrainerMultimodal.py:625-718, ``_gradient_accumulation``) issues, per
batch, ``model.zero_grad()`` -> ``model(src, tgt, lengths, tgt_lengths, img_feats)`` ->
``train_loss.sharded_compute_loss(...)`` (which back-propagates) -> ``optim.step()``.  At batch 40 the
~180 kernels of that sequence take less device time than Python needs to launch them, so the same
sequence is captured ONCE per input-shape bucket (src_len, tgt_len, batch, normalization) into a CUDA
graph and replayed: inputs are copied into static device buffers, the statistics come back in a static
device vector.  The optimiser step (gradient all-reduce + clip + Adam: three launches) stays outside
the graph so that the NCCL collective is issued normally.

What makes replays differ from each other although every kernel argument is baked into the graph:
the Philox offsets of dropout / latent noise are taken relative to a device-resident counter
(``ops.rng_base``) that a one-thread kernel at the end of the graph advances.
"""
import os

import torch

from . import ops
from . import _lib


class _Batch(object):
    """The two attributes NMTVIModel1LossCompute reads (TrainerMultimodal.py:668-677)."""

    def __init__(self, tgt, batch_size):
        self.tgt, self.batch_size = tgt, batch_size


class GraphedTrainStep(object):
    """forward + sharded loss + backward as a replayed CUDA graph.

    step = GraphedTrainStep(model, loss_compute, shard_size=32)
    stats_vec = step(src [S,B], src_lengths [B], tgt [Tf,B], tgt_lengths [B], img_feats [B,D], normalization)
    optim.step()

    ``stats_vec`` is a device tensor of 8 floats {nmt_loss, n_words, n_correct, kl, img_logprob, img_cos,
    kl_after, elbo} (VIStatistics order) that is overwritten by the next call with the same shapes.
    Inputs may live on the host (pinned) or on the device; they are copied into the graph's static buffers.
    """

    def __init__(self, model, loss_compute, shard_size=32, max_graphs=128, optim=None):
        assert not loss_compute.use_kl_annealing, \
            "KL annealing changes a host-side weight every update: use the eager path"
        self.model, self.loss = model, loss_compute
        self.optim = optim               # told after every replay that the captured early gradient exchange has run
        self.shard_size, self.max_graphs = shard_size, max_graphs
        self._graphs = {}
        self._pool = None
        self._zero_stream = None
        self._capture_stream = None
        self.kernels_per_replay = 0          # libvmmt kernels inside the most recently captured graph
        self.device = next(model.parameters()).device
        ops.rng_base(self.device)

    def _param_addresses(self):
        """(parameter, gradient) addresses a captured graph has baked in; they move when Optim.set_parameters re-homes the
        flat buffers into the NVLink peer segment or the model changes device."""
        p = next(self.model.parameters())
        return (p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0)

    def _run(self, src, sl, tgt, tl, img, normalization):
        out, attns, _ = self.model(src.unsqueeze(2), tgt.unsqueeze(2), sl, tl, img)
        # the 171 MB gradient memset runs beside the generator / loss forward (nothing touches a gradient before the
        # backward pass).  Not earlier: its thousands of short blocks keep the encoder recurrences' thread-block
        # clusters from being placed (measured +30 us on the source encoder's first layer)
        cur = torch.cuda.current_stream(self.device)
        if self._zero_stream is None:
            self._zero_stream = torch.cuda.Stream(device=self.device)
        self._zero_stream.wait_stream(cur)
        with torch.cuda.stream(self._zero_stream):
            self.model.zero_grad()
        self.loss.before_backward = lambda: cur.wait_stream(self._zero_stream)
        st = self.loss.sharded_compute_loss(_Batch(tgt, tgt.size(1)), out, attns, 0, tgt.size(0),
                                            self.shard_size, normalization)
        if self.optim is not None and getattr(self.optim, "_early", None) is not None:
            self.optim.join_early()      # the early reduce-scatter forked inside the backward pass rejoins the step here
        return st._vec

    def _load_inputs(self, static, inputs):
        """Copies this step's batch into the graph's static buffers: one kernel launch for device-resident inputs (no copy
        engine: see vmmt_copy_list), cudaMemcpyAsync for host (pinned) inputs."""
        if all(t.is_cuda and t.is_contiguous() and t.dtype == s.dtype for s, t in zip(static, inputs)) and len(inputs) <= 5:
            import ctypes as C
            n = len(inputs)
            src = (C.c_void_p * n)(*[t.data_ptr() for t in inputs])
            dst = (C.c_void_p * n)(*[s.data_ptr() for s in static])
            nb = (C.c_int64 * n)(*[t.numel() * t.element_size() for t in inputs])
            _lib.call("vmmt_copy_list", src, dst, nb, n, _lib.stream())
            return
        for s, t in zip(static, inputs):
            s.copy_(t, non_blocking=True)

    def _capture(self, key, inputs, normalization):
        if len(self._graphs) >= self.max_graphs:
            raise RuntimeError("GraphedTrainStep: too many shape buckets; bucket the batches by length")
        static = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in inputs]
        for s, t in zip(static, inputs):
            s.copy_(t, non_blocking=True)
        n_updates = self.loss.n_model_updates
        # warm-up on a side stream: lazy one-time work (function attributes, occupancy queries, allocator growth)
        # must not happen inside the capture
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        # The warm-up passes must NOT fire the data-parallel early exchange: it is a cross-rank barrier + peer reads of every
        # rank's gradient segment, and the ranks meet new shape buckets at different steps (io.OrderedIterator deals each
        # rank its own batches) -- a capturing rank would run two barrier generations ahead of a replaying one.  The
        # exchange is part of the CAPTURED step only (recorded here, executed by every rank once per replay).
        hook = getattr(self.model, "early_exchange_hook", None)
        early_done = getattr(self.optim, "_early_done", False) if self.optim is not None else False
        self.model.early_exchange_hook = None
        try:
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._run(*static, normalization)
        finally:
            self.model.early_exchange_hook = hook
            if self.optim is not None:
                self.optim._early_done = early_done
        torch.cuda.current_stream(self.device).wait_stream(side)
        dot = os.environ.get("VMMT_GRAPH_DOT")              # debug: dump the captured step as DOT (cudaGraphDebugDotPrint)
        g = torch.cuda.CUDAGraph(keep_graph=True) if dot else torch.cuda.CUDAGraph()
        l0 = _lib.lib.vmmt_launch_count()
        # captured on a HIGH-priority stream (as are the branch / loss streams): kernel-node priorities are recorded, so
        # the critical chain's CTAs are placed before those of the weight-gradient lanes (default = lowest priority)
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(device=self.device, priority=int(os.environ.get("VMMT_MAIN_PRIO", "-1")))
        with torch.cuda.graph(g, pool=self._pool, stream=self._capture_stream):
            ops.begin_step()
            vec = self._run(*static, normalization)
            ops.advance_rng()
        if dot:
            g.debug_dump(dot)
            g.instantiate()
        self.kernels_per_replay = int(_lib.lib.vmmt_launch_count() - l0)
        if self._pool is None:
            self._pool = g.pool()
        self.loss.n_model_updates = n_updates           # the warm-up / capture passes are not model updates
        self._graphs[key] = (g, static, vec, self._param_addresses())
        return self._graphs[key]

    def __call__(self, src, src_lengths, tgt, tgt_lengths, img_feats, normalization):
        if src.dim() == 3:
            src = src[:, :, 0]
        if tgt.dim() == 3:
            tgt = tgt[:, :, 0]
        inputs = (src, src_lengths, tgt, tgt_lengths, img_feats)
        key = (tuple(src.shape), tuple(tgt.shape), tuple(img_feats.shape), float(normalization), self.model.training)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(key, inputs, normalization)
        g, static, vec, gen = entry
        if gen != self._param_addresses():
            raise RuntimeError("GraphedTrainStep: the parameters were moved to other buffers after this graph was captured "
                               "(Optim.set_parameters / model.to()): create the GraphedTrainStep after Optim.set_parameters")
        self._load_inputs(static, inputs)
        g.replay()
        if self.optim is not None and getattr(self.optim, "_early", None) is not None:
            self.optim._early_done = True          # the graph contains Optim.early_reduce_scatter (fired in its backward)
        self.loss.n_model_updates += 1
        return vec
