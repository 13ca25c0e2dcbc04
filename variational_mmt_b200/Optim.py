"""Optimiser wrapper (reference: onmt/Optim.py:5-114): same constructor / set_parameters / step /
update_learning_rate contract; 'adam' (the published method) runs as a fused global-norm clip +
Adam(eps=1e-9) over the flat buffers in two launches.  Data parallel:
on one NVSwitch box the flat buffers live in NVLink peer memory and the step is reduce-scatter -> clip +
Adam on this rank's 1/N slice -> all-gather inside the update kernels (distributed.PeerExchange,
csrc/peer.cu); otherwise one NCCL SUM all-reduce of the flat gradient buffer precedes the update.
``exchange``: "auto" (peer when torch.distributed is up on CUDA, else none), "peer", "nccl"; the
environment variable VMMT_DP_EXCHANGE overrides "auto".
"""
import ctypes
import os

import torch

from . import _lib as L
from ._lib import fptr, stream
from .flat import flatten_parameters, owner_of, padded_numel
from . import distributed, ops


def ctypes_i64():
    return ctypes.byref(ctypes.c_int64())


class _ParamList(torch.nn.Module):
    def __init__(self, params):
        super().__init__()
        self.plist = torch.nn.ParameterList(params)


class _AdamStateView(object):
    """What the reference's callers reach through ``optim.optimizer`` (a torch.optim.Adam there, onmt/Optim.py:69-70):
    ``state_dict()`` / ``load_state_dict()`` (train_mm_vi_model1.py:433-437) and ``param_groups[0]['lr']``.  The state is
    the flat Adam moments of the fused kernels: {"step", "exp_avg", "exp_avg_sq"} as host tensors over the flat parameter
    buffer.  With the NVLink peer exchange every rank holds only its slice of the moments: state_dict() is then COLLECTIVE
    (one all-reduce assembles the full vectors on every rank) -- call it on all ranks, write the file on one."""

    def __init__(self, owner):
        self._o = owner

    @property
    def param_groups(self):
        o = self._o
        return [{"lr": o.lr, "betas": tuple(o.betas), "eps": 1e-9}]

    def state_dict(self):
        return self._o._adam_state_dict()

    def load_state_dict(self, sd):
        self._o._load_adam_state(sd)


# device / process bound attributes: dropped when an Optim is pickled into a checkpoint (TrainerMultimodal.py:576-587)
_UNPICKLED = ("params", "flat", "gflat", "exp_avg", "exp_avg_sq", "_gsum", "_pws", "_sq", "_ws", "peer", "_early",
              "_owner", "optimizer")


class Optim(object):
    def __init__(self, method, lr, max_grad_norm, lr_decay=1, start_decay_at=None, beta1=0.9, beta2=0.999,
                 adagrad_accum=0.0, decay_method=None, warmup_steps=4000, model_size=None, exchange="auto"):
        assert method in ("adam", "sgd"), "published runs use adam; sgd is kept for completeness"
        self.last_ppl = None
        self.lr, self.original_lr = lr, lr
        self.max_grad_norm = max_grad_norm
        self.method = method
        self.lr_decay, self.start_decay_at, self.start_decay = lr_decay, start_decay_at, False
        self._step = 0
        self.betas = [beta1, beta2]
        self.adagrad_accum = adagrad_accum   # kept for checkpoint compatibility (adagrad is not on the published path)
        self.decay_method, self.warmup_steps, self.model_size = decay_method, warmup_steps, model_size
        self.sync_gradients = True           # all-reduce across ranks when torch.distributed is up
        self.exchange = exchange
        self.peer = None                     # distributed.PeerExchange when the NVLink peer path is active
        self._early = None                   # early (overlapped) exchange of the tail of the flat buffer, see enable_early_exchange
        self._early_done = False
        self._adam_t = 0                     # Adam's bias-correction step (torch.optim.Adam's state['step']; restarts with
        #                                      every set_parameters, exactly as the reference's fresh optim.Adam does)
        self._pending_state = None           # Adam state carried by an unpickled / not yet bound Optim
        self.optimizer = _AdamStateView(self)

    # ---- checkpointing (TrainerMultimodal.py:576-587 pickles the whole Optim; train_mm_vi_model1.py:433-452 resumes) ----
    def __getstate__(self):
        d = {k: v for k, v in self.__dict__.items() if k not in _UNPICKLED}
        d["_pending_state"] = self._adam_state_dict() if getattr(self, "flat", None) is not None else self._pending_state
        d["_early_done"] = False
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.peer, self._early = None, None
        self.optimizer = _AdamStateView(self)

    def _moment_slices(self):
        """[(moment tensor m, moment tensor v, first float, count)] of what this rank holds."""
        n = self.flat.numel()
        if self.peer is not None and self._early is not None:
            e, pe, b0 = self._early, self.peer, self._early["begin"]
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            out = []
            L.lib.vmmt_peer_slice(n - b0, pe.world, pe.rank, ctypes.byref(lo), ctypes.byref(hi))
            out.append((e["m_a"], e["v_a"], b0 + lo.value, hi.value - lo.value))
            L.lib.vmmt_peer_slice(b0, pe.world, pe.rank, ctypes.byref(lo), ctypes.byref(hi))
            out.append((e["m_b"], e["v_b"], lo.value, hi.value - lo.value))
            return out
        if self.peer is not None:
            lo, hi, _cap = self.peer.slice_bounds()
            return [(self.exp_avg, self.exp_avg_sq, lo, hi - lo)]
        return [(self.exp_avg, self.exp_avg_sq, 0, n)]

    def _adam_state_dict(self):
        if getattr(self, "flat", None) is None:
            return self._pending_state
        self.wait_params()
        n = self.flat.numel()
        m = torch.zeros(n, device=self.flat.device, dtype=torch.float32)
        v = torch.zeros_like(m)
        for sm, sv, lo, cnt in self._moment_slices():
            m[lo:lo + cnt].copy_(sm[:cnt])
            v[lo:lo + cnt].copy_(sv[:cnt])
        if self.peer is not None and self.peer.world > 1:
            torch.distributed.all_reduce(m)              # disjoint slices: the sum assembles the full vectors
            torch.distributed.all_reduce(v)
        return {"step": int(self._adam_t), "numel": n, "exp_avg": m.cpu(), "exp_avg_sq": v.cpu(),
                "param_groups": self.optimizer.param_groups}

    def _load_adam_state(self, sd):
        if sd is None:
            return
        if getattr(self, "flat", None) is None:
            self._pending_state = sd
            return
        if int(sd["numel"]) != self.flat.numel():
            raise ValueError("Optim: optimizer state of %d elements does not fit %d parameters" % (sd["numel"], self.flat.numel()))
        dev = self.flat.device
        m, v = sd["exp_avg"].to(dev), sd["exp_avg_sq"].to(dev)
        for sm, sv, lo, cnt in self._moment_slices():
            sm[:cnt].copy_(m[lo:lo + cnt])
            sv[:cnt].copy_(v[lo:lo + cnt])
        self._adam_t = int(sd["step"])

    def set_parameters(self, params, keep_state=False):
        """Bind to the model's parameters (onmt/Optim.py:55-70).  As in the reference, where this builds a NEW
        torch.optim.Adam, the Adam moments and bias-correction step start from zero -- also after ``-train_from``
        (train_mm_vi_model1.py:433-452 loads the state into the old optimiser object and then rebuilds it here), while
        ``_step`` / ``lr`` / decay bookkeeping survive in the pickled Optim.  ``keep_state=True`` restores the moments an
        unpickled Optim carries (or that ``optimizer.load_state_dict`` stored) instead: a true resume."""
        self.params = [p for p in params if p.requires_grad]
        ptrs = {p.data.untyped_storage().data_ptr() for p in self.params}
        grads_ok = all(p.grad is not None for p in self.params) and \
            len({p.grad.untyped_storage().data_ptr() for p in self.params}) == 1
        if len(ptrs) != 1 or not grads_ok:
            holder = _ParamList(self.params)
            flatten_parameters(holder)
            holder._flat_owner_keepalive = holder
        p0 = self.params[0]
        mode = self.exchange
        if mode == "auto":
            mode = os.environ.get("VMMT_DP_EXCHANGE", "auto")
        if mode == "auto":
            mode = "peer" if (distributed.is_active() and p0.is_cuda and self.method == "adam") else "nccl"
        if mode == "peer" and (self.method != "adam" or not p0.is_cuda):
            raise ValueError("exchange='peer' fuses the gradient exchange into the CUDA clip+Adam kernels: it needs "
                             "method='adam' and parameters on a CUDA device")
        if mode == "peer" and self.peer is None:
            # move the flat buffers into the cudaIpc segment every peer maps (before any CUDA-graph capture)
            owner = owner_of(p0)
            assert owner is not None, "peer exchange needs the parameters in one flat buffer"
            self.peer = distributed.PeerExchange.create(padded_numel(owner), p0.device)
            if self.peer is not None:
                flatten_parameters(owner, buffers=(self.peer.flat, self.peer.gflat))
                self._owner = owner
        if self.peer is not None:
            self.flat, self.gflat = self.peer.flat, self.peer.gflat
            _lo, _hi, cap = self.peer.slice_bounds()
            self._gsum = torch.zeros(cap, device=p0.device, dtype=torch.float32)
            self.exp_avg = torch.zeros(cap, device=p0.device, dtype=torch.float32)       # moments of the slice only
            self.exp_avg_sq = torch.zeros(cap, device=p0.device, dtype=torch.float32)
            self._pws = torch.empty(L.lib.vmmt_peer_adam_workspace_bytes() // 4, device=p0.device,
                                    dtype=torch.float32)
            torch.cuda.synchronize(p0.device)
            if distributed.is_active():
                torch.distributed.barrier()          # every rank's segment is mapped and initialised
        else:
            n = p0.data.untyped_storage().nbytes() // 4
            self.flat = torch.empty(0, device=p0.device, dtype=torch.float32).set_(p0.data.untyped_storage(), 0, (n,))
            self.gflat = torch.empty(0, device=p0.device, dtype=torch.float32).set_(p0.grad.untyped_storage(), 0, (n,))
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        self._sq = torch.zeros(1, device=p0.device, dtype=torch.float32)
        self._ws = torch.empty(L.lib.vmmt_sqnorm_workspace_bytes() // 4, device=p0.device, dtype=torch.float32)
        self._adam_t = 0
        self._early, self._early_done = None, False
        if keep_state and self._pending_state is not None:
            self._load_adam_state(self._pending_state)
        self._pending_state = None

    def enable_early_exchange(self, model):
        """Split the peer exchange in two: the gradients of the latent / image networks and the generator (the tail of
        the flat buffer, final before the encoders' backward pass starts) are reduce-scattered on their own stream
        BESIDE the encoder backward; only the remainder is exchanged inside step().  The first encoder-stack backward
        node of a backward pass fires the early phase (NMTVIModel.forward arms it through ``model.early_exchange_hook``).
        Call right after set_parameters (the slice-wise Adam moments are re-partitioned).  Assumes ONE backward pass per
        step(): with gradient accumulation over several backward passes the tail is not final when the first pass
        reaches the encoders -- leave it disabled there.  Returns True when active (False: single rank, NCCL exchange,
        VMMT_DP_OVERLAP=0, or a model whose loss-side modules are not the tail of the flat buffer)."""
        if self.peer is None or os.environ.get("VMMT_DP_OVERLAP", "1") == "0":
            return False
        carried = self._adam_state_dict() if self._adam_t != 0 else None      # re-partitioning keeps restored moments
        begin = distributed.early_final_begin(model)
        n = self.flat.numel()
        if begin is None or begin % 4 or begin <= 0 or begin >= n:
            return False
        pe, dev = self.peer, self.flat.device
        lo, hi = ctypes_i64(), ctypes_i64()
        cap_b = int(L.lib.vmmt_peer_slice(begin, pe.world, pe.rank, lo, hi))
        cap_a = int(L.lib.vmmt_peer_slice(n - begin, pe.world, pe.rank, lo, hi))
        z = lambda k: torch.zeros(max(k, 4), device=dev, dtype=torch.float32)        # noqa: E731
        self._early = {"begin": begin, "gsum_a": z(cap_a), "m_a": z(cap_a), "v_a": z(cap_a), "gsum_b": z(cap_b),
                       "m_b": z(cap_b), "v_b": z(cap_b),
                       "ws_a": torch.empty(L.lib.vmmt_peer_adam_workspace_bytes() // 4, device=dev, dtype=torch.float32),
                       "stream": torch.cuda.Stream(device=dev, priority=-1)}
        self.exp_avg = self.exp_avg_sq = self._gsum = None       # replaced by the per-range slices above
        model.early_exchange_hook = self.early_reduce_scatter
        if os.environ.get("VMMT_DP_AG_OVERLAP", "0") == "1":
            # overlapped all-gather of the tail range: an EXTERNAL event (recorded by step() outside any captured graph,
            # waited for inside the captured forward pass) marks "the tail parameters of the previous update have landed".
            # OFF by default: measured on 2 / 8 B200s (tools/dp_ab.sh, profiles/dp_exchange_r2.txt) it hides the ~130 us tail
            # all-gather but the arriving NVLink traffic holds up the first copies / kernels of the next step by as much
            # (2.166 vs 2.162 ms at 8 GPUs): no net gain, so the simpler in-order schedule stays the default
            ag = {"stream": torch.cuda.Stream(device=dev, priority=-1), "event": torch.cuda.Event(external=True)}
            ag["event"].record(torch.cuda.current_stream(dev))      # created before any capture waits for it
            self._early["ag"] = ag
            model.params_ready_hook = self.wait_params
        if carried is not None:
            self._load_adam_state(carried)
        torch.cuda.synchronize(dev)
        return True

    def early_reduce_scatter(self, after_event=None):
        """Phase 1 of the split exchange (called from the first encoder backward node; no host synchronisation,
        graph-capturable).  ``after_event``: what was on the calling stream when that node started (the node's own
        recurrence kernel, already launched, is NOT waited for)."""
        e, pe = self._early, self.peer
        cur = torch.cuda.current_stream(self.flat.device)
        xs = e["stream"]
        if after_event is not None:
            xs.wait_event(after_event)
        else:
            xs.wait_stream(cur)
        capturing = torch.cuda.is_current_stream_capturing()
        for st in ops.aux_streams():                         # every weight gradient issued so far
            if capturing:                                    # helper streams that are not part of this capture hold nothing
                with torch.cuda.stream(st):                  # of this step (and may not be waited on from inside it)
                    if not torch.cuda.is_current_stream_capturing():
                        continue
            xs.wait_stream(st)
        n = self.flat.numel()
        with torch.cuda.stream(xs):
            L.call("vmmt_peer_reduce_scatter", pe.segments, pe.mc_base, pe.grad_off, pe.rank, pe.world, e["begin"], n - e["begin"],
                   fptr(e["gsum_a"]), 1, fptr(e["ws_a"]), stream())
        self._early_done = True

    def wait_params(self):
        """The current stream waits until the overlapped all-gather of the previous update has delivered every parameter
        (no-op without it).  NMTVIModel.forward calls this before the latent networks; call it yourself before reading the
        parameters from another stream / the host (evaluation, checkpointing) between steps."""
        ag = (self._early or {}).get("ag")
        if ag is not None:
            torch.cuda.current_stream(self.flat.device).wait_event(ag["event"])

    def join_early(self):
        """The current stream waits for the early reduce-scatter (a captured step must join it before the capture ends)."""
        if self._early is not None and self._early_done:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._early["stream"])

    @property
    def exchange_in_use(self):
        how = "nvls multimem.ld_reduce reduce-scatter + sharded clip/Adam + multimem.st all-gather" \
            if (self.peer is not None and self.peer.mc_base) else "nvlink p2p reduce-scatter + sharded clip/Adam + all-gather"
        if self.peer is not None and self._early is not None:
            return (how + " (csrc/peer.cu); tail of the buffer (latent / image networks, generator: %.0f%%) "
                    "reduce-scattered beside the encoder backward%s"
                    % (100.0 * (self.flat.numel() - self._early["begin"]) / self.flat.numel(),
                       " and all-gathered under the next step's encoder phase" if self._early.get("ag") else ""))
        if self.peer is not None:
            return how + " (csrc/peer.cu)"
        return "nccl all-reduce + replicated clip/Adam" if distributed.is_active() else "single rank"

    def _set_rate(self, lr):
        self.lr = lr

    def grad_norm(self):
        """Global L2 norm of the (all-reduced) gradient as a device scalar tensor.  With the peer exchange the summed
        gradient never exists in one place: the value is the norm the last step() clipped with (before any step: the
        norm of this rank's own gradients)."""
        if self.peer is not None and self._adam_t > 0:
            return self._sq.sqrt()
        L.call("vmmt_sqnorm", fptr(self.gflat), self.gflat.numel(), fptr(self._sq), 0, fptr(self._ws), stream())
        return self._sq.sqrt()

    def step(self):
        """clip + update (+ gradient exchange).  Stays OUTSIDE captured CUDA graphs: the learning rate and Adam's bias
        corrections are host scalars of each call."""
        self._step += 1
        self._adam_t += 1
        ops.join_side()
        ops.weights_changed()                # bf16 variant: cached casts of the parameters are stale after this update
        if self.decay_method == "noam":
            self._set_rate(self.original_lr * (self.model_size ** (-0.5) *
                           min(self._step ** (-0.5), self._step * self.warmup_steps ** (-1.5))))
        n = self.flat.numel()
        max_norm = float(self.max_grad_norm) if self.max_grad_norm else 0.0
        if self.peer is not None and self._early is not None:
            e, pe = self._early, self.peer
            if not self._early_done:                         # no backward hook fired (eval forward, custom loop): do it here
                self.early_reduce_scatter()
            torch.cuda.current_stream(self.flat.device).wait_stream(e["stream"])
            self._early_done = False
            b0, sm = e["begin"], stream()
            hyp = (max_norm, float(self.lr), float(self.betas[0]), float(self.betas[1]), 1e-9, self._adam_t)
            L.call("vmmt_peer_reduce_scatter", pe.segments, pe.mc_base, pe.grad_off, pe.rank, pe.world, 0, b0, fptr(e["gsum_b"]), 0,
                   fptr(self._pws), sm)
            # head of the buffer first (embeddings, encoders, decoder: what the next forward pass reads first), bracketed by
            # barriers: norm shares landed / nobody reads my gradients any more -> update + all-gather -> every rank's head
            # parameters are complete
            L.call("vmmt_peer_adam_allgather", pe.segments, pe.mc_base, pe.param_off, pe.rank, pe.world, 0, b0, fptr(e["gsum_b"]),
                   fptr(e["m_b"]), fptr(e["v_b"]), fptr(self._sq), 2, *hyp, 1, 1, 0, sm)
            # the tail (latent / image networks, generator: first read ~250 us into the next forward pass) is updated and
            # all-gathered on its own stream with its own barrier channel, UNDER the next step's encoder phase; the model
            # waits for `ag["event"]` (NMTVIModel.params_ready_hook) before it touches those parameters
            ag = e.get("ag")
            cur = torch.cuda.current_stream(self.flat.device)
            if ag is not None:
                ag["stream"].wait_stream(cur)
                with torch.cuda.stream(ag["stream"]):
                    L.call("vmmt_peer_adam_allgather", pe.segments, pe.mc_base, pe.param_off, pe.rank, pe.world, b0, n - b0,
                           fptr(e["gsum_a"]), fptr(e["m_a"]), fptr(e["v_a"]), None, 2, *hyp, 0, 1, 1, stream())
                    ag["event"].record(ag["stream"])
            else:
                L.call("vmmt_peer_adam_allgather", pe.segments, pe.mc_base, pe.param_off, pe.rank, pe.world, b0, n - b0,
                       fptr(e["gsum_a"]), fptr(e["m_a"]), fptr(e["v_a"]), None, 2, *hyp, 0, 1, 0, sm)
            return
        if self.peer is not None:
            pe = self.peer
            L.call("vmmt_peer_adam_step", pe.segments, pe.mc_base, pe.param_off, pe.grad_off, pe.rank, pe.world, n,
                   fptr(self._gsum), fptr(self.exp_avg), fptr(self.exp_avg_sq), fptr(self._sq), max_norm,
                   float(self.lr), float(self.betas[0]), float(self.betas[1]), 1e-9, self._adam_t, fptr(self._pws),
                   stream())
            return
        if self.sync_gradients:
            distributed.all_reduce_gradients(self.gflat)
        if max_norm > 0:
            L.call("vmmt_sqnorm", fptr(self.gflat), n, fptr(self._sq), 0, fptr(self._ws), stream())
        if self.method == "adam":
            L.call("vmmt_adam_clip_step", fptr(self.flat), fptr(self.gflat), fptr(self.exp_avg),
                   fptr(self.exp_avg_sq), n, fptr(self._sq), max_norm, 1.0, float(self.lr),
                   float(self.betas[0]), float(self.betas[1]), 1e-9, self._adam_t, stream())
        else:
            coef = 1.0
            if max_norm > 0:                                  # clip_grad_norm semantics (onmt/Optim.py:94-95); sgd is not the
                norm = float(self._sq) ** 0.5                 # published method: the host read of the norm is accepted here
                coef = min(1.0, max_norm / (norm + 1e-6))
            L.call("vmmt_axpy", fptr(self.flat), fptr(self.gflat), -float(self.lr) * coef, n, stream())

    def update_learning_rate(self, ppl, epoch):
        if self.start_decay_at is not None and epoch >= self.start_decay_at:
            self.start_decay = True
        if self.last_ppl is not None and ppl > self.last_ppl:
            self.start_decay = True
        if self.start_decay:
            self.lr = self.lr * self.lr_decay
            print("Decaying learning rate to %g" % self.lr)
        self.last_ppl = ppl
