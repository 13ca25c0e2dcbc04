"""Optimiser wrapper (reference: onmt/Optim.py:5-114): same constructor / set_parameters / step /
update_learning_rate contract; 'adam' (the published method) runs as a fused global-norm clip +
Adam(eps=1e-9) over the flat buffers in two launches.  Data parallel:
on one NVSwitch box the flat buffers live in NVLink peer memory and the step is reduce-scatter -> clip +
Adam on this rank's 1/N slice -> all-gather inside the update kernels (distributed.PeerExchange,
csrc/peer.cu); otherwise one NCCL SUM all-reduce of the flat gradient buffer precedes the update.
``exchange``: "auto" (peer when torch.distributed is up on CUDA, else none), "peer", "nccl"; the
environment variable VMMT_DP_EXCHANGE overrides "auto".
"""
import os

import torch

from . import _lib as L
from ._lib import fptr, stream
from .flat import flatten_parameters, owner_of, padded_numel
from . import distributed, ops


class _ParamList(torch.nn.Module):
    def __init__(self, params):
        super().__init__()
        self.plist = torch.nn.ParameterList(params)


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
        self.decay_method, self.warmup_steps, self.model_size = decay_method, warmup_steps, model_size
        self.sync_gradients = True           # all-reduce across ranks when torch.distributed is up
        self.exchange = exchange
        self.peer = None                     # distributed.PeerExchange when the NVLink peer path is active

    def set_parameters(self, params):
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

    @property
    def exchange_in_use(self):
        if self.peer is not None:
            return "nvlink p2p reduce-scatter + sharded clip/Adam + all-gather (csrc/peer.cu)"
        return "nccl all-reduce + replicated clip/Adam" if distributed.is_active() else "single rank"

    def _set_rate(self, lr):
        self.lr = lr

    def grad_norm(self):
        """Global L2 norm of the (all-reduced) gradient as a device scalar tensor."""
        L.call("vmmt_sqnorm", fptr(self.gflat), self.gflat.numel(), fptr(self._sq), 0, fptr(self._ws), stream())
        return self._sq.sqrt()

    def step(self):
        self._step += 1
        ops.join_side()
        if self.decay_method == "noam":
            self._set_rate(self.original_lr * (self.model_size ** (-0.5) *
                           min(self._step ** (-0.5), self._step * self.warmup_steps ** (-1.5))))
        n = self.flat.numel()
        max_norm = float(self.max_grad_norm) if self.max_grad_norm else 0.0
        if self.peer is not None:
            pe = self.peer
            L.call("vmmt_peer_adam_step", pe.segments, pe.param_off, pe.grad_off, pe.rank, pe.world, n,
                   fptr(self._gsum), fptr(self.exp_avg), fptr(self.exp_avg_sq), fptr(self._sq), max_norm,
                   float(self.lr), float(self.betas[0]), float(self.betas[1]), 1e-9, self._step, fptr(self._pws),
                   stream())
            return
        if self.sync_gradients:
            distributed.all_reduce_gradients(self.gflat)
        if max_norm > 0:
            L.call("vmmt_sqnorm", fptr(self.gflat), n, fptr(self._sq), 0, fptr(self._ws), stream())
        if self.method == "adam":
            L.call("vmmt_adam_clip_step", fptr(self.flat), fptr(self.gflat), fptr(self.exp_avg),
                   fptr(self.exp_avg_sq), n, fptr(self._sq), max_norm, 1.0, float(self.lr),
                   float(self.betas[0]), float(self.betas[1]), 1e-9, self._step, stream())
        else:
            if max_norm > 0:
                raise NotImplementedError("sgd with clipping is not on the published path")
            L.call("vmmt_axpy", fptr(self.flat), fptr(self.gflat), -float(self.lr), n, stream())

    def update_learning_rate(self, ppl, epoch):
        if self.start_decay_at is not None and epoch >= self.start_decay_at:
            self.start_decay = True
        if self.last_ppl is not None and ppl > self.last_ppl:
            self.start_decay = True
        if self.start_decay:
            self.lr = self.lr * self.lr_decay
            print("Decaying learning rate to %g" % self.lr)
        self.last_ppl = ppl
