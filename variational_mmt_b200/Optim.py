"""Optimiser wrapper (reference: onmt/Optim.py:5-114): same constructor / set_parameters / step /
update_learning_rate contract; 'adam' (the published method) runs as a fused global-norm clip +
Adam(eps=1e-9) over the flat buffers in two launches, with the data-parallel gradient all-reduce
(NCCL, SUM -- SURVEY.md section 8e) issued on the same flat buffer right before it.
"""
import torch

from . import _lib as L
from ._lib import fptr, stream
from .flat import flatten_parameters
from . import distributed, ops


class _ParamList(torch.nn.Module):
    def __init__(self, params):
        super().__init__()
        self.plist = torch.nn.ParameterList(params)


class Optim(object):
    def __init__(self, method, lr, max_grad_norm, lr_decay=1, start_decay_at=None, beta1=0.9, beta2=0.999,
                 adagrad_accum=0.0, decay_method=None, warmup_steps=4000, model_size=None):
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

    def set_parameters(self, params):
        self.params = [p for p in params if p.requires_grad]
        ptrs = {p.data.untyped_storage().data_ptr() for p in self.params}
        grads_ok = all(p.grad is not None for p in self.params) and \
            len({p.grad.untyped_storage().data_ptr() for p in self.params}) == 1
        if len(ptrs) != 1 or not grads_ok:
            holder = _ParamList(self.params)
            flatten_parameters(holder)
        p0 = self.params[0]
        n = p0.data.untyped_storage().nbytes() // 4
        self.flat = torch.empty(0, device=p0.device, dtype=torch.float32).set_(p0.data.untyped_storage(), 0, (n,))
        self.gflat = torch.empty(0, device=p0.device, dtype=torch.float32).set_(p0.grad.untyped_storage(), 0, (n,))
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self._sq = torch.zeros(1, device=p0.device, dtype=torch.float32)
        self._ws = torch.empty(L.lib.vmmt_sqnorm_workspace_bytes() // 4, device=p0.device, dtype=torch.float32)

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
        if self.sync_gradients:
            distributed.all_reduce_gradients(self.gflat)
        n = self.flat.numel()
        max_norm = float(self.max_grad_norm) if self.max_grad_norm else 0.0
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
