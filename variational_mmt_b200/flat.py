"""One contiguous fp32 buffer for all parameters and one for all gradients.

The backward kernels accumulate straight into ``param.grad`` views of the flat gradient buffer, so
``zero_grad`` is one memset, the data-parallel exchange runs over the flat buffers (peer-memory kernels, or one NCCL
all-reduce as the fallback) and clip+Adam is two
launches (see Optim.py).  Parameter identity, names and shapes are untouched, so state_dicts saved
by the reference load unchanged.
"""
import weakref

import torch

_OWNERS = {}          # data_ptr of a flat parameter buffer -> weakref of the module that owns it


def owner_of(param):
    """The module whose flat buffer holds ``param`` (None when it is not in one)."""
    ref = _OWNERS.get(param.data.untyped_storage().data_ptr())
    return ref() if ref is not None else None


def padded_numel(module):
    """Length (floats) of the flat buffers ``flatten_parameters(module)`` needs."""
    seen, total = set(), 0
    for p in module.parameters():
        if id(p) not in seen:
            seen.add(id(p))
            total += ((p.numel() + 3) // 4) * 4
    return total


def flatten_parameters(module, buffers=None):
    """Re-point every parameter of ``module`` (and its ``.grad``) into flat buffers.
    Idempotent; call again after moving the module to another device.  ``buffers=(flat, gflat)`` places
    them in caller-provided fp32 storage (the NVLink peer segment of distributed.PeerExchange)."""
    params = []
    seen = set()
    for p in module.parameters():
        if id(p) not in seen:
            seen.add(id(p))
            params.append(p)
    if not params:
        return None, None
    dev = params[0].device
    sizes = [((p.numel() + 3) // 4) * 4 for p in params]          # keep every tensor 16-byte aligned
    total = sum(sizes)
    if buffers is None:
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        gflat = torch.zeros(total, device=dev, dtype=torch.float32)
    else:
        flat, gflat = buffers
        assert flat.numel() == total and gflat.numel() == total and flat.device == dev, "flat buffers: wrong size/device"
        gflat.zero_()
    off = 0
    with torch.no_grad():
        for p, n in zip(params, sizes):
            view = flat[off: off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = gflat[off: off + p.numel()].view(p.shape)
            off += n
    module._flat_params, module._flat_grads = flat, gflat
    _OWNERS[flat.untyped_storage().data_ptr()] = weakref.ref(module)
    return flat, gflat


class FlatParamsMixin:
    """nn.Module mixin: flat buffers + a zero_grad that keeps the gradient views alive."""

    _flat_params = None
    _flat_grads = None

    def flatten_parameters(self):
        return flatten_parameters(self)

    def zero_grad(self, set_to_none=False, max_blocks=None):
        """``max_blocks``: zero the flat gradient buffer with that many resident blocks (vmmt_fill_zero) instead of a
        full-width memset -- for callers that overlap it with latency-critical work (GraphedTrainStep)."""
        if self._flat_grads is not None and not set_to_none:
            if max_blocks is not None and self._flat_grads.is_cuda:
                from . import _lib as L
                L.call("vmmt_fill_zero", L.fptr(self._flat_grads), self._flat_grads.numel(), int(max_blocks), L.stream())
            else:
                self._flat_grads.zero_()
            return
        if self._flat_grads is not None:
            raise RuntimeError("zero_grad(set_to_none=True) would detach the flat gradient views")
        return super().zero_grad(set_to_none=set_to_none)

    def _apply(self, fn, *a, **k):
        had = self._flat_params is not None
        out = super()._apply(fn, *a, **k)
        if had:
            flatten_parameters(self)
        return out
