"""Normal (reference: onmt/modules/Dists.py:11-60): params() -> [mean, std], mean(), sample().

``sample()`` is mean + std * eps with NO pathwise gradient, which is what the reference's
torch.normal call amounts to (hazard H2).  ``eps`` can be injected (``Normal.inject_noise``) so that
parity tests do not depend on an RNG stream; otherwise it is drawn by the in-kernel Philox generator.
The scale may be lazy (a thunk): the image head's scale branch is dead compute in the loss
(VILoss.py:321) and is only evaluated if somebody asks for it.
"""
import contextlib

import torch

from .. import ops


class Normal(object):
    _injected_eps = None

    def __init__(self, mean, std):
        self._mean, self._std = mean, std

    def mean(self):
        return self._mean

    @property
    def std(self):
        if callable(self._std):
            self._std = self._std()
        return self._std

    def params(self):
        return [self._mean, self.std]

    def sample(self):
        return ops.normal_sample(self._mean, self.std, Normal._injected_eps)

    def log_prob(self, value):
        raise NotImplementedError("not on the hot path: the loss kernels evaluate log-probs directly")

    @staticmethod
    @contextlib.contextmanager
    def inject_noise(eps):
        Normal._injected_eps = eps
        try:
            yield
        finally:
            Normal._injected_eps = None
