"""Multi-layer (bi)LSTM with torch.nn.LSTM's parameter names, running on the persistent
recurrence kernels (reference call sites: onmt/Models.py:124-129,139-147; onmt/VI_Model1.py:106,149-152).

Differences from nn.LSTM that the callers in this package rely on:
  * ``lengths`` replaces pack_padded_sequence / pad_packed_sequence: rows are frozen past their
    length and their outputs are zero (exactly what unpacking produces);
  * ``in_bias`` ([N, 4H]) is a per-example additive gate term for layer 0 and ``in_cols`` selects
    the columns of ``weight_ih_l0`` that multiply ``input`` -- the decoder uses both so that the
    [T,B,E+Z] concatenation of embeddings and z (VI_Model1.py:99-100) is never materialised.
"""
import math

import torch
import torch.nn as nn

from .. import ops


class LSTM(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers=1, dropout=0.0, bidirectional=False):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.dropout, self.bidirectional = float(dropout), bidirectional
        self.cluster_budget = None       # cap on clusters per recurrence launch (set when two stacks run concurrently)
        self.fires_early_exchange = False   # encoder stacks: their backward marks the point where the loss-side gradients are final
        ndir = 2 if bidirectional else 1
        k = 1.0 / math.sqrt(hidden_size)
        for l in range(num_layers):
            for sfx in ("", "_reverse")[:ndir]:
                i = input_size if l == 0 else hidden_size * ndir
                for name, shape in ((f"weight_ih_l{l}{sfx}", (4 * hidden_size, i)),
                                    (f"weight_hh_l{l}{sfx}", (4 * hidden_size, hidden_size)),
                                    (f"bias_ih_l{l}{sfx}", (4 * hidden_size,)),
                                    (f"bias_hh_l{l}{sfx}", (4 * hidden_size,))):
                    self.register_parameter(name, nn.Parameter(torch.empty(*shape).uniform_(-k, k)))

    def _weights(self, l):
        out = []
        for sfx in ("", "_reverse")[: 2 if self.bidirectional else 1]:
            out += [getattr(self, f"{n}_l{l}{sfx}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        return out

    def forward(self, input, hx=None, lengths=None, in_bias=None, in_cols=None, gx_given=False):
        """input [T,N,In]; hx = (h0, c0) each [layers*ndir, N, H] or None ->
        output [T,N,ndir*H], (h_n, c_n) each [layers*ndir, N, H].  ``gx_given``: ``input`` already is layer 0's
        input projection x W_ih[:, in_cols]^T [T,N,4H] (computed ahead of time on another stream)."""
        ndir = 2 if self.bidirectional else 1
        save = torch.is_grad_enabled()
        x = input
        hs, cs = [], []
        for l in range(self.num_layers):
            if l > 0:
                x = ops.dropout(x, self.dropout, self.training)
            h0 = hx[0][l * ndir:(l + 1) * ndir] if hx is not None else None
            c0 = hx[1][l * ndir:(l + 1) * ndir] if hx is not None else None
            cfg = {"save": save, "in_cols": in_cols if l == 0 else None, "gx_given": gx_given and l == 0,
                   "cluster_budget": self.cluster_budget if self.training else None,
                   "cluster_budget_bwd": getattr(self, "cluster_budget_bwd", None) if self.training else None,
                   # weight gradients of a bidirectional stack (the target encoder: lanes 2 + direction) and of a
                   # unidirectional one (source encoder / decoder: lane = layer parity) go to different side streams: the
                   # two encoders' backward passes run concurrently and their last weight-gradient blocks are the tail of
                   # the step
                   "side_lane": 2 if self.bidirectional else l % 2,
                   "fires_early_exchange": self.fires_early_exchange}
            x, hT, cT = ops.lstm_layer(x, h0, c0, in_bias if l == 0 else None, lengths, cfg, self._weights(l))
            hs.append(hT)
            cs.append(cT)
        return x, (torch.cat(hs, 0), torch.cat(cs, 0))
