"""Luong global attention, 'general' scoring (reference: onmt/modules/GlobalAttention.py:60-217).

Same constructor, parameters (``linear_in.weight`` [dim,dim], ``linear_out.weight`` [dim,2*dim], no
biases) and forward contract (3-D input = sequence mode, 2-D input = one-step mode).  The
[c ; q] concatenation before linear_out (GlobalAttention.py:187) is never materialised: linear_out
is applied as two accumulating GEMMs over the column halves of its weight with tanh in the epilogue
of the second.  'mlp' scoring and coverage are outside the hot path.
"""
import torch
import torch.nn as nn

from .. import ops


class _NoBiasLinear(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        k = 1.0 / in_features ** 0.5
        self.weight = nn.Parameter(torch.empty(out_features, in_features).uniform_(-k, k))


class GlobalAttention(nn.Module):
    def __init__(self, dim, coverage=False, attn_type="dot"):
        super().__init__()
        self.dim, self.attn_type = dim, attn_type
        assert attn_type in ("dot", "general"), "only 'dot' and 'general' attention run on this path"
        assert not coverage, "coverage attention is not supported (the reference asserts the same)"
        if attn_type == "general":
            self.linear_in = _NoBiasLinear(dim, dim)
        self.linear_out = _NoBiasLinear(dim * 2, dim)

    def forward_time_major(self, q, context, context_lengths=None):
        """q [T,B,dim], context [S,B,dim] (both time-major) -> attn_h [T,B,dim], align [T,B,S]."""
        dim = self.dim
        # (the join node is created BEFORE linear_in: autograd runs ready nodes in reverse creation order, so in backward
        # linear_in's input-gradient GEMM is issued before the join makes the stream wait for the context-side kernel)
        context, token = ops.attention_context(context)
        qp = ops.linear(q, self.linear_in.weight) if self.attn_type == "general" else q
        cvec, align = ops.AttentionCoreFn.apply(qp, context, context_lengths, token)
        attn_h = ops.dual_linear(cvec, q, self.linear_out.weight, act=ops.ACT_TANH)     # tanh(linear_out([c ; q]))
        return attn_h, align

    def forward(self, input, context, context_lengths=None, coverage=None):
        """Reference layout: input [batch, tgt_len, dim] (or [batch, dim] for one step), context
        [batch, src_len, dim] -> attn_h [tgt_len, batch, dim], align [tgt_len, batch, src_len]
        (one step: [batch, dim], [batch, src_len])."""
        assert coverage is None
        one_step = input.dim() == 2
        if one_step:
            input = input.unsqueeze(1)
        q = input.transpose(0, 1).contiguous()
        ctx = context.transpose(0, 1).contiguous()
        attn_h, align = self.forward_time_major(q, ctx, context_lengths)
        if one_step:
            return attn_h.squeeze(0), align.squeeze(0)
        return attn_h, align
