"""Inference / generative networks of VI model 1
(reference: onmt/modules/NormalVariationalEncoder.py:12-43, 47-110, 113-228, 231-304).

Same classes, constructor arguments, parameter names and return values ``(Normal, h)``.
"""
import torch
import torch.nn as nn

from .. import ops
from .Dists import Normal


class _Linear(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        k = 1.0 / in_features ** 0.5
        self.weight = nn.Parameter(torch.empty(out_features, in_features).uniform_(-k, k))
        self.bias = nn.Parameter(torch.empty(out_features).uniform_(-k, k))


class LocationLayer(nn.Module):
    """fc2(relu(fc1 x))"""

    def __init__(self, input_size, hidden_size, output_size=None):
        super().__init__()
        self.fc1 = _Linear(input_size, hidden_size)
        self.fc2 = _Linear(hidden_size, output_size if output_size is not None else hidden_size)

    def forward(self, x):
        h = ops.linear(x, self.fc1.weight, self.fc1.bias, ops.ACT_RELU)
        return ops.linear(h, self.fc2.weight, self.fc2.bias)


class ScaleLayer(nn.Module):
    """softplus(fc2(relu(fc1 x)))"""

    def __init__(self, input_size, hidden_size, output_size=None):
        super().__init__()
        self.fc1 = _Linear(input_size, hidden_size)
        self.fc2 = _Linear(hidden_size, output_size if output_size is not None else hidden_size)

    def forward(self, x):
        h = ops.linear(x, self.fc1.weight, self.fc1.bias, ops.ACT_RELU)
        return ops.linear(h, self.fc2.weight, self.fc2.bias, ops.ACT_SOFTPLUS)


def _mlp_heads(location, scale=None):
    h = [(location.fc1.weight, location.fc1.bias, location.fc2.weight, location.fc2.bias)]
    if scale is not None:
        h.append((scale.fc1.weight, scale.fc1.bias, scale.fc2.weight, scale.fc2.bias))
    return h


class _LazyCat(object):
    """The concatenated network input ``h`` the reference returns beside the distribution (never read by its callers,
    onmt/Models.py:889-930): the row-MLP kernel consumes the parts in place, ``h()`` materialises the concatenation."""

    def __init__(self, parts):
        self.parts = parts

    def __call__(self):
        return torch.cat(list(self.parts), -1)


class GlobalInferenceNetwork(nn.Module):
    """Z | x ~ N(loc(h), scale(h)), h = masked mean of the source encodings."""

    def __init__(self, z_dim, input_size, dist_type):
        assert dist_type == "normal", "only the Normal family is instantiated by the reference configs"
        super().__init__()
        self.location = LocationLayer(input_size, z_dim)
        self.scale = ScaleLayer(input_size, z_dim)
        self.dist_type = dist_type

    def encode_seq(self, seq, seq_lengths):
        return ops.MaskedMeanFn.apply(seq, seq_lengths)

    def forward(self, x, x_lengths):
        h = self.encode_seq(x, x_lengths)
        # both MLPs in exact fp32 over the batch rows: two launches (csrc/rowlin.cu)
        loc, scale = ops.row_mlp([h], _mlp_heads(self.location, self.scale), (ops.ACT_NONE, ops.ACT_SOFTPLUS))
        return Normal(loc, scale), h


class GlobalFullInferenceNetwork(GlobalInferenceNetwork):
    """Z | x, y, v ~ N(loc(h), scale(h)), h = [mean(x); mean(y); v] (global image features)."""

    def __init__(self, z_dim, input_size, dist_type, image_features_type="global"):
        assert image_features_type in ("global", "posterior"), \
            "local image features are outside the hot path (SURVEY.md section 2, row 5)"
        super().__init__(z_dim, input_size, dist_type)
        self.image_features_type = image_features_type

    def forward(self, x, x_lengths, y, y_lengths, v):
        hx = self.encode_seq(x, x_lengths)
        hy = self.encode_seq(y, y_lengths)
        # q(z|x,y,v) sits on the step's critical path (target encoder -> here -> decoder): h = [hx ; hy ; v]
        # (Models.py:911) is consumed in place by the row-MLP kernel, location and scale heads share each launch
        loc, scale = ops.row_mlp([hx, hy, v], _mlp_heads(self.location, self.scale), (ops.ACT_NONE, ops.ACT_SOFTPLUS))
        return Normal(loc, scale), _LazyCat((hx, hy, v))


class ImageGlobalInferenceNetwork(GlobalInferenceNetwork):
    """V | z ~ N(loc(z * g), scale(z * g)), g = sigmoid(affine(z)); use_source_encodings=False is the
    only configuration the reference constructs (ModelConstructor.py:536-539)."""

    def __init__(self, latent_dim, image_feats_dim, src_encodings_dim, use_source_encodings, dist_type):
        assert not use_source_encodings, "use_source_encodings=True is never constructed by the reference"
        super().__init__(image_feats_dim, latent_dim, dist_type)
        self.use_source_encodings = use_source_encodings
        self.gate_affine_transform = _Linear(latent_dim, 1)

    def forward(self, z, x=None, x_lengths=None):
        gated = ops.gate(z, self.gate_affine_transform.weight, self.gate_affine_transform.bias)
        (loc,) = ops.row_mlp([gated], _mlp_heads(self.location), (ops.ACT_NONE,))
        # the scale branch never reaches the loss (VILoss.py:321): evaluate it only on demand
        return Normal(loc, lambda: self.scale(gated.detach()).detach()), None
