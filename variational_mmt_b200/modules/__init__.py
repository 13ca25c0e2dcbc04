"""Building blocks of the VI-model-1 hot path, mirroring ``onmt/modules`` of the reference
(same class names, constructor arguments, forward signatures and state_dict keys), with all
arithmetic in libvmmt kernels."""
from .Embeddings import Embeddings
from .LSTM import LSTM
from .GlobalAttention import GlobalAttention
from .Dists import Normal
from .NormalVariationalEncoder import (LocationLayer, ScaleLayer, GlobalInferenceNetwork,
                                       GlobalFullInferenceNetwork, ImageGlobalInferenceNetwork)

__all__ = ["Embeddings", "LSTM", "GlobalAttention", "Normal", "LocationLayer", "ScaleLayer",
           "GlobalInferenceNetwork", "GlobalFullInferenceNetwork", "ImageGlobalInferenceNetwork"]
