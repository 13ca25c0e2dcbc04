"""variational_mmt_b200 -- B200-native (sm_100a) implementation of the training / decoding hot path
of iacercalixto/variational_mmt's conditional-VAE multimodal translator (VI model 1).

The modules mirror the reference's operator surface (onmt/Models.py, onmt/VI_Model1.py,
onmt/VILoss.py, onmt/modules/*, onmt/Optim.py, onmt/translate/*): same class names, signatures and
state_dict keys.  All arithmetic runs in hand-written CUDA kernels behind the C ABI in
include/vmmt.h (libvmmt.so); there is no CPU or PyTorch-op fallback.
"""
from . import _lib, ops                                           # noqa: F401  (fails loudly if libvmmt.so is missing)
from .Models import NMTVIModel, RNNEncoder, RNNDecoderState, RNNVIDecoderBase
from .VI_Model1 import StdRNNVIModel1Decoder
from .VILoss import NMTVIModel1LossCompute, VIStatistics
from .Optim import Optim
from .ModelConstructor import make_vi_model_mmt, Generator
from .modules import (Embeddings, LSTM, GlobalAttention, Normal, LocationLayer, ScaleLayer,
                      GlobalInferenceNetwork, GlobalFullInferenceNetwork, ImageGlobalInferenceNetwork)
from .ops import manual_seed, set_gemm_mode, get_gemm_mode
from .graph import GraphedTrainStep
from . import translate, distributed, io
from .translate import TranslatorMultimodalVI, GNMTGlobalScorer

__version__ = "0.1.0"
