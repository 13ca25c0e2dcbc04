"""Model construction (reference: onmt/ModelConstructor.py:28-60 make_embeddings, 63-111
make_encoder, 328-620 make_vi_model_mmt).  Builds the same module tree with the same parameter
names, initialises every parameter uniform(-param_init, param_init) or loads a reference checkpoint
(keys 'model' / 'generator'), attaches the generator and lays all parameters out in one flat buffer.
"""
import torch
import torch.nn as nn

from . import ops
from .Models import NMTVIModel, RNNEncoder
from .VI_Model1 import StdRNNVIModel1Decoder
from .modules import (Embeddings, GlobalInferenceNetwork, GlobalFullInferenceNetwork,
                      ImageGlobalInferenceNetwork)

PAD_WORD = "<blank>"


class _GenLinear(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        k = 1.0 / in_features ** 0.5
        self.weight = nn.Parameter(torch.empty(out_features, in_features).uniform_(-k, k))
        self.bias = nn.Parameter(torch.empty(out_features).uniform_(-k, k))


class Generator(nn.Sequential):
    """nn.Sequential(nn.Linear(rnn_size, V), nn.LogSoftmax()) -- state_dict keys 0.weight / 0.bias.
    ``forward`` materialises log-probabilities (decode); training goes through the fused loss."""

    def __init__(self, rnn_size, vocab_size):
        super().__init__(_GenLinear(rnn_size, vocab_size))

    def forward(self, x):
        return ops.generator_logprobs(x.reshape(-1, x.size(-1)), self[0].weight, self[0].bias)


def make_embeddings(opt, word_dict, feature_dicts=(), for_encoder=True):
    dim = opt.src_word_vec_size if for_encoder else opt.tgt_word_vec_size
    assert len(feature_dicts) == 0
    return Embeddings(word_vec_size=dim, word_vocab_size=len(word_dict),
                      word_padding_idx=word_dict.stoi[PAD_WORD])


def make_encoder(opt, embeddings):
    assert opt.encoder_type in ("rnn", "brnn"), "the VI-model-1 path uses the RNN encoder"
    return RNNEncoder(opt.rnn_type, opt.encoder_type == "brnn", opt.enc_layers, opt.rnn_size,
                      opt.dropout, embeddings)


def make_vi_model_mmt(model_opt, fields, gpu, checkpoint=None):
    assert model_opt.model_type == "text"
    assert not getattr(model_opt, "use_posterior_image_features", False)
    feat_size = 4096 if "vgg" in model_opt.path_to_train_img_feats.lower() else 2048
    model_opt.global_image_features_dim = feat_size
    src_dict, tgt_dict = fields["src"].vocab, fields["tgt"].vocab
    src_embeddings = make_embeddings(model_opt, src_dict)
    encoder = make_encoder(model_opt, src_embeddings)
    tgt_embeddings = make_embeddings(model_opt, tgt_dict, for_encoder=False)
    if getattr(model_opt, "share_embeddings", False):
        tgt_embeddings.word_lut.weight = src_embeddings.word_lut.weight
    assert model_opt.use_global_image_features, "global image features are the published configuration"
    brnn = model_opt.encoder_type == "brnn"
    decoder = StdRNNVIModel1Decoder(model_opt.rnn_type, brnn, model_opt.dec_layers, model_opt.rnn_size,
                                    model_opt.global_attention, model_opt.coverage_attn,
                                    model_opt.context_gate, model_opt.copy_attn, model_opt.dropout,
                                    model_opt.word_dropout, tgt_embeddings, model_opt.z_latent_dim,
                                    model_opt.reuse_copy_attn)
    if model_opt.conditional:
        input_dims = 2 * model_opt.rnn_size + feat_size
        inf_net_global = GlobalFullInferenceNetwork(model_opt.z_latent_dim, input_dims, "normal",
                                                    image_features_type="global")
        gen_net_global = GlobalInferenceNetwork(model_opt.z_latent_dim, model_opt.rnn_size, "normal")
        encoder_tgt = RNNEncoder(model_opt.rnn_type, True, model_opt.enc_layers, model_opt.rnn_size,
                                 model_opt.dropout, tgt_embeddings)
    else:
        inf_net_global = GlobalInferenceNetwork(model_opt.z_latent_dim, model_opt.rnn_size, "normal")
        gen_net_global, encoder_tgt = None, None
    inf_net_image = ImageGlobalInferenceNetwork(model_opt.z_latent_dim, feat_size, model_opt.rnn_size,
                                                False, "normal")
    model = NMTVIModel(encoder, decoder, encoder_inference=None, inf_net_global=inf_net_global,
                       gen_net_global=gen_net_global, inf_net_image=inf_net_image,
                       multimodal_model_type="vi-model1", image_loss_type=model_opt.image_loss,
                       image_features_type="global", image_features_projector=None,
                       two_step_image_prediction=False, conditional=model_opt.conditional,
                       encoder_tgt=encoder_tgt)
    model.model_type = model_opt.model_type
    assert not model_opt.copy_attn
    generator = Generator(model_opt.rnn_size, len(tgt_dict))
    if getattr(model_opt, "share_decoder_embeddings", False):
        generator[0].weight = decoder.embeddings.word_lut.weight
    if checkpoint is not None:
        model.load_state_dict(checkpoint["model"])
        generator.load_state_dict(checkpoint["generator"])
    elif model_opt.param_init != 0.0:
        with torch.no_grad():
            for p in list(model.parameters()) + list(generator.parameters()):
                p.uniform_(-model_opt.param_init, model_opt.param_init)
    model.generator = generator
    if gpu:
        model.cuda()
    model.flatten_parameters()
    return model
