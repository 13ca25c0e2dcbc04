"""Word embeddings (reference: onmt/modules/Embeddings.py:12-188, lookup path only).

state_dict key: ``make_embedding.emb_luts.0.weight``.  Feature embeddings, positional encodings and
the 'mlp' feature merge of the reference are outside the VI-model-1 hot path.
"""
import torch
import torch.nn as nn

from .. import ops


class _Lut(nn.Module):
    """nn.Embedding(vocab, dim, padding_idx) parameter holder (weight only)."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx):
        super().__init__()
        self.num_embeddings, self.embedding_dim, self.padding_idx = num_embeddings, embedding_dim, padding_idx
        self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim))
        with torch.no_grad():
            self.weight.normal_(0, 1)
            self.weight[padding_idx].fill_(0)

    def forward(self, idx):
        return ops.embedding(idx, self.weight, self.padding_idx)


class Embeddings(nn.Module):
    def __init__(self, word_vec_size, word_vocab_size, word_padding_idx, position_encoding=False,
                 feat_merge="concat", feat_vec_exponent=0.7, feat_vec_size=-1, feat_padding_idx=(),
                 feat_vocab_sizes=(), dropout=0):
        super().__init__()
        if position_encoding or len(feat_vocab_sizes) > 0:
            raise NotImplementedError("positional encodings / feature embeddings are outside the "
                                      "VI-model-1 hot path")
        self.word_padding_idx = word_padding_idx
        self.embedding_size = word_vec_size
        self.make_embedding = nn.Sequential()
        self.make_embedding.add_module("emb_luts", nn.ModuleList(
            [_Lut(word_vocab_size, word_vec_size, word_padding_idx)]))

    @property
    def word_lut(self):
        return self.make_embedding[0][0]

    @property
    def emb_luts(self):
        return self.make_embedding[0]

    def load_pretrained_vectors(self, emb_file, fixed):
        if emb_file:
            pretrained = torch.load(emb_file)
            self.word_lut.weight.data.copy_(pretrained)
            if fixed:
                self.word_lut.weight.requires_grad = False

    def forward(self, input):
        """input: LongTensor [len, batch, nfeat=1] -> [len, batch, embedding_size]"""
        assert input.dim() == 3 and input.size(2) == 1, "expected [len x batch x 1] token ids"
        return self.word_lut(input[:, :, 0])
