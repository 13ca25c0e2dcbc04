"""In-process translation of a validation set on the training GPU.

The reference's model selection (onmt/TrainerMultimodal.py:372-394 -> onmt/EarlyStop.py:91-145, 245-272) drops a
temporary checkpoint every ``evaluate_every_nupdates`` updates, forks ``python translate_mm_vi.py -model <ckpt> -src
<valid> -beam_size k -output <tmp>`` (a fresh interpreter that reloads the 170 MB checkpoint and decodes the 1 014
validation sentences ONE AT A TIME, translate_mm_vi.py:80-82), then scores the file with perl / java subprocesses.
Here the live model is decoded in place, all sentences of a batch together (batched device beam search,
TranslatorMultimodalVI.translate_batch), and the hypotheses are returned in corpus order -- or written one sentence
per line, the file format ``EarlyStop.compute_bleus`` / ``compute_meteors`` read.  BLEU / METEOR scoring itself stays
external (multi-bleu.perl / the METEOR jar are not part of the hot path).
"""
import torch

from .Beam import GNMTGlobalScorer
from .TranslatorMultimodalVI import TranslatorMultimodalVI, EOS_WORD


def translate_dataset(model, fields, dataset, batch_size=128, beam_size=1, max_length=100, output=None,
                      translator=None):
    """-> list (corpus order) of token-string lists.  ``dataset``: variational_mmt_b200.io.TripletDataset (only the
    source side is read).  ``beam_size`` 1 = greedy.  The model's train / eval mode is restored on return.
    Pass a ``translator`` to keep its CUDA-graph buckets alive between validation rounds."""
    import numpy as np
    from .. import io as vio
    dev = next(model.parameters()).device
    was_training = model.training
    model.eval()
    try:
        if translator is None:
            translator = TranslatorMultimodalVI(model, fields, beam_size=beam_size, n_best=1, max_length=max_length,
                                                global_scorer=GNMTGlobalScorer(0., -0.), cuda=True,
                                                test_img_feats=np.zeros((1, 1), np.float32),
                                                multimodal_model_type="vi-model1")
            translator.return_attention = False
        it = vio.OrderedIterator(dataset, batch_size, train=False, device=dev, rank=0, world=1)
        itos = fields["tgt"].vocab.itos
        hyps = [None] * len(dataset)
        with torch.no_grad():
            for batch in it:
                ret = translator.translate_batch(batch, None, None)
                for j, i in enumerate(batch.indices.tolist()):
                    toks = ret["predictions"][j][0]
                    words = [itos[t] for t in toks]
                    if words and words[-1] == EOS_WORD:
                        words = words[:-1]                      # translate_mm_vi.py writes the sentence without </s>
                    hyps[i] = words
    finally:
        model.train(was_training)
    if output is not None:
        with open(output, "w", encoding="utf-8") as f:
            for w in hyps:
                f.write(" ".join(w) + "\n")
    return hyps
