"""Scorer object of the reference's beam search (onmt/translate/Beam.py:156-183).

The published decode flags are ``-alpha 0. -beta -0.`` (opts.py:425-429): the GNMT length / coverage
penalties vanish and a hypothesis' score is its summed log-probability.  That is the only scoring the
device-side beam (csrc/decode.cu: beam_advance_kernel) implements; other values are rejected up front.
The per-sentence ``Beam`` object of the reference (Beam.py:5-153) has no counterpart here: its state
(scores, back pointers, finished list) lives in device arrays for all sentences of a batch at once.
"""


class GNMTGlobalScorer(object):
    def __init__(self, alpha, beta):
        self.alpha, self.beta = float(alpha), float(beta)

    def is_plain_logprob(self):
        return self.alpha == 0.0 and self.beta == 0.0
