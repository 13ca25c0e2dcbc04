"""Prior-only beam search for VI model 1 (reference: onmt/translate/TranslatorMultimodalVI.py,
onmt/translate/Beam.py) with the beam bookkeeping on the device."""
from .Beam import GNMTGlobalScorer
from .TranslatorMultimodalVI import TranslatorMultimodalVI
from .validation import translate_dataset

__all__ = ["TranslatorMultimodalVI", "GNMTGlobalScorer", "translate_dataset"]
