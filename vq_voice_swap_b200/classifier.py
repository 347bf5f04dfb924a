"""Place-holders for the two auxiliary guidance models the reference's scripts import
(reference models/classifier.py, models/encoder_predictor.py).  They are named in SURVEY.md 8(f)
as the next rows after the sampling path; importing works, constructing says what is missing."""

from .base import Savable


class _NotOnPathYet(Savable):
    what = ""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__} ({self.what}) is not implemented on the sm_100a path yet; "
            "ddpm_sample accepts any cond_fn callable, so a guidance model evaluated elsewhere still works"
        )

    def save_kwargs(self):
        return {}


class Classifier(_NotOnPathYet):
    what = "noised-audio classifier for classifier guidance"


class ClassifierStem(_NotOnPathYet):
    what = "classifier feature stem"


class EncoderPredictor(_NotOnPathYet):
    what = "VQ-code predictor for decode-time guidance"
