from vq_voice_swap_b200.base import Encoder, Predictor, Savable, atomic_save  # noqa: F401
from vq_voice_swap_b200.classifier import Classifier, ClassifierStem, EncoderPredictor  # noqa: F401
from vq_voice_swap_b200.make import make_encoder, make_predictor  # noqa: F401
from vq_voice_swap_b200.unet import ResBlock, UNetEncoder, UNetPredictor  # noqa: F401
