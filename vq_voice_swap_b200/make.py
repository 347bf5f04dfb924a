"""Name -> model factories (reference models/make.py).  The UNet family and the MFCC conv encoder (version 1, the
published checkpoint's) are implemented in CUDA; WaveGrad and the version-2 MFCC encoder are outside the accelerated path."""

from typing import Optional

from .base import Encoder, Predictor
from .unet import UNetEncoder, UNetPredictor

_UNSUPPORTED_PREDICTORS = ("wavegrad",)
_UNSUPPORTED_ENCODERS = ("wavegrad", "conv-mfcc-ulaw-v2")


def make_predictor(pred_name: str, base_channels: int = 32, num_labels: Optional[int] = None,
                   cond_channels: Optional[int] = None, dropout: float = 0.0) -> Predictor:
    if pred_name == "unet":
        return UNetPredictor(base_channels=base_channels, cond_channels=cond_channels, num_labels=num_labels,
                             dropout=dropout)
    if pred_name in _UNSUPPORTED_PREDICTORS:
        raise NotImplementedError(f"predictor '{pred_name}' is not part of the sm_100a sampling path (SURVEY.md 8)")
    raise ValueError(f"unknown predictor: {pred_name}")


def make_encoder(enc_name: str, base_channels: int = 32, cond_mult: int = 16) -> Encoder:
    out = base_channels * cond_mult
    if enc_name == "unet":
        return UNetEncoder(base_channels=base_channels, out_channels=out)
    if enc_name == "unet128":  # downsample rate 128 instead of 256
        return UNetEncoder(base_channels=base_channels, channel_mult=(1, 1, 2, 2, 2, 4, 4, 8), out_channels=out)
    if enc_name == "unet128-dilated":
        return UNetEncoder(base_channels=base_channels, channel_mult=(1, 1, 2, 2, 2, 4, 4, 8),
                           out_dilations=(4, 8, 16, 32), out_channels=out)
    if enc_name in ("conv-mfcc-ulaw", "conv-mfcc-linear"):  # reference models/make.py:66-81
        from .conv_encoder import ConvMFCCEncoder

        return ConvMFCCEncoder(base_channels=base_channels, out_channels=out, input_ulaw=enc_name == "conv-mfcc-ulaw")
    if enc_name in _UNSUPPORTED_ENCODERS:
        raise NotImplementedError(f"encoder '{enc_name}' is not part of the sm_100a sampling path (SURVEY.md 8)")
    raise ValueError(f"unknown encoder: {enc_name}")
