"""ConvMFCCEncoder (reference models/conv_encoder.py): the encoder of the published `vqvae-unet-mfcc` checkpoint.

Same constructor, attributes and state-dict layout as the reference (`mfcc.*` buffers of torchaudio's MFCC transform,
`blocks.{i}...` convs), so reference checkpoints load.  The forward pass is libvqvs only:

    inverse mu-law -> MFCC -> deltas / delta-deltas        vqvs_mfcc39 (DFT against the module's window, filter bank, DCT)
    Conv1d(39 -> mid, 3) -> GELU                            vqvs_conv1d_umma (39 channels zero-padded to 48) + vqvs_gelu_add
    ResConv(mid, 3)                                         conv + gelu_add with the residual
    Conv1d(mid -> mid, 4, stride 2, pad 1) -> GELU          even/odd split + a k = 3 conv over [even ; odd] (2*mid channels)
    2 x ResConv(mid, 3), 4 x ResConv(mid, 1), Conv1d(mid -> out, 1)

All convs use the bf16x3 operand format (the outputs decide VQ code indices).  torchaudio is needed at construction, as
in the reference, only to own the transform's buffers; its kernels are never called.  Version 2 (80 mels, dB scale with a
data-dependent top_db clamp) is not implemented on the CUDA path and says so."""
import ctypes as C
import math
from typing import Optional

import torch
import torch.nn as nn

from . import engine
from . import lib as L
from .base import Encoder
from .unet import _Slot


class ResConv(nn.Module):
    """reference conv_encoder.py:121-133: x + gelu(conv(x)); evaluated by ConvMFCCEncoder's program."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.conv = nn.Conv1d(*args, **kwargs)


class ConvMFCCEncoder(Encoder):
    def __init__(self, base_channels: int, out_channels: int = 64, input_ulaw: bool = True, input_rate: int = 16000,
                 mfcc_rate: int = 100, version: int = 1):
        super().__init__()
        self.base_channels = base_channels
        self.out_channels = out_channels
        self.input_ulaw = input_ulaw
        self.input_rate = input_rate
        self.mfcc_rate = mfcc_rate
        self.mid_channels = mid = base_channels * 12
        self.version = version
        assert mfcc_rate % 2 == 0, "must be able to downsample MFCCs once"
        assert input_rate % mfcc_rate == 0, "must evenly downsample input sequences"
        if version != 1:
            raise NotImplementedError("ConvMFCCEncoder version 2 (dB-scaled 80-mel MFCC) is not implemented on the sm_100a path")
        from torchaudio.transforms import MFCC  # owns window / filter-bank / DCT buffers under the reference's key names

        self.hop = input_rate // mfcc_rate
        self.n_fft = 2 * self.hop
        self.mfcc = MFCC(sample_rate=input_rate, n_mfcc=13, log_mels=True,
                         melkwargs=dict(n_fft=self.n_fft, hop_length=self.hop, n_mels=40, normalized=False))
        self.blocks = nn.ModuleList([
            nn.Sequential(nn.Conv1d(13 * 3, mid, 3, padding=1), _Slot("gelu")),
            ResConv(mid, mid, 3, padding=1),
            nn.Sequential(nn.Conv1d(mid, mid, 4, stride=2, padding=1), _Slot("gelu")),
            *[ResConv(mid, mid, 3, padding=1) for _ in range(2)],
            *[ResConv(mid, mid, 1) for _ in range(4)],
            nn.Conv1d(mid, out_channels, 1),
        ])
        with torch.no_grad():  # zero output by default (reference :88-92)
            for p in self.blocks[-1].parameters():
                p.zero_()
        self._plans = engine.PlanCache()

    @property
    def downsample_rate(self) -> int:
        return self.input_rate // (self.mfcc_rate // 2)

    # -----------------------------------------------------------------------------------------
    def _tables(self, device):
        """Device copies of the transform's buffers plus the DFT basis, and the packed weight images (per parameter version)."""
        sig = (engine._signature(self), str(device))
        cache = getattr(self, "_tables_cache", None)
        if cache is not None and cache["sig"] == sig:
            return cache
        spec = self.mfcc.MelSpectrogram
        k = torch.arange(self.n_fft // 2 + 1, dtype=torch.float64)[:, None]
        i = torch.arange(self.n_fft, dtype=torch.float64)[None, :]
        ang = 2.0 * math.pi * k * i / self.n_fft
        f32 = lambda a: a.detach().to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        mid = self.mid_channels
        w0 = self.blocks[0][0].weight.detach()
        w0p = torch.zeros(mid, 48, 3, device=w0.device, dtype=w0.dtype)
        w0p[:, :39] = w0
        w2 = self.blocks[2][0].weight.detach()  # [mid, mid, 4], stride 2, pad 1: y[t] = sum_k w[k] x[2t + k - 1]
        w2p = torch.zeros(mid, 2 * mid, 3, device=w2.device, dtype=w2.dtype)
        w2p[:, :mid, 1], w2p[:, :mid, 2] = w2[:, :, 1], w2[:, :, 3]   # even phase: x_e[t], x_e[t+1]
        w2p[:, mid:, 0], w2p[:, mid:, 1] = w2[:, :, 0], w2[:, :, 2]   # odd phase:  x_o[t-1], x_o[t]
        convs = [w0p, self.blocks[1].conv.weight, w2p] + [self.blocks[j].conv.weight for j in range(3, 9)] + [self.blocks[9].weight]
        packed = [engine.pack_weights(w.to(device), None, L.PREC_BF16X3) for w in convs]
        if any(p is None for p in packed):
            raise ValueError("ConvMFCCEncoder: mid_channels (12 * base_channels) and out_channels must be multiples of 16")
        cache = dict(sig=sig, window=f32(spec.spectrogram.window), cos=f32(torch.cos(ang)), sin=f32(torch.sin(ang)),
                     fb=f32(spec.mel_scale.fb), dct=f32(self.mfcc.dct_mat), packed=packed)
        self._tables_cache = cache
        return cache

    def _conv(self, src, c_in, t, c_out, ksize, packed, bias, dst, xb=None, c_b=0):
        d = L.Conv()
        d.batch, d.c_a, d.c_b, d.t_in, d.c_out, d.t_out = src.shape[0], c_in, c_b, t, c_out, t
        d.ksize, d.dilation, d.resize, d.act, d.skip_mode = ksize, 1, L.RESIZE_NONE, 0, L.SKIP_NONE
        d.xa, d.xb, d.out = src.data_ptr(), (xb.data_ptr() if xb is not None else 0), dst.data_ptr()
        d.bias = bias.data_ptr()
        d.w_packed = packed.img.data_ptr()
        d.reserved_ = packed.prec << L.CONV_PREC_SHIFT
        L.check(L.load().vqvs_conv1d_umma(C.byref(d), L.stream_ptr(src.device)), "vqvs_conv1d_umma")

    def forward(self, x: torch.Tensor, use_checkpoint: bool = False) -> torch.Tensor:
        engine._require_cuda(x)
        assert x.shape[1] == 1, "input must only have one channel"
        engine._check_module(self)
        lib = L.load()
        dev = x.device
        with torch.no_grad(), torch.cuda.device(dev):
            tb = self._tables(dev)
            x = engine._f32(x)
            n, _, t = x.shape
            frames = t // self.hop + 1
            mid = self.mid_channels
            stream = L.stream_ptr(dev)
            mf = torch.empty(n, 13, frames, device=dev)
            h39 = torch.empty(n, 48, frames, device=dev)
            d = L.Mfcc()
            d.batch, d.t, d.n_fft, d.hop, d.n_bins, d.n_mels, d.n_mfcc = n, t, self.n_fft, self.hop, self.n_fft // 2 + 1, 40, 13
            d.frames, d.c_pad, d.ulaw = frames, 48, 1 if self.input_ulaw else 0
            d.x, d.window, d.cos_t, d.sin_t = x.data_ptr(), tb["window"].data_ptr(), tb["cos"].data_ptr(), tb["sin"].data_ptr()
            d.fb, d.dct, d.mfcc, d.out = tb["fb"].data_ptr(), tb["dct"].data_ptr(), mf.data_ptr(), h39.data_ptr()
            L.check(lib.vqvs_mfcc39(C.byref(d), stream), "vqvs_mfcc39")
            pk = tb["packed"]
            f32 = engine._f32
            tmp = torch.empty(n, mid, frames, device=dev)
            h = torch.empty(n, mid, frames, device=dev)
            # block 0: conv(39 -> mid) -> GELU
            self._conv(h39, 48, frames, mid, 3, pk[0], f32(self.blocks[0][0].bias), tmp)
            L.check(lib.vqvs_gelu_add(tmp.data_ptr(), frames, 0, 0, h.data_ptr(), frames, n * mid, frames, stream), "vqvs_gelu_add")
            # block 1: ResConv
            self._conv(h, mid, frames, mid, 3, pk[1], f32(self.blocks[1].conv.bias), tmp)
            L.check(lib.vqvs_gelu_add(tmp.data_ptr(), frames, h.data_ptr(), frames, h.data_ptr(), frames, n * mid, frames, stream),
                    "vqvs_gelu_add")
            # block 2: stride-2 k = 4 conv as a k = 3 conv over [even ; odd]
            t_half = (frames + 1) // 2
            t_out = (frames + 2 - 4) // 2 + 1
            even, odd = torch.empty(n, mid, t_half, device=dev), torch.empty(n, mid, t_half, device=dev)
            L.check(lib.vqvs_deinterleave2(h.data_ptr(), n * mid, frames, even.data_ptr(), odd.data_ptr(), t_half, stream),
                    "vqvs_deinterleave2")
            tmp2 = torch.empty(n, mid, t_half, device=dev)
            self._conv(even, mid, t_half, mid, 3, pk[2], f32(self.blocks[2][0].bias), tmp2, xb=odd, c_b=mid)
            h = torch.empty(n, mid, t_out, device=dev)
            L.check(lib.vqvs_gelu_add(tmp2.data_ptr(), t_half, 0, 0, h.data_ptr(), t_out, n * mid, t_out, stream), "vqvs_gelu_add")
            tmp = torch.empty(n, mid, t_out, device=dev)
            for j in range(3, 9):  # two k = 3 and four k = 1 residual convs
                conv = self.blocks[j].conv
                self._conv(h, mid, t_out, mid, conv.kernel_size[0], pk[j], f32(conv.bias), tmp)
                L.check(lib.vqvs_gelu_add(tmp.data_ptr(), t_out, h.data_ptr(), t_out, h.data_ptr(), t_out, n * mid, t_out, stream),
                        "vqvs_gelu_add")
            out = torch.empty(n, self.out_channels, t_out, device=dev)
            self._conv(h, mid, t_out, self.out_channels, 1, pk[9], f32(self.blocks[9].bias), out)
            return out


def _unused() -> Optional[int]:
    return None
