"""ffmpeg-free audio I/O with the call pattern of reference dataset.py:167-303.

The reference pipes every file through an `ffmpeg` subprocess (absent on the B200 box);
the sampling scripts only ever read/write 16 kHz mono WAV, which the standard library handles:
16-bit PCM, `int16 = clip(x) * 32767` on write (reference :296-299), `/ 32768` on read (:225).
Other containers still need ffmpeg and say so.
"""

import os
import wave
from typing import Optional

import numpy as np


def encode_u_law(x: np.ndarray, mu: float = 255.0) -> np.ndarray:
    return np.sign(x) * (np.log1p(mu * np.abs(x)) / np.log1p(mu))


def decode_u_law(x: np.ndarray, mu: float = 255.0) -> np.ndarray:
    return np.sign(x) * ((1.0 + mu) ** np.abs(x) - 1.0) / mu


def _check_encoding(encoding: str):
    if encoding not in ("linear", "ulaw"):
        raise ValueError(f"unknown audio encoding: {encoding}")


def encode_from_linear(x: np.ndarray, encoding: str) -> np.ndarray:
    _check_encoding(encoding)
    return x if encoding == "linear" else encode_u_law(x)


def decode_to_linear(x: np.ndarray, encoding: str) -> np.ndarray:
    _check_encoding(encoding)
    return x if encoding == "linear" else decode_u_law(x)


def _require_wav(path: str):
    if os.path.splitext(path)[1].lower() not in (".wav", ".wave"):
        raise ValueError(f"{path}: only WAV is handled without ffmpeg; convert with `ffmpeg -i in -ar 16000 -ac 1 out.wav`")


class ChunkReader:
    """Sequential reader: read(n) -> float32 [<= n] in [-1, 1], or None at end of stream."""

    def __init__(self, path: str, sample_rate: int, encoding: str = "linear"):
        _check_encoding(encoding)
        _require_wav(path)
        self.path, self.sample_rate, self.encoding = path, sample_rate, encoding
        with wave.open(path, "rb") as f:
            if f.getsampwidth() != 2:
                raise ValueError(f"{path}: expected 16-bit PCM, got {8 * f.getsampwidth()}-bit")
            pcm = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").astype(np.float32)
            if f.getnchannels() > 1:
                pcm = pcm.reshape(-1, f.getnchannels()).mean(axis=1)
            rate = f.getframerate()
        if rate != sample_rate:  # linear-interpolation resample (ffmpeg's -ar uses a polyphase filter)
            n_out = int(round(len(pcm) * sample_rate / rate))
            pcm = np.interp(np.arange(n_out) * (rate / sample_rate), np.arange(len(pcm)), pcm).astype(np.float32)
        self._samples = np.round(pcm).astype("<i2")
        self._pos = 0

    def read_raw(self, chunk_size: int) -> Optional[bytes]:
        if self._pos >= len(self._samples):
            return None
        buf = self._samples[self._pos:self._pos + chunk_size]
        self._pos += chunk_size
        return buf.tobytes()

    def read(self, chunk_size: int) -> Optional[np.ndarray]:
        buf = self.read_raw(chunk_size)
        if buf is None:
            return None
        linear = np.frombuffer(buf, dtype="<i2").astype("float32") / (2 ** 15)
        return encode_from_linear(linear, self.encoding)

    def close(self):
        self._pos = len(self._samples)


class ChunkWriter:
    """Sequential writer: write(float chunk in [-1, 1]) ..., close()."""

    def __init__(self, path: str, sample_rate: int, encoding: str = "linear"):
        _check_encoding(encoding)
        _require_wav(path)
        self.path, self.sample_rate, self.encoding = path, sample_rate, encoding
        self._file = wave.open(path, "wb")
        self._file.setnchannels(1)
        self._file.setsampwidth(2)
        self._file.setframerate(sample_rate)

    def write(self, chunk: np.ndarray):
        chunk = decode_to_linear(np.clip(np.asarray(chunk), -1, 1), self.encoding)
        self._file.writeframes((chunk * (2 ** 15 - 1)).astype("<i2").tobytes())

    def close(self):
        self._file.close()
