"""Abstract model types and the checkpoint container.

Mirrors the public surface of reference models/base.py (Predictor :12-44,
Encoder :47-60, Savable :63-127, atomic_save :130-134): same method names,
same on-disk format ``{"kwargs": ..., "state_dict": ...}`` so reference
checkpoints load here and ours load there.
"""

import functools
import os
import tempfile
from abc import abstractmethod
from typing import Any, Callable, Dict, List

import numpy as np
import torch
import torch.nn as nn


class Predictor(nn.Module):
    """An epsilon predictor: forward(xs, ts, **kwargs) -> tensor shaped like xs."""

    @abstractmethod
    def forward(self, xs: torch.Tensor, ts: torch.Tensor, **kwargs) -> torch.Tensor:
        ...

    def condition(self, **kwargs) -> Callable:
        return functools.partial(self, **kwargs)

    @abstractmethod
    def add_labels(self, n: int, end: bool = True):
        ...

    @abstractmethod
    def label_parameters(self) -> List[nn.Parameter]:
        ...

    @property
    @abstractmethod
    def downsample_rate(self) -> int:
        ...


class Encoder(nn.Module):
    """A waveform encoder: forward(xs) -> lower-resolution feature sequence."""

    @abstractmethod
    def forward(self, xs: torch.Tensor, **kwargs) -> torch.Tensor:
        ...

    @property
    @abstractmethod
    def downsample_rate(self) -> int:
        ...


def atomic_save(state: Any, path: str):
    """Write to a temporary file in a scratch directory, then rename over `path`."""
    with tempfile.TemporaryDirectory() as scratch:
        tmp = os.path.join(scratch, "out.pt")
        torch.save(state, tmp)
        os.rename(tmp, path)


class Savable(nn.Module):
    """A module that can be rebuilt from (constructor kwargs, state dict)."""

    @abstractmethod
    def save_kwargs(self) -> Dict[str, Any]:
        ...

    def save_dict(self) -> Dict[str, Any]:
        return {"kwargs": self.save_kwargs(), "state_dict": self.state_dict()}

    @classmethod
    def load_dict(cls, state: Dict[str, Any]) -> Any:
        model = cls(**state["kwargs"])
        model.load_state_dict(state["state_dict"])
        return model

    def save(self, path: str):
        atomic_save(self.save_dict(), path)

    @classmethod
    def load(cls, path: str):
        return cls.load_dict(torch.load(path, map_location="cpu"))

    def load_from_pretrained(self, model: nn.Module) -> int:
        """Copy every parameter that exists under the same name in `model`; returns elements copied."""
        theirs = dict(model.named_parameters())
        copied = 0
        with torch.no_grad():
            for name, mine in self.named_parameters():
                src = theirs.get(name)
                if src is None:
                    continue
                if src.shape != mine.shape:
                    raise RuntimeError(
                        f"Parameter {name} has shape {mine.shape} in destination but {src.shape} in source."
                    )
                mine.copy_(src)
                copied += int(np.prod(mine.shape))
        return copied
