"""
Deterministic synthetic weights and inputs.

There is no network for checkpoints or datasets, and the reference zero-inits
every ResBlock output conv (reference models/unet.py:286-294,352-356), which
would make parity tests blind to half the network.  Everything
here is a pure function of (name, shape): values come from numpy's PCG64
bit-generator seeded by a hash of the tensor name, mapped through Box-Muller in
float64, so the same numbers are produced in the build container (where the
golden vectors are made against the live reference) and on the GPU box.
"""

import hashlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def _uniforms(name: str, n: int) -> np.ndarray:
    seed = int.from_bytes(hashlib.sha256(name.encode()).digest()[:8], "little")
    return np.random.Generator(np.random.PCG64(seed)).random(n)


def normal(name: str, shape: Iterable[int], std: float = 1.0, mean: float = 0.0) -> torch.Tensor:
    """float32 tensor of N(mean, std^2) values that depends only on (name, shape)."""
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape)) if shape else 1
    m = (n + 1) // 2
    u = _uniforms(name, 2 * m)
    r = np.sqrt(-2.0 * np.log(1.0 - u[:m]))
    ang = 2.0 * np.pi * u[m:]
    z = np.concatenate([r * np.cos(ang), r * np.sin(ang)])[:n]
    return torch.from_numpy((mean + std * z).astype(np.float32).reshape(shape))


def integers(name: str, shape: Iterable[int], high: int) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape)) if shape else 1
    return torch.from_numpy((_uniforms(name, n) * high).astype(np.int64).clip(0, high - 1).reshape(shape))


def synth_state_dict(shapes: Dict[str, Tuple[Tuple[int, ...], torch.dtype]], tag: str = "w") -> Dict[str, torch.Tensor]:
    """Fill every floating tensor of a state dict (given as name -> (shape, dtype)).

    dim>1 weights ~ N(0, 1/fan_in); 1-D ``.weight`` (GroupNorm gamma) ~ N(1, 0.1^2);
    biases ~ N(0, 0.05^2); the VQ dictionary ~ N(0, 1); integer buffers keep the
    reference's constructor value (dead_rate = 100, reference vq.py:96).
    """
    out = {}
    for name, (shape, dtype) in shapes.items():
        if not dtype.is_floating_point:
            out[name] = torch.full(shape, 100, dtype=dtype)
            continue
        key = f"{tag}/{name}"
        if name.endswith("dictionary"):
            t = normal(key, shape)
        elif len(shape) > 1:
            fan_in = int(np.prod(shape[1:]))
            t = normal(key, shape, std=fan_in ** -0.5)
        elif name.endswith(".weight"):
            t = normal(key, shape, std=0.1, mean=1.0)
        else:
            t = normal(key, shape, std=0.05)
        out[name] = t.to(dtype)
    return out


def shapes_of(module: torch.nn.Module) -> Dict[str, Tuple[Tuple[int, ...], torch.dtype]]:
    return {k: (tuple(v.shape), v.dtype) for k, v in module.state_dict().items()}


def load_synth(module: torch.nn.Module, tag: str = "w") -> torch.nn.Module:
    """Overwrite a module's parameters/buffers with the deterministic synthetic set."""
    module.load_state_dict(synth_state_dict(shapes_of(module), tag))
    return module
