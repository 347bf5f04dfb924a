"""Builders shared by the oracle tests (CPU) and the parity tests (GPU).

Every builder returns (state_dict, inputs...) made only from
vq_voice_swap_b200.synth, i.e. exactly what tests/golden/make_golden.py fed to
the live reference.
"""

import json
import os

import numpy as np
import torch

from vq_voice_swap_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b) -> float:
    a = torch.as_tensor(np.asarray(a)).double().flatten()
    b = torch.as_tensor(np.asarray(b)).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def key_table():
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        return json.load(f)


_DT = {"torch.float32": torch.float32, "torch.int64": torch.int64}


def resblock_shapes(ctor):
    """State-dict layout of one reference ResBlock (models/unet.py:248-305)."""
    c_in = ctor["channels"]
    c_out = ctor.get("out_channels") or c_in
    emb = ctor.get("emb_channels")
    f32 = torch.float32
    shapes = {}
    if c_in != c_out:
        shapes["skip.1.weight"] = ((c_out, c_in, 1), f32)
        shapes["skip.1.bias"] = ((c_out,), f32)
    if emb:
        shapes["cond_layers.1.weight"] = ((2 * c_out, emb), f32)
        shapes["cond_layers.1.bias"] = ((2 * c_out,), f32)
    shapes["pre_cond.0.0.weight"] = ((c_in,), f32)
    shapes["pre_cond.0.0.bias"] = ((c_in,), f32)
    shapes["pre_cond.2.weight"] = ((c_out, c_in, 3), f32)
    shapes["pre_cond.2.bias"] = ((c_out,), f32)
    shapes["pre_cond.3.weight"] = ((c_out,), f32)
    shapes["pre_cond.3.bias"] = ((c_out,), f32)
    shapes["post_cond.1.weight"] = ((c_out, c_out, 3), f32)
    shapes["post_cond.1.bias"] = ((c_out,), f32)
    return shapes


def resblock_case(name, kw):
    ctor = kw["ctor"]
    sd = synth.synth_state_dict(resblock_shapes(ctor), tag=f"rb/{name}")
    x = synth.normal(f"rb/{name}/x", (kw["batch"], ctor["channels"], kw["t"]))
    emb = None
    if ctor.get("emb_channels"):
        emb = synth.normal(f"rb/{name}/emb", (kw["batch"], ctor["emb_channels"]))
    return sd, x, emb


def model_sd(table_name, tag):
    shapes = {k: (tuple(s), _DT[d]) for k, s, d in key_table()[table_name]}
    return synth.synth_state_dict(shapes, tag=tag)


def shapes_from_table(rows):
    return {k: (tuple(s), _DT[d]) for k, s, d in rows}
