"""The plain-C restatement of the ATen primitives (oracle/ref_ops.c) agrees with the torch-CPU oracle."""

import ctypes as C
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2
from oracle import hotpath as O
from vq_voice_swap_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    ge.build()
    return C.CDLL(ge.ORACLE_LIB)


def _p(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("k,dil", [(3, 1), (3, 2), (3, 8), (1, 1)])
def test_conv1d(ref, k, dil):
    x = synth.normal("c/x", (2, 5, 37))
    w = synth.normal("c/w", (7, 5, k))
    b = synth.normal("c/b", (7,))
    y = torch.empty(2, 7, 37)
    ref.ref_conv1d(_p(x), _p(w), _p(b), _p(y), 2, 5, 7, 37, k, dil)
    assert rel_l2(y, F.conv1d(x, w, b, padding=dil * (k // 2), dilation=dil)) < 1e-6


def test_group_norm_gelu_resize(ref):
    x = synth.normal("g/x", (2, 12, 30), std=2.0, mean=0.5)
    gamma, beta = synth.normal("g/g", (12,)), synth.normal("g/b", (12,))
    y = torch.empty_like(x)
    ref.ref_group_norm(_p(x), _p(gamma), _p(beta), _p(y), 2, 12, 30, 4)
    assert rel_l2(y, F.group_norm(x, 4, gamma, beta, 1e-5)) < 1e-6
    ref.ref_gelu(_p(x), _p(y), C.c_long(x.numel()))
    assert rel_l2(y, F.gelu(x)) < 1e-6
    pooled = torch.empty(2, 12, 15)
    ref.ref_avg_pool2(_p(x), _p(pooled), 24, 30)
    assert torch.equal(pooled, F.avg_pool1d(x, 2))
    up = torch.empty(2, 12, 60)
    ref.ref_upsample_nearest(_p(x), _p(up), 24, 30, 60)
    assert torch.equal(up, F.interpolate(x, scale_factor=2.0))
    up = torch.empty(2, 12, 7680)
    ref.ref_upsample_nearest(_p(x), _p(up), 24, 30, 7680)
    assert torch.equal(up, F.interpolate(x, 7680))


def test_vq_argmin_matches_wherever_decisive(ref):
    d = synth.normal("v/d", (40, 24))
    x = synth.normal("v/x", (3, 24, 11))
    idx = torch.empty(3, 11, dtype=torch.int64)
    ref.ref_vq_argmin(_p(x), _p(d), _p(idx), 3, 24, 11, 40)
    want = O.vq_encode(d, x)
    gap, mag = O.vq_top2_gap(d, x)
    decisive = gap > 64 * np.finfo(np.float32).eps * mag
    assert torch.equal(idx[decisive], want[decisive]) and decisive.float().mean() > 0.95


def test_ddpm_step(ref):
    x, eps, noise = (synth.normal(f"d/{n}", (1, 1, 64)) for n in "xen")
    out = torch.empty_like(x)
    ab = O.make_alpha_bar("exp")
    t, step = torch.tensor([0.6]), 0.02
    ref.ref_ddpm_step(_p(x), _p(eps), _p(noise), _p(out), C.c_long(64), C.c_float(float(ab(t))), C.c_float(float(ab(t - step))), 0)
    assert rel_l2(out, O.ddpm_previous(ab, x, t, step, eps, noise)) < 1e-5
