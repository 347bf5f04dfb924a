"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the ctypes
mirrors match the C structs, and the host modules keep the reference's API surface."""

import ctypes as C
import json
import os
import re
import subprocess
import sys

import pytest
import torch

from helpers import key_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vqvs.h")


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    ge.build()
    from vq_voice_swap_b200 import lib

    return lib


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vqvs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = built.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vqvs.h but not exported"
    assert sorted(built.SIGNATURES) == declared
    assert lib.vqvs_abi_version() == built.ABI_VERSION


def test_ctypes_structs_match_header(built, tmp_path):
    names = {"VqvsConv": built.Conv, "VqvsGnFinalize": built.GnFinalize, "VqvsConvIn": built.ConvIn,
             "VqvsConvOut": built.ConvOut, "VqvsDdpmFinish": built.DdpmFinish, "VqvsTimeEmbed": built.TimeEmbed,
             "VqvsFilm": built.Film, "VqvsMemset": built.Memset, "VqvsOp": built.Op}
    lines = []
    for cname, ct in names.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in ct._fields_:
            lines.append(f'printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vqvs.h"\nint main(void){' + "".join(lines) + "return 0;}")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in names.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for field, _ in ct._fields_:
            assert int(out[f"{cname}.{field}"]) == getattr(ct, field).offset, f"{cname}.{field}"


def test_bad_arguments_return_errors_not_crashes(built):
    lib = built.load()
    d = built.Conv()  # all zero
    assert lib.vqvs_conv1d_fused(C.byref(d), None) == -1
    assert b"conv" in lib.vqvs_last_error()
    assert lib.vqvs_run(None, 3, None) == -1
    assert lib.vqvs_packed_weight_bytes(64, 24, 3, 0, 0) == -1  # 24 channels: not a multiple of 16
    assert lib.vqvs_packed_weight_bytes(64, 128, 3, 128, 0) == (128 * 3 + 128) * 64 * 4   # bf16 hi + lo
    assert lib.vqvs_packed_weight_bytes(64, 128, 3, 128, 1) == (128 * 3 + 128) * 64 * 2   # one fp16 image
    assert lib.vqvs_packed_weight_bytes(64, 128, 3, 128, 7) == -1                         # unknown operand format


STATE_TABLES = {
    "diffusion_unet32": lambda: _dm("unet", 32),
    "diffusion_unet16": lambda: _dm("unet", 16),
    "diffusion_unet16_cond": lambda: _dm("unet", 16, num_labels=5, cond_channels=48),
    "diffusion_unet16_dropout": lambda: _dm("unet", 16, dropout=0.1),
    "vqvae_unet32": lambda: _vqvae(base_channels=32, pred_name="unet", num_labels=8),
    "vqvae16": lambda: _vqvae(base_channels=16, num_labels=3, cond_mult=3, dictionary_size=64, pred_name="unet"),
    "classifier16": lambda: _classifier(num_labels=7, base_channels=16),
    "classifier32": lambda: _classifier(num_labels=100, base_channels=32),
}


def _classifier(**k):
    from vq_voice_swap_b200.classifier import Classifier

    return Classifier(**k)


def _dm(*a, **k):
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    return DiffusionModel(*a, **k)


def _vqvae(**k):
    from vq_voice_swap_b200.vq_vae import VQVAE

    return VQVAE(**k)


@pytest.mark.parametrize("name", sorted(STATE_TABLES))
def test_state_dict_layout_equals_reference(name):
    m = STATE_TABLES[name]()
    mine = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    assert mine == key_table()[name]
    assert json.loads(json.dumps(m.save_kwargs())) == key_table()["save_kwargs"][name]  # (JSON turns tuples into lists)


def test_checkpoint_roundtrip(tmp_path):
    m = _dm("unet", 16, num_labels=3)
    path = str(tmp_path / "m.pt")
    m.save(path)
    blob = torch.load(path, map_location="cpu")
    assert set(blob) == {"kwargs", "state_dict"}
    m2 = type(m).load(path)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_zero_init_conventions():
    """Output convs start at zero, FiLM Linear at 0.1x (reference models/unet.py:277,286-294)."""
    from vq_voice_swap_b200.unet import ResBlock

    torch.manual_seed(0)
    blk = ResBlock(32, emb_channels=64, out_channels=64)
    assert float(blk.post_cond[1].weight.abs().max()) == 0.0
    assert float(blk.post_cond[1].bias.abs().max()) == 0.0
    torch.manual_seed(0)
    torch.nn.Conv1d(32, 64, 1)
    ref = torch.nn.Linear(64, 128)
    assert torch.allclose(blk.cond_layers[1].weight, ref.weight * 0.1)


def test_factories_and_errors():
    from vq_voice_swap_b200.diffusion import CosSchedule, ExpSchedule, make_schedule
    from vq_voice_swap_b200.make import make_encoder, make_predictor

    assert isinstance(make_schedule("exp"), ExpSchedule) and isinstance(make_schedule("cos"), CosSchedule)
    with pytest.raises(ValueError, match="unknown schedule"):
        make_schedule("linear")
    with pytest.raises(ValueError, match="unknown predictor"):
        make_predictor("nope")
    with pytest.raises(ValueError, match="unknown encoder"):
        make_encoder("nope")
    assert make_encoder("unet128", base_channels=16, cond_mult=2).downsample_rate == 128
    p = make_predictor("unet", base_channels=16)
    assert p.downsample_rate == 256
    with pytest.raises(AssertionError, match="labels"):
        p(torch.zeros(1, 1, 256), torch.zeros(1), labels=torch.zeros(1, dtype=torch.long))


def test_schedules_match_closed_form():
    from vq_voice_swap_b200.diffusion import make_schedule

    t = torch.linspace(0, 1, 11)
    assert torch.allclose(make_schedule("exp")(t), torch.exp(torch.log(torch.tensor(1e-5)) * t ** 2), atol=1e-7)
    assert torch.allclose(make_schedule("cos")(t), torch.cos(t * torch.pi / 2) ** 2)
    assert abs(float(make_schedule("exp")(torch.tensor(1.0))) - 1e-5) < 1e-9


def test_no_cuda_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    p = _dm("unet", 16).predictor
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p(torch.zeros(1, 1, 256), torch.zeros(1))


def test_reference_namespace_shim():
    """`import vq_voice_swap...` (what the reference's scripts do) resolves to this implementation."""
    import vq_voice_swap
    from vq_voice_swap.dataset import ChunkReader, ChunkWriter  # noqa: F401
    from vq_voice_swap.diffusion_model import DiffusionModel
    from vq_voice_swap.models import Classifier, EncoderPredictor  # noqa: F401
    from vq_voice_swap.vq_vae import VQVAE

    assert vq_voice_swap.__file__.startswith(ROOT)
    assert DiffusionModel is type(_dm("unet", 16)) and VQVAE.__mro__[1] is DiffusionModel


def test_classifier_matches_reference_on_cpu(golden):
    """Parameter layout / semantics of the guidance classifier through its ATen diagnostic (`forward_aten`, any device):
    logits and the guidance gradient of reference sample_diffusion.py:34-42 against the live reference's outputs.  The
    product path (`Classifier.forward`) is the native program and refuses CPU tensors (checked below)."""
    import pytest
    import torch.nn.functional as F

    from vq_voice_swap_b200.classifier import forward_aten

    from helpers import model_sd, rel_l2
    from vq_voice_swap_b200 import synth

    g = golden("classifier_bc16.npz")
    clf = _classifier(num_labels=7, base_channels=16).eval()
    clf.load_state_dict(model_sd("classifier16", "clf16"))
    x = synth.normal("clf16/x", (2, 1, 1024))
    ts = torch.tensor([0.8, 0.25])
    labels = torch.tensor([3, 6])
    assert rel_l2(forward_aten(clf, x, ts).detach(), g["logits"]) <= 1e-5
    assert rel_l2(forward_aten(clf, x, ts, use_checkpoint=True).detach(), g["logits"]) <= 1e-5
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clf(x, ts)
    with pytest.raises(RuntimeError, match="ATen diagnostic"):
        clf.stem(x, ts)
    xg = x.clone().requires_grad_()
    logp = F.log_softmax(forward_aten(clf, xg, ts), dim=-1)
    grad = torch.autograd.grad(logp[range(2), labels].sum(), xg)[0] * 2.5
    assert rel_l2(grad, g["grad"]) <= 1e-4


def test_classifier_zero_head_and_checkpoint_roundtrip(tmp_path):
    torch.manual_seed(0)
    clf = _classifier(num_labels=5, base_channels=16)
    assert float(clf.out[1].weight.abs().max()) == 0.0  # reference models/classifier.py:27-29
    path = str(tmp_path / "clf.pt")
    clf.save(path)
    again = type(clf).load(path)
    assert again.save_kwargs() == clf.save_kwargs() and again.num_labels == 5
    pred = _dm("unet", 16).predictor
    assert clf.stem.load_from_predictor(pred) > 0  # reference models/classifier.py:123-130


def test_audio_io_wav_roundtrip_uses_the_reference_sample_convention(tmp_path):
    """ffmpeg-free ChunkWriter / ChunkReader (reference dataset.py:167-303 pipes s16le through ffmpeg): a float chunk is
    clipped, scaled by 2**15 - 1 and truncated toward zero on write (:296-299) and divided by 2**15 on read (:225)."""
    import numpy as np
    import wave

    from vq_voice_swap_b200.audio_io import ChunkReader, ChunkWriter, decode_u_law, encode_u_law

    x = np.concatenate([np.linspace(-1.2, 1.2, 4001), [0.0, 1e-5, -1e-5, 0.5, -0.5]]).astype(np.float32)
    path = str(tmp_path / "a.wav")
    wr = ChunkWriter(path, 16000)
    wr.write(x[:1000])
    wr.write(x[1000:])
    wr.close()
    with wave.open(path, "rb") as f:
        assert (f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()) == (16000, 1, 2, len(x))
        pcm = np.frombuffer(f.readframes(len(x)), dtype="<i2")
    assert np.array_equal(pcm, (np.clip(x, -1, 1) * (2 ** 15 - 1)).astype("int16"))  # the reference's exact expression
    rd = ChunkReader(path, 16000)
    a, b, c = rd.read(1500), rd.read(10 ** 6), rd.read(10)
    rd.close()
    assert c is None and len(a) == 1500 and len(a) + len(b) == len(x)
    assert np.array_equal(np.concatenate([a, b]), pcm.astype("float32") / 2 ** 15)
    # u-law: written decoded to linear, read re-encoded (reference encode_from_linear / decode_to_linear)
    u = np.linspace(-1, 1, 513).astype(np.float32)
    assert np.allclose(encode_u_law(decode_u_law(u)), u, atol=1e-6)
    path_u = str(tmp_path / "u.wav")
    wr = ChunkWriter(path_u, 16000, encoding="ulaw")
    wr.write(u)
    wr.close()
    rd = ChunkReader(path_u, 16000, encoding="ulaw")
    back = rd.read(len(u))
    rd.close()
    assert np.abs(back - u).max() < 0.02  # 16-bit linear quantisation seen through the u-law curve near zero
    import pytest

    with pytest.raises(ValueError):
        ChunkReader(str(tmp_path / "a.mp3"), 16000)


def test_operand_format_policy(monkeypatch):
    """engine.conv_precision: the default rule (fp16 single products for C_out >= 4*bc and for 2*bc blocks at <= 1/16 of the input
    length, bf16x3 elsewhere), the opt-in thresholds and the study knobs, on the unet64 block list."""
    from vq_voice_swap_b200 import engine
    from vq_voice_swap_b200 import lib as L
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    for k in ("VQVS_PREC", "VQVS_F16_FROM", "VQVS_F16_SHORT", "VQVS_BF16X3_FIRST", "VQVS_BF16X3_LAST", "VQVS_F16_BLOCKS"):
        monkeypatch.delenv(k, raising=False)
    blocks = engine._predictor_blocks(DiffusionModel("unet", 64).predictor)
    assert len(blocks) == 65

    def policy():
        rel, out = 1.0, []
        for b in blocks:
            rel *= float(b.scale_factor)
            out.append(engine.conv_precision(b.out_channels, 64, rel))
        return out

    p = policy()
    f16 = [i for i, v in enumerate(p) if v == L.PREC_F16]
    # 128-channel blocks whose output is <= 4000 samples (11-14 down, 46-48 up) + every 256/512-channel block (15-45)
    assert f16 == list(range(11, 49))
    assert all(p[i] == L.PREC_BF16X3 for i in list(range(0, 11)) + list(range(49, 65)))
    assert engine.conv_precision(512, None) == L.PREC_BF16X3  # encoders / single blocks: no reduced formats
    monkeypatch.setenv("VQVS_F16_SHORT", "0")
    assert [i for i, v in enumerate(policy()) if v == L.PREC_F16] == list(range(15, 46))
    monkeypatch.delenv("VQVS_F16_SHORT")
    monkeypatch.setenv("VQVS_F16_FROM", "2")
    assert [i for i, v in enumerate(policy()) if v == L.PREC_F16] == list(range(6, 58))
    monkeypatch.setenv("VQVS_F16_FROM", "1")
    assert all(v == L.PREC_F16 for v in policy())
    monkeypatch.setenv("VQVS_PREC", "bf16x3")
    assert all(v == L.PREC_BF16X3 for v in policy())


def test_workspace_bytes_contract(built):
    """vqvs_workspace_bytes: 0 for ops that live off their descriptor's buffers, the attention pool's own figure
    for it, -1 with a message for unknown kinds -- all host-side, no GPU."""
    from vq_voice_swap_b200 import lib as L

    lib = L.load()
    conv = L.Conv()
    assert lib.vqvs_workspace_bytes(L.OP_CONV_UMMA, C.addressof(conv)) == 0
    ap = L.AttnPool()
    ap.batch, ap.c, ap.t, ap.heads = 4, 512, 125, 8
    want = lib.vqvs_attnpool_workspace_bytes(4, 512, 125, 8)
    assert want > 0 and lib.vqvs_workspace_bytes(L.OP_ATTNPOOL_FWD, C.addressof(ap)) == want
    assert lib.vqvs_workspace_bytes(L.OP_ATTNPOOL_BWD, C.addressof(ap)) == want
    assert lib.vqvs_workspace_bytes(99, C.addressof(conv)) == -1
    assert b"unknown op kind" in lib.vqvs_last_error()
    assert lib.vqvs_workspace_bytes(L.OP_CONV_UMMA, None) == -1
