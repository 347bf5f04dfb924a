"""Native classifier guidance (BASELINE configs[4]): libvqvs forward + dgrad programs against the oracle.

Golden logits / gradient from the live reference at bc16 are in test_gpu_parity.py (test_classifier_guided_step_golden);
here: the attention pool and head kernels alone against torch autograd, and the config-5 classifier (bc32, 100 labels,
T = 64000) against the CPU oracle's logits and guidance gradient (tolerance 1e-3, north star)."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2
from oracle import hotpath as O
from vq_voice_swap_b200 import lib as L
from vq_voice_swap_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _attnpool_ref(h, gamma, beta, wqkv, bqkv, wproj, bproj, heads):
    """reference models/classifier.py:133-191 on gelu(GroupNorm(h)) (fp64 torch on the CPU)."""
    n, c, _ = h.shape
    x = F.gelu(F.group_norm(h, 32 if c % 32 == 0 else 16, gamma, beta, 1e-5))
    x = torch.cat([torch.zeros_like(x[:, :, :1]), x], dim=-1)
    qkv = F.conv1d(x, wqkv[:, :, None], bqkv)
    ch = c // heads
    q, k, v = (p.reshape(n * heads, ch, -1) for p in qkv.chunk(3, dim=1))
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(n, c, -1)
    return F.conv1d(a, wproj[:, :, None], bproj)[:, :, 0]


@pytest.mark.parametrize("n,c,t,heads,c_out", [(3, 256, 125, 4, 512), (2, 128, 2, 2, 256), (1, 64, 7, 1, 64)])
def test_attention_pool_forward_backward(n, c, t, heads, c_out):
    import ctypes as C

    lib = L.load()
    tag = f"ap/{n}/{c}/{t}"
    h = synth.normal(tag + "/h", (n, c, t)).double()
    gamma, beta = synth.normal(tag + "/g", (c,), 0.1, 1.0).double(), synth.normal(tag + "/b", (c,), 0.1).double()
    wqkv, bqkv = synth.normal(tag + "/wq", (3 * c, c), c ** -0.5).double(), synth.normal(tag + "/bq", (3 * c,), 0.5).double()
    wproj, bproj = synth.normal(tag + "/wp", (c_out, c), c ** -0.5).double(), synth.normal(tag + "/bp", (c_out,), 0.1).double()
    d_out = synth.normal(tag + "/do", (n, c_out)).double()
    hr = h.clone().requires_grad_()
    ref = _attnpool_ref(hr, gamma, beta, wqkv, bqkv, wproj, bproj, heads)
    x_act = F.gelu(F.group_norm(hr, 32 if c % 32 == 0 else 16, gamma, beta, 1e-5))
    # gradient w.r.t. the ACTIVATED tokens (what vqvs_attnpool_bwd returns; the GroupNorm part is vqvs_gelu_bwd's job)
    xa = x_act.detach().clone().requires_grad_()
    xx = torch.cat([torch.zeros_like(xa[:, :, :1]), xa], dim=-1)
    qkv = F.conv1d(xx, wqkv[:, :, None], bqkv)
    ch = c // heads
    q, k, v = (p.reshape(n * heads, ch, -1) for p in qkv.chunk(3, dim=1))
    sc = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * sc, k * sc), dim=-1)
    out2 = F.conv1d(torch.einsum("bts,bcs->bct", w, v).reshape(n, c, -1), wproj[:, :, None], bproj)[:, :, 0]
    d_act_ref = torch.autograd.grad((out2 * d_out).sum(), xa)[0]

    f32 = lambda a: a.float().to(DEV).contiguous()  # noqa: E731
    hd = f32(h)
    groups = 32 if c % 32 == 0 else 16
    stats = torch.stack([hd.double().sum(-1), (hd.double() ** 2).sum(-1)], dim=-1).contiguous()
    fin = L.GnFinalize()
    fin.batch, fin.c_a, fin.c_b, fin.groups, fin.count = n, c, 0, groups, t
    g32, b32 = f32(gamma), f32(beta)
    fin.stats_a, fin.gamma, fin.beta = stats.data_ptr(), g32.data_ptr(), b32.data_ptr()
    prep = torch.empty(5 * n * c, device=DEV)
    L.check(lib.vqvs_gn_bwd_prep(C.byref(fin), prep.data_ptr(), L.stream_ptr()))
    ws = torch.empty(lib.vqvs_attnpool_workspace_bytes(n, c, t, heads) // 4, device=DEV)
    out = torch.empty(n, c_out, device=DEV)
    d_act = torch.empty(n, c, t, device=DEV)
    tensors = [f32(wqkv), f32(bqkv), f32(wproj), f32(bproj), f32(d_out)]
    ap = L.AttnPool()
    ap.batch, ap.c, ap.t, ap.heads, ap.c_out = n, c, t, heads, c_out
    ap.h, ap.prep = hd.data_ptr(), prep.data_ptr()
    ap.w_qkv, ap.b_qkv, ap.w_proj, ap.b_proj = (x.data_ptr() for x in tensors[:4])
    ap.ws, ap.out, ap.d_out, ap.d_act = ws.data_ptr(), out.data_ptr(), tensors[4].data_ptr(), d_act.data_ptr()
    L.check(lib.vqvs_attnpool_fwd(C.byref(ap), L.stream_ptr()))
    L.check(lib.vqvs_attnpool_bwd(C.byref(ap), L.stream_ptr()))
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), ref.detach()) <= 2e-5
    assert rel_l2(d_act.cpu(), d_act_ref) <= 2e-5


def _clf(bc, labels, tag):
    from vq_voice_swap_b200.classifier import Classifier

    clf = Classifier(num_labels=labels, base_channels=bc).eval()
    sd = synth.synth_state_dict(synth.shapes_of(clf), tag=tag)
    clf.load_state_dict(sd)
    return clf.to(DEV), sd


@pytest.mark.parametrize("bc,n_labels,batch,t", [(16, 7, 3, 2048), (32, 100, 2, 64000)])
def test_classifier_logits_and_gradient_vs_oracle(bc, n_labels, batch, t):
    clf, sd = _clf(bc, n_labels, f"guid{bc}")
    x = synth.normal(f"guid{bc}/x", (batch, 1, t))
    ts = torch.linspace(0.2, 0.9, batch)
    labels = synth.integers(f"guid{bc}/labels", (batch,), n_labels)
    ref_logits = O.classifier_logits(sd, x, ts)
    ref_grad = O.classifier_cond_fn(sd, labels, 1.5)(x, ts)
    before = L.launch_counts()
    xg = x.to(DEV).requires_grad_()
    with torch.enable_grad():
        logits = clf(xg, ts.to(DEV))
        logp = F.log_softmax(logits, dim=-1)
        grad = torch.autograd.grad(logp[range(batch), labels.to(DEV)].sum(), xg)[0] * 1.5
    assert rel_l2(logits.detach().cpu(), ref_logits) <= 1e-4
    assert rel_l2(grad.cpu(), ref_grad) <= 1e-3
    # the program launched tcgen05 convs for the forward AND the transposed convs, and no CUDA-core fallback
    counts = {k: v - before[k] for k, v in L.launch_counts().items()}
    assert counts["conv_umma"] >= 27 * 4 and counts["gelu_bwd"] == 55 and counts["attnpool_bwd"] == 1 and counts["conv_simt"] == 0


def test_backward_after_second_forward_is_refused():
    clf, _ = _clf(16, 5, "guid_stale")
    x = synth.normal("guid_stale/x", (1, 1, 1024)).to(DEV).requires_grad_()
    ts = torch.tensor([0.5], device=DEV)
    with torch.enable_grad():
        first = clf(x, ts)
        clf(x, ts)
        with pytest.raises(RuntimeError, match="evaluated again"):
            first.sum().backward()


def test_encoder_predictor_golden_logits_and_guidance_gradient(golden):
    """EncoderPredictor (reference models/encoder_predictor.py) against the live reference: the backward runs through the
    whole UNet -- concat-fed up blocks, upsampling blocks, the skip stack's two-consumer gradients."""
    from vq_voice_swap_b200.classifier import EncoderPredictor

    g = golden("encoder_predictor16.npz")
    m = EncoderPredictor(base_channels=16, downsample_rate=256, num_latents=32, bottleneck_dim=16).eval()
    assert [f"{k}|{tuple(v.shape)}" for k, v in m.state_dict().items()] == list(g["keys"])
    synth.load_synth(m, tag="encpred16")
    m = m.to(DEV)
    x = synth.normal("encpred16/x", (2, 1, 1024)).to(DEV).requires_grad_()
    ts = torch.tensor([0.7, 0.3], device=DEV)
    targets = synth.integers("encpred16/targets", (2, 4), 32).to(DEV)
    before = L.launch_counts()
    with torch.enable_grad():
        logits = m(x, ts)
        assert rel_l2(logits.detach().cpu(), g["logits"]) <= 1e-4
        losses = m.losses(x, ts, targets) * targets.shape[-1]
        grad = torch.autograd.grad(losses.sum(), x)[0]
    assert rel_l2(losses.detach().cpu(), g["losses"]) <= 1e-4
    assert rel_l2(grad.cpu(), g["grad"]) <= 1e-3
    counts = {k: v - before[k] for k, v in L.launch_counts().items()}
    assert counts["conv_simt"] == 0 and counts["gelu_bwd"] >= 2 * 65 and counts["conv_in_bwd"] == 1


def test_vqvae_decode_with_encoder_predictor_guidance(monkeypatch):
    """VQVAE.decode(enc_pred=...) (reference vq_vae.py:123-145): fused decoder steps + native guidance gradient."""
    from vq_voice_swap_b200.classifier import EncoderPredictor
    from vq_voice_swap_b200.vq_vae import VQVAE

    m = VQVAE(base_channels=16, num_labels=3, cond_mult=3, dictionary_size=32, pred_name="unet")
    synth.load_synth(m, "vqvae16g")
    m = m.to(DEV).eval()
    ep = EncoderPredictor(base_channels=16, downsample_rate=256, num_latents=32, bottleneck_dim=16)
    synth.load_synth(ep, "encpred16")
    ep = ep.to(DEV).eval()
    codes = synth.integers("vqvae16g/codes", (2, 2), 32).to(DEV)
    audio = m.decode(codes, torch.tensor([1, 2], device=DEV), steps=3, constrain=True, enc_pred=ep, enc_pred_scale=0.5)
    plain = m.decode(codes, torch.tensor([1, 2], device=DEV), steps=3, constrain=True)
    assert audio.shape == (2, 1, 512) and torch.isfinite(audio).all()
    assert audio.shape == plain.shape
