"""
TEST INFRASTRUCTURE -- CPU oracle for the diffusion-sampling hot path.

A functional (module-free) restatement of the reference's algorithm, driven by
a flat ``state_dict`` with the reference's key names.  All arithmetic is strict
fp32 on the CPU through ``torch.nn.functional`` -- the same third-party
dependency (PyTorch ATen, unpinned in the reference's setup.py:6; 2.11.0 here)
that executes the reference's own CPU path.  ``oracle/ref_ops.c`` restates the
primitive ATen ops themselves in plain C and is checked against these on small
cases (tests/test_oracle_c.py).

Each function cites the reference file:line it follows (paths relative to
/root/reference/vq_voice_swap/).  Pinned by tests/golden (see oracle/__init__).
"""

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

DEFAULT_MULT = (1, 1, 2, 2, 2, 4, 4, 8, 8)
DEFAULT_MIDDLE = (4, 8, 16, 32)


# ---------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------
def group_count(ch: int) -> int:
    """models/unet.py:345-349 -- largest power-of-two <= 32 dividing ch."""
    g = 32
    while ch % g:
        g //= 2
    return g


def gn(x: torch.Tensor, sd: SD, key: str) -> torch.Tensor:
    """nn.GroupNorm(eps=1e-5, affine) as built by models/unet.py:345-349."""
    return F.group_norm(x, group_count(x.shape[1]), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def gelu(x: torch.Tensor) -> torch.Tensor:
    """models/unet.py:341-342 -- nn.GELU() default = exact erf form."""
    return F.gelu(x)


def resize(x: torch.Tensor, factor: float) -> torch.Tensor:
    """models/unet.py:324-334 -- avg_pool1d(x, 2) / nearest x2 / identity."""
    if factor == 1.0:
        return x
    if factor < 1.0:
        return F.avg_pool1d(x, int(round(1.0 / factor)))
    return F.interpolate(x, scale_factor=factor)


def conv(x: torch.Tensor, sd: SD, key: str, dilation: int = 1) -> torch.Tensor:
    w = sd[key + ".weight"]
    pad = dilation * (w.shape[-1] // 2)
    return F.conv1d(x, w, sd[key + ".bias"], padding=pad, dilation=dilation)


# ---------------------------------------------------------------------------
# ResBlock -- models/unet.py:248-316
# ---------------------------------------------------------------------------
def resblock(
    x: torch.Tensor,
    emb: Optional[torch.Tensor],
    sd: SD,
    p: str,
    scale_factor: float = 1.0,
    dilation: int = 2,
) -> torch.Tensor:
    # pre_cond = [norm_act, Resize, Conv3(pad 1), GroupNorm]   (unet.py:280-285)
    h = gelu(gn(x, sd, p + "pre_cond.0.0"))
    h = resize(h, scale_factor)
    h = conv(h, sd, p + "pre_cond.2")
    h = gn(h, sd, p + "pre_cond.3")
    # FiLM (unet.py:311-314)
    if p + "cond_layers.1.weight" in sd:
        ab = F.linear(gelu(emb), sd[p + "cond_layers.1.weight"], sd[p + "cond_layers.1.bias"])
        c_out = ab.shape[1] // 2
        a, b = ab[:, :c_out, None], ab[:, c_out:, None]
        h = h * (a + 1) + b
    # post_cond = [GELU, (Dropout), Conv3 dilated]   (unet.py:286-305); eval mode -> no dropout
    post = p + ("post_cond.1" if p + "post_cond.1.weight" in sd else "post_cond.2")
    h = conv(gelu(h), sd, post, dilation=dilation)
    # skip = [Resize, Conv1x1 | Identity] on the raw input   (unet.py:265-271, 316)
    s = resize(x, scale_factor)
    if p + "skip.1.weight" in sd:
        s = conv(s, sd, p + "skip.1")
    return s + h


# ---------------------------------------------------------------------------
# time embedding -- models/wavegrad.py:352-373, models/unet.py:40-45,133-135
# ---------------------------------------------------------------------------
def time_embedding(ts: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    ch = sd[p + "time_embed.proj.weight"].shape[0]
    half = ch // 2
    freqs = torch.exp(-math.log(100.0 / 0.1) * torch.arange(half, dtype=torch.float32) / (half - 1)) * 100.0
    args = ts[:, None] * freqs[None].to(ts)
    e = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    e = F.linear(e, sd[p + "time_embed.proj.weight"], sd[p + "time_embed.proj.bias"])
    return F.linear(gelu(e), sd[p + "time_embed_extra.1.weight"], sd[p + "time_embed_extra.1.bias"])


# ---------------------------------------------------------------------------
# UNetPredictor.forward -- models/unet.py:118-163 (structure :51-116)
# ---------------------------------------------------------------------------
def unet_predictor(
    sd: SD,
    x: torch.Tensor,
    ts: torch.Tensor,
    cond: Optional[torch.Tensor] = None,
    labels: Optional[torch.Tensor] = None,
    prefix: str = "predictor.",
    channel_mult: Sequence[int] = DEFAULT_MULT,
    middle_dilations: Sequence[int] = DEFAULT_MIDDLE,
    depth_mult: int = 2,
) -> torch.Tensor:
    p = prefix
    has_labels = p + "class_embed.weight" in sd
    has_cond = p + "cond_proj.weight" in sd
    assert (labels is None) == (not has_labels)
    assert (cond is None) == (not has_cond)

    emb = time_embedding(ts, sd, p)
    if labels is not None:
        emb = emb + F.embedding(labels, sd[p + "class_embed.weight"])

    h = conv(x, sd, p + "in_conv")
    if cond is not None:
        h = h + F.interpolate(conv(cond, sd, p + "cond_proj"), h.shape[-1])

    skips: List[torch.Tensor] = [h]
    bi = 0
    n_levels = len(channel_mult)
    for depth in range(n_levels):
        for _ in range(depth_mult):
            h = resblock(h, emb, sd, f"{p}down_blocks.{bi}.")
            skips.append(h)
            bi += 1
        if depth != n_levels - 1:
            h = resblock(h, emb, sd, f"{p}down_blocks.{bi}.", scale_factor=0.5)
            skips.append(h)
            bi += 1
    for i, d in enumerate(middle_dilations):
        h = resblock(h, emb, sd, f"{p}middle_blocks.{i}.", dilation=d)
    bi = 0
    for depth in reversed(range(n_levels)):
        for _ in range(depth_mult + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resblock(h, emb, sd, f"{p}up_blocks.{bi}.")
            bi += 1
        if depth:
            h = resblock(h, emb, sd, f"{p}up_blocks.{bi}.", scale_factor=2.0)
            bi += 1
    h = gelu(gn(h, sd, p + "out.0.0"))
    return conv(h, sd, p + "out.1")


# ---------------------------------------------------------------------------
# UNetEncoder.forward -- models/unet.py:187-241
# ---------------------------------------------------------------------------
def unet_encoder(
    sd: SD,
    x: torch.Tensor,
    prefix: str = "encoder.",
    channel_mult: Sequence[int] = DEFAULT_MULT,
    out_dilations: Sequence[int] = (),
    depth_mult: int = 2,
) -> torch.Tensor:
    p = prefix
    h = conv(x, sd, p + "in_conv")
    bi = 0
    n_levels = len(channel_mult)
    for depth in range(n_levels):
        for _ in range(depth_mult):
            h = resblock(h, None, sd, f"{p}blocks.{bi}.")
            bi += 1
        if depth != n_levels - 1:
            h = resblock(h, None, sd, f"{p}blocks.{bi}.", scale_factor=0.5)
            bi += 1
    for d in out_dilations:
        h = resblock(h, None, sd, f"{p}blocks.{bi}.", dilation=d)
        bi += 1
    h = gelu(gn(h, sd, p + "out.0.0"))
    return conv(h, sd, p + "out.1")


# ---------------------------------------------------------------------------
# VQ -- vq.py:98-143, 199-243
# ---------------------------------------------------------------------------
def vq_distances(dictionary: torch.Tensor, flat: torch.Tensor) -> torch.Tensor:
    """vq.py:199-221: ((-2*dots) + |d|^2) + |x|^2 in fp32, dots through bmm."""
    dict_norms = torch.sum(torch.pow(dictionary, 2), dim=-1)
    x_norms = torch.sum(torch.pow(flat, 2), dim=-1)
    dots = torch.bmm(dictionary[None].expand(flat.shape[0], *dictionary.shape), flat[:, :, None])[..., 0]
    return -2 * dots + dict_norms + x_norms[..., None]


def vq_encode(dictionary: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """vq.py:127-131,142: [N,C,T1] -> int64 [N,T1], first minimum wins."""
    n, c = x.shape[:2]
    flat = x.reshape(n, c, -1).permute(0, 2, 1).reshape(-1, c)
    idx = torch.argmin(vq_distances(dictionary, flat), dim=-1)
    return idx.reshape(n, *x.shape[2:])


def vq_top2_gap(dictionary: torch.Tensor, x: torch.Tensor):
    """Checker helper: gap between best and second-best distance per vector."""
    n, c = x.shape[:2]
    flat = x.reshape(n, c, -1).permute(0, 2, 1).reshape(-1, c)
    d = vq_distances(dictionary, flat)
    top2 = torch.topk(d, 2, dim=-1, largest=False).values
    return (top2[:, 1] - top2[:, 0]).reshape(n, *x.shape[2:]), d.abs().max(dim=-1).values.reshape(n, *x.shape[2:])


def vq_embed(dictionary: torch.Tensor, idxs: torch.Tensor) -> torch.Tensor:
    """vq.py:98-110: [N,T1] -> [N,C,T1]."""
    n = idxs.shape[0]
    e = F.embedding(idxs.reshape(n, -1), dictionary)
    return e.permute(0, 2, 1).reshape(n, dictionary.shape[1], *idxs.shape[1:])


# ---------------------------------------------------------------------------
# schedules -- diffusion/schedule.py:15-41
# ---------------------------------------------------------------------------
def make_alpha_bar(name: str) -> Callable[[torch.Tensor], torch.Tensor]:
    if name == "exp":
        k = -math.log(1e-5)
        return lambda t: torch.exp(-k * (t ** 2))
    if name == "cos":
        return lambda t: torch.cos(t * math.pi / 2) ** 2
    raise ValueError(f"unknown schedule: {name}")


def _bc(v: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """diffusion/diffusion.py:154-157 (materialised broadcast)."""
    while v.dim() < like.dim():
        v = v[:, None]
    return v.to(like) + torch.zeros_like(like)


# ---------------------------------------------------------------------------
# DDPM step -- diffusion/diffusion.py:48-90 (with :28-46)
# ---------------------------------------------------------------------------
def ddpm_previous(
    alpha_bar: Callable,
    x_t: torch.Tensor,
    ts: torch.Tensor,
    step,
    eps: torch.Tensor,
    noise: torch.Tensor,
    sigma_large: bool = False,
    constrain: bool = False,
    cond_fn: Optional[Callable] = None,
) -> torch.Tensor:
    ab_t = _bc(alpha_bar(ts), x_t)
    ab_p = _bc(alpha_bar(ts - step), x_t)
    alpha = ab_t / ab_p
    beta = 1 - alpha

    def to_prev(e):
        return alpha.rsqrt() * (x_t - beta * (1 - ab_t).rsqrt() * e)

    def to_eps(prev):
        return (-prev * alpha.sqrt() + x_t) * (1 - ab_t).sqrt() / beta

    sig2 = beta if sigma_large else beta * (1 - ab_p) / (1 - ab_t)
    if cond_fn is not None:
        mean = to_prev(eps)
        mean = mean + sig2 * cond_fn(mean, ts - step)
        eps = to_eps(mean)
    if constrain:
        x0 = (x_t - (1 - ab_t).sqrt() * eps) * ab_t.rsqrt()          # eps_to_x0, :28-36
        x0 = (x0 - x0.mean(dim=-1, keepdim=True)).clamp(-1, 1)       # :87
        eps = (x_t - x0 * ab_t.sqrt()) * (1 - ab_t).rsqrt()          # x0_to_eps, :38-46
    return to_prev(eps) + sig2.sqrt() * noise


# ---------------------------------------------------------------------------
# sampler loop -- diffusion/diffusion.py:92-133, noise injected explicitly
# ---------------------------------------------------------------------------
def ddpm_sample(
    alpha_bar: Callable,
    x_T: torch.Tensor,
    predictor: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
    steps: int,
    noises: Sequence[torch.Tensor],
    sigma_large: bool = False,
    constrain: bool = False,
    cond_fn: Optional[Callable] = None,
    schedule: Optional[Callable] = None,
    trace: Optional[list] = None,
) -> torch.Tensor:
    """``noises[i]`` is the Gaussian draw of loop iteration i (the reference calls
    randn_like, :62-63); the last iteration uses zeros (:127) and ignores noises."""
    x_t = x_T
    grid = [(i + 1) / steps for i in range(steps)]
    t_step = 1 / steps
    for i, t in enumerate(grid[::-1]):
        ts = torch.tensor([t] * x_T.shape[0]).to(x_T)
        if schedule is not None:
            t_step = schedule(ts) - schedule(ts - 1 / steps)
            ts = schedule(ts)
        with torch.no_grad():
            eps = predictor(x_t, ts)
            noise = torch.zeros_like(x_T) if i + 1 == steps else noises[i]
            x_t = ddpm_previous(alpha_bar, x_t, ts, t_step, eps, noise, sigma_large, constrain, cond_fn)
        if trace is not None:
            trace.append((eps, x_t))
    return x_t


# ---------------------------------------------------------------------------
# VQVAE.encode / decode -- vq_vae.py:82-145
# ---------------------------------------------------------------------------
def vqvae_encode(sd: SD, x: torch.Tensor) -> torch.Tensor:
    return vq_encode(sd["vq.dictionary"], unet_encoder(sd, x))


def vqvae_decode(
    sd: SD,
    schedule_name: str,
    codes: torch.Tensor,
    labels: Optional[torch.Tensor],
    steps: int,
    x_T: torch.Tensor,
    noises: Sequence[torch.Tensor],
    constrain: bool = False,
) -> torch.Tensor:
    cond_seq = vq_embed(sd["vq.dictionary"], codes) if codes.dim() == 2 else codes
    pred = lambda xs, ts: unet_predictor(sd, xs, ts, cond=cond_seq, labels=labels)
    return ddpm_sample(make_alpha_bar(schedule_name), x_T, pred, steps, noises, constrain=constrain)


# ---------------------------------------------------------------------------
# Classifier guidance (config 5) -- models/classifier.py:18-191, sample_diffusion.py:34-42
# ---------------------------------------------------------------------------
def attention_pool(x: torch.Tensor, sd: SD, p: str, head_channels: int = 64) -> torch.Tensor:
    """AttentionPool1d.forward + QKVAttention.forward (models/classifier.py:153-191): zero token first,
    1x1 qkv conv, per-head softmax(q k^T) v with ch^-1/4 applied to q and k, 1x1 projection, take t = 0."""
    n, c, _ = x.shape
    heads = c // min(c, head_channels)
    ch = c // heads
    x = torch.cat([torch.zeros_like(x[..., :1]), x], dim=-1)
    qkv = F.conv1d(x, sd[p + "qkv_proj.weight"], sd[p + "qkv_proj.bias"])
    q, k, v = qkv.chunk(3, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    length = x.shape[-1]
    w = torch.einsum("bct,bcs->bts", (q * scale).reshape(n * heads, ch, length), (k * scale).reshape(n * heads, ch, length))
    w = torch.softmax(w, dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v.reshape(n * heads, ch, length)).reshape(n, -1, length)
    return F.conv1d(a, sd[p + "c_proj.weight"], sd[p + "c_proj.bias"])[..., 0]


def classifier_logits(sd: SD, x: torch.Tensor, ts: torch.Tensor, channel_mult: Sequence[int] = DEFAULT_MULT,
                      depth_mult: int = 2) -> torch.Tensor:
    """Classifier.forward (models/classifier.py:31-36) over ClassifierStem.forward (:111-121): every level ends with a
    downsampling ResBlock (:84-99), then GN -> GELU -> attention pool -> GELU -> Linear."""
    p = "stem."
    emb = time_embedding(ts, sd, p)
    h = conv(x, sd, p + "in_conv")
    bi = 0
    for _ in channel_mult:
        for _ in range(depth_mult):
            h = resblock(h, emb, sd, f"{p}blocks.{bi}.")
            bi += 1
        h = resblock(h, emb, sd, f"{p}blocks.{bi}.", scale_factor=0.5)
        bi += 1
    h = gelu(gn(h, sd, p + "out.0.0"))
    h = attention_pool(h, sd, p + "out.1.")
    return F.linear(gelu(h), sd["out.1.weight"], sd["out.1.bias"])


def classifier_cond_fn(sd: SD, labels: torch.Tensor, scale: float = 1.0, **kw) -> Callable:
    """cond_fn of sample_diffusion.py:34-42: scale * d/dx sum_n log softmax(logits)[n, label_n]."""

    def cond_fn(x, ts):
        with torch.enable_grad():
            xg = x.detach().clone().requires_grad_()
            logp = F.log_softmax(classifier_logits(sd, xg, ts, **kw), dim=-1)
            grads = torch.autograd.grad(logp[range(len(xg)), labels].sum(), xg)[0]
        return grads.detach() * scale

    return cond_fn
