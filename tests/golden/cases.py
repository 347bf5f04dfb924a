"""Case tables shared by make_golden.py (reference side) and the tests (our side)."""

# ResBlock constructor kwargs follow reference models/unet.py:249-257.
RESBLOCK_CASES = {
    # name: ctor kwargs, batch, input length
    "plain_film": dict(ctor=dict(channels=32, emb_channels=64), batch=2, t=320),
    "widen_film": dict(ctor=dict(channels=32, emb_channels=64, out_channels=64), batch=2, t=256),
    "concat_narrow": dict(ctor=dict(channels=96, emb_channels=64, out_channels=32), batch=2, t=192),
    "down_film": dict(ctor=dict(channels=64, emb_channels=64, scale_factor=0.5), batch=2, t=256),
    "up_film": dict(ctor=dict(channels=64, emb_channels=64, scale_factor=2.0), batch=2, t=100),
    "dilated32": dict(ctor=dict(channels=64, emb_channels=128, dilation=32), batch=3, t=250),
    "dilated4": dict(ctor=dict(channels=32, emb_channels=128, dilation=4), batch=1, t=250),
    "enc_plain": dict(ctor=dict(channels=32), batch=2, t=200),
    "enc_widen": dict(ctor=dict(channels=16, out_channels=32), batch=2, t=256),
    "enc_down": dict(ctor=dict(channels=48, scale_factor=0.5), batch=1, t=130),
}

# ddpm_previous(x_t, ts, step, eps, noise, ...) -- reference diffusion/diffusion.py:48-90
DDPM_CASES = {
    "plain": dict(schedule="exp", ts=[0.9, 0.5, 0.04], step=0.02),
    "cos": dict(schedule="cos", ts=[0.9, 0.5, 0.04], step=0.02),
    "sigma_large": dict(schedule="exp", ts=[1.0, 0.5, 0.1], step=0.1, sigma_large=True),
    "constrain": dict(schedule="exp", ts=[0.8, 0.4, 0.02], step=0.01, constrain=True, x_std=0.5),
    "cond_fn": dict(schedule="exp", ts=[0.8, 0.4, 0.02], step=0.01, cond_fn=True),
    "cond_fn_constrain": dict(schedule="exp", ts=[0.8, 0.4, 0.02], step=0.01, cond_fn=True, constrain=True, x_std=0.5),
    "last_step": dict(schedule="exp", ts=[0.02, 0.02, 0.02], step=0.02),
}
