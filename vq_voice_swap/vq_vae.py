from vq_voice_swap_b200.vq_vae import VQVAE  # noqa: F401
