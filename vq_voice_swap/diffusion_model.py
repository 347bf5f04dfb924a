from vq_voice_swap_b200.diffusion_model import DiffusionModel  # noqa: F401
