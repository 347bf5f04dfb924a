from vq_voice_swap_b200.vq import VQ  # noqa: F401
