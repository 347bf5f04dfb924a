"""vq_voice_swap_b200 -- the diffusion-sampling hot path of unixpickle/vq-voice-swap on B200.

Host code is Python/PyTorch (tensors, streams, RNG); all arithmetic on the path runs in the
hand-written sm_100a kernels of libvqvs.so (csrc/), bound through the C ABI in include/vqvs.h.
"""

__all__ = ["lib", "engine", "synth"]
