from vq_voice_swap_b200.diffusion import CosSchedule, Diffusion, ExpSchedule, Schedule, make_schedule  # noqa: F401

__all__ = ["Diffusion", "make_schedule", "CosSchedule", "Schedule", "ExpSchedule"]
