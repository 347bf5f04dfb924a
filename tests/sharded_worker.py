"""torchrun worker for tests/test_sharding_gpu.py: every rank samples its shard, rank 0 saves the gathered batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vq_voice_swap_b200 import sharding, synth  # noqa: E402
from vq_voice_swap_b200.diffusion_model import DiffusionModel  # noqa: E402


def build(device, num_labels=None):
    model = DiffusionModel("unet", 16, num_labels=num_labels)
    synth.load_synth(model, "sharded16")
    return model.to(device).eval()


def main():
    out, total, steps, length = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model = build(dev, num_labels=3)
    labels = torch.arange(total) % 3
    full = sharding.sample_sharded(model, total, steps, seed=1234, length=length, device=dev, labels=labels)
    if dist.get_rank() == 0:
        torch.save(full.cpu(), out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
