#!/usr/bin/env python
"""Headline benchmark: 4-second 16 kHz audio samples per second, 50-step DDPM, unet64.

    python bench.py --gpus N --steps K --warmup W            # the sm_100a path (this repo)
    python bench.py --impl reference --gpus N ...            # the reference's CPU algorithm (oracle port)

One bench *step* = one complete `Diffusion.ddpm_sample` (50 reverse-diffusion steps) over one batch
of 64 synthetic waveforms per GPU (BASELINE.json configs[1]; at N GPUs the batch is 64*N, sharded
with no data-path collective and one NCCL all_gather of the finished samples = configs[3]).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BASE_CHANNELS = 64
PER_GPU_BATCH = 64
T = 64000
DIFFUSION_STEPS = 50
METRIC = "4-s 16 kHz audio samples/sec (50-step DDPM, unet64)"
UNIT = "samples/s"
# algorithmic HBM bytes of one UNetPredictor forward per sample (unet64, fused-ResBlock model)
BYTES_PER_SAMPLE_STEP = 1.7603e9
FLOPS_PER_SAMPLE_STEP = 144.88e9


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


def build_model(device):
    from vq_voice_swap_b200 import synth
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    model = DiffusionModel("unet", BASE_CHANNELS)
    # random-init architecture with the zero-initialised tensors re-randomised
    model.load_state_dict(synth.synth_state_dict(synth.shapes_of(model), tag=f"bench{BASE_CHANNELS}"))
    return model.to(device).eval()


def conv_algorithmic_bytes(plan):
    """Per-launch algorithmic bytes of every conv op: inputs read once (+ raw skip input), output written once."""
    from vq_voice_swap_b200 import lib as L

    out = []
    for kind, d in plan.descs:
        if kind in (L.OP_CONV_UMMA, L.OP_CONV_SIMT):
            elems = (d.c_a + d.c_b) * d.t_in + d.c_out * d.t_out
            if d.skip_mode:
                elems += (d.s_a + d.s_b) * d.t_skip
            out.append(4.0 * d.batch * elems)
        else:
            out.append(0.0)
    return out


def profile_kernels(model, x, reps=3):
    """Per-op device time of one UNet step (CUDA events between launches on the launching stream)."""
    import ctypes as C

    from vq_voice_swap_b200 import engine
    from vq_voice_swap_b200 import lib as L

    ts = torch.full((x.shape[0],), 0.5, device=x.device)
    plan = engine._predictor_plan(model.predictor, x, None)
    engine.stage_predictor_inputs(model.predictor, plan, x, ts, None, None)
    co = plan.slots["conv_out"]
    co.mode, co.out = L.OUT_EPS, plan.eps.data_ptr()
    n = len(plan.descs)
    acc = [0.0] * n
    buf = (C.c_float * n)()
    plan.run()
    for _ in range(reps):
        L.check(L.load().vqvs_run_timed(plan.ops, n, L.stream_ptr(), buf), "vqvs_run_timed")
        for i in range(n):
            acc[i] += buf[i] / reps
    by_kind = {}
    for (kind, _), ms in zip(plan.descs, acc):
        by_kind[kind] = by_kind.get(kind, 0.0) + ms
    alg = conv_algorithmic_bytes(plan)
    umma = [(ms, b) for (kind, _), ms, b in zip(plan.descs, acc, alg) if kind == L.OP_CONV_UMMA]
    return plan, acc, by_kind, umma


def traffic_from_profiles():
    """dram bytes per launch of the dominant kernel from the committed ncu summary, if any."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("conv_umma_dram_bytes_per_launch_avg")
    return None


def run_ours(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs CUDA; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from vq_voice_swap_b200 import lib as L

    L.load()
    model = build_model(dev)
    B = args.batch
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)  # each rank owns samples [rank*B, (rank+1)*B)
    x_T = torch.randn(B, 1, T, device=dev, generator=gen)
    gathered = [torch.empty_like(x_T) for _ in range(world)] if world > 1 else None

    def one_step(x):
        out = model.diffusion.ddpm_sample(x, model.predictor, args.diffusion_steps)
        if world > 1:
            dist.all_gather(gathered, out)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(x_T)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        start.record()
        for _ in range(args.steps):
            one_step(x_T)
        stop.record()
        barrier()
    ms = start.elapsed_time(stop)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---- end to end through the public API with HOST buffers --------------------------------
    host_in = torch.randn(B, 1, T).pin_memory()
    host_out = torch.empty(B, 1, T).pin_memory()

    def e2e_step():
        x = host_in.to(dev, non_blocking=True)
        out = one_step(x)
        host_out.copy_(out, non_blocking=True)

    e2e_step()
    barrier()
    start.record()
    for _ in range(args.steps):
        e2e_step()
    stop.record()
    barrier()
    e2e_ms = start.elapsed_time(stop)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B / (e2e_ms / args.steps / 1e3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel timing, roofline -----------------------------------------------------------
    plan, per_op, by_kind, umma = profile_kernels(model, x_T)
    peak, peak_src = _peaks()
    umma_ms = sum(m for m, _ in umma)
    umma_bytes = sum(b for _, b in umma)
    n_umma = max(len(umma), 1)
    achieved = (umma_bytes / n_umma) / ((umma_ms / n_umma) * 1e-3) / 1e9 if umma_ms > 0 else 0.0
    step_ms = sum(per_op)
    whole_path_gbs = B * BYTES_PER_SAMPLE_STEP * args.diffusion_steps / (ms_per_step * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "conv_umma_kernel", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "traffic": traffic_from_profiles(), "peak_source": peak_src,
        "launches_per_unet_step": len(umma), "avg_launch_ms": round(umma_ms / n_umma, 4),
        "avg_algorithmic_bytes_per_launch": round(umma_bytes / n_umma),
        "kernel_share_of_step": round(umma_ms / step_ms, 4) if step_ms else None,
        "whole_path_gbs": round(whole_path_gbs, 1), "whole_path_frac": round(whole_path_gbs / peak, 4),
        "tensor_tflops_bf16x3": round(3 * B * FLOPS_PER_SAMPLE_STEP * args.diffusion_steps / (ms_per_step * 1e-3) / 1e12, 1),
    }
    launches = (plan.n_launch) * args.diffusion_steps * args.steps
    cpu = cpu_baseline_sample(args) if (world == 1 and not args.no_cpu_baseline) else None
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32 (bf16x3 tensor-core products, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"unet{BASE_CHANNELS} unconditional DDPM, batch {B}/GPU, {args.diffusion_steps} steps, "
                               f"{T}-sample waveform (BASELINE configs[1]; x N GPUs = configs[3])",
                   "global_batch": world * B, "l2": "activations (>=1 GB per tensor) exceed the 126 MB L2; no flush needed",
                   "weights": "random-init architecture, zero-init tensors re-randomised (seeded)",
                   "backend": plan.backend, "parallelism": f"batch-sharded x{world}, one all_gather at the end"},
        "clocks": clocks.summary(),
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": B * T * 4, "d2h_bytes_per_step": B * T * 4},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port; the reference is pure Python/PyTorch)
# ---------------------------------------------------------------------------------------------
def _oracle_setup(batch):
    from oracle import hotpath as O  # the one place bench.py executes oracle/
    from vq_voice_swap_b200 import synth
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    torch.set_num_threads(os.cpu_count())
    shapes = synth.shapes_of(DiffusionModel("unet", BASE_CHANNELS))
    sd = synth.synth_state_dict(shapes, tag=f"bench{BASE_CHANNELS}")
    x = synth.normal("bench/cpu/x", (batch, 1, T))
    return O, sd, x


def _oracle_time_steps(O, sd, x, n_steps, total_steps):
    """Time `n_steps` reverse-diffusion steps (predictor + update) of a `total_steps` sampler."""
    alpha_bar = O.make_alpha_bar("exp")
    grid = [(i + 1) / total_steps for i in range(total_steps)][::-1]
    t0 = time.perf_counter()
    with torch.no_grad():
        for t in grid[:n_steps]:
            ts = torch.tensor([t] * x.shape[0])
            eps = O.unet_predictor(sd, x, ts)
            x = O.ddpm_previous(alpha_bar, x, ts, 1 / total_steps, eps, torch.randn_like(x))
    return time.perf_counter() - t0


def cpu_baseline_sample(args, batch=2, timed=2):
    O, sd, x = _oracle_setup(batch)
    _oracle_time_steps(O, sd, x, 1, args.diffusion_steps)  # warm-up
    dt = _oracle_time_steps(O, sd, x, timed, args.diffusion_steps)
    per_sampler = dt / timed * args.diffusion_steps
    return {"value": round(batch / per_sampler, 5), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle (torch CPU fp32, {torch.get_num_threads()} threads) unet{BASE_CHANNELS} batch {batch}: "
                      f"{timed} of {args.diffusion_steps} diffusion steps timed after 1 warm-up, extrapolated linearly"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    batch, timed = 2, 1
    O, sd, x = _oracle_setup(batch)
    for _ in range(args.warmup):
        _oracle_time_steps(O, sd, x, 1, args.diffusion_steps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _oracle_time_steps(O, sd, x, timed, args.diffusion_steps)
    dt = (time.perf_counter() - t0) / args.steps
    per_sampler = dt / timed * args.diffusion_steps
    value = batch / per_sampler
    sample = (f"each bench step = {timed} of {args.diffusion_steps} diffusion steps of unet{BASE_CHANNELS} at batch {batch} "
              f"on {torch.get_num_threads()} host threads; samples/s extrapolated linearly to {args.diffusion_steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"unet{BASE_CHANNELS} unconditional DDPM, {args.diffusion_steps} steps, {T}-sample waveform "
                               "(BASELINE configs[1]), reference algorithm on host CPU cores"},
        "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch (64 = BASELINE config)")
    ap.add_argument("--diffusion-steps", type=int, default=DIFFUSION_STEPS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:  # convenience: self-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517")] + sys.argv
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
