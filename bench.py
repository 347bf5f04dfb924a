#!/usr/bin/env python
"""Headline benchmark: 4-second 16 kHz audio samples per second, 50-step DDPM, unet64.

    python bench.py --gpus N --steps K --warmup W            # the sm_100a path (this repo), BASELINE configs[1] / [3]
    python bench.py --impl reference --gpus N ...            # the reference's own CPU implementation (baseline/_ref)
    python bench.py --config vqvae | guided                  # BASELINE configs[2] / [4] (same JSON contract)

One bench *step* = one complete sampling call over one batch of synthetic waveforms per GPU:
  uncond  `Diffusion.ddpm_sample`, 50 reverse-diffusion steps, unet64, 64 waveforms per GPU (x N GPUs = configs[3]);
  vqvae   `VQVAE.encode` + `VQVAE.decode(steps=100, constrain=True)`, base_channels 32, 32 waveforms per GPU;
  guided  `ddpm_sample(..., cond_fn)` with the classifier gradient of reference sample_diffusion.py:34-42, 32 per GPU.
At N GPUs the batch is sharded with no data-path collective (noise keyed by the global sample index) and one NCCL
all_gather of the finished samples.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
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

T = 64000
UNIT = "samples/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

# algorithmic HBM bytes / FLOPs per sample of one UNetPredictor forward (fused-ResBlock model) and of one
# UNetEncoder / Classifier forward
CONFIGS = {
    "uncond": dict(bc=64, batch=64, steps=50, metric="4-s 16 kHz audio samples/sec (50-step DDPM, unet64)",
                   bytes_step=1.7603e9, flops_step=144.88e9, bytes_once=0.0,
                   workload="unet{bc} unconditional DDPM, batch {b}/GPU, {s} steps, 64000-sample waveform "
                            "(BASELINE configs[1]; x N GPUs = configs[3])"),
    "vqvae": dict(bc=32, batch=32, steps=100, metric="VQ-VAE conversions/sec (encode + 100-step constrained decode, bc32)",
                  bytes_step=0.8804e9, flops_step=36.23e9, bytes_once=0.2872e9,
                  workload="VQVAE bc{bc}: UNetEncoder + VQ arg-min + {s}-step conditional decode (constrain=True), batch {b}/GPU "
                           "(BASELINE configs[2], sample_vqvae.py path)"),
    "guided": dict(bc=64, batch=32, steps=50, metric="4-s 16 kHz audio samples/sec (50-step classifier-guided DDPM, unet64)",
                   bytes_step=1.7603e9 + 3 * 0.2881e9, flops_step=144.88e9 + 3 * 9.08e9, bytes_once=0.0,
                   workload="unet{bc} DDPM with classifier guidance (Classifier bc32, 100 labels, cond_fn of "
                            "sample_diffusion.py:34-42), batch {b}/GPU, {s} steps (BASELINE configs[4])"),
}


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


# ---------------------------------------------------------------------------------------------
# synthetic models (random-init architecture, zero-initialised tensors re-randomised: SURVEY.md D5)
# ---------------------------------------------------------------------------------------------
def _state(module, tag):
    from vq_voice_swap_b200 import synth

    return synth.synth_state_dict(synth.shapes_of(module), tag=tag)


def build_model(device, bc=64):
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    model = DiffusionModel("unet", bc)
    model.load_state_dict(_state(model, f"bench{bc}"))
    return model.to(device).eval()


def build_vqvae(device, bc=32):
    from vq_voice_swap_b200.vq_vae import VQVAE

    model = VQVAE(base_channels=bc, pred_name="unet", enc_name="unet", cond_mult=16, dictionary_size=512, num_labels=8)
    model.load_state_dict(_state(model, f"bench_vqvae{bc}"))
    return model.to(device).eval()


def build_classifier(device, bc=32, num_labels=100):
    from vq_voice_swap_b200.classifier import Classifier

    clf = Classifier(num_labels=num_labels, base_channels=bc)
    clf.load_state_dict(_state(clf, f"bench_clf{bc}"))
    return clf.to(device).eval()


def make_cond_fn(classifier, labels, scale=1.0):
    """reference sample_diffusion.py:34-42 (the caller's closure, restated because it lives in the script)."""
    import torch.nn.functional as F

    def cond_fn(x, ts):
        with torch.enable_grad():
            x = x.detach().clone().requires_grad_()
            logits = classifier(x, ts)
            logprobs = F.log_softmax(logits, dim=-1)
            grads = torch.autograd.grad(logprobs[range(len(x)), labels].sum(), x)[0]
            return grads.detach() * scale

    return cond_fn


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


def profile_kernels(model, x, reps=3, cond=None, labels=None):
    """Per-op device time of one UNet step (CUDA events between launches on the launching stream)."""
    import ctypes as C

    from vq_voice_swap_b200 import engine
    from vq_voice_swap_b200 import lib as L

    ts = torch.full((x.shape[0],), 0.5, device=x.device)
    plan = engine._predictor_plan(model.predictor, x, cond)
    engine.stage_predictor_inputs(model.predictor, plan, x, ts, cond, labels)
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


PRECISION_FROM = {"default": None, "fast128": "2", "fast": "1"}  # VQVS_F16_FROM: fp16 single products from this multiple of bc
PRECISION_DTYPE = {
    "default": "fp32 storage and accumulation; tensor-core products bf16x3 (hi/lo split), one fp16 product for C_out >= 4*bc and for "
               "2*bc blocks at <= 1/16 of the input length (unet64 forward 8.9e-5, 50-step sample 2.0e-5 vs the fp32 oracle)",
    "fast128": "fp32 storage and accumulation; tensor-core products bf16x3 for C_out < 2*bc, one fp16 product from 2*bc "
               "(NOT the headline: unet64 forward 3.7e-4 vs the fp32 oracle)",
    "fast": "fp32 storage and accumulation; one fp16 tensor-core product per tap everywhere (NOT the headline: TF32-class, "
            "unet64 forward 1.0e-3, 50-step sample 2.8e-4 vs the fp32 oracle)",
}


def traffic_from_profiles():
    """Launch-weighted dram bytes per launch of the dominant kernel from the committed ncu summary, if any."""
    for name in ("r2_dram_traffic.json", "ncu_summary.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                v = json.load(f).get("conv_umma_dram_bytes_per_launch_avg")
            if v:
                return v
    return None


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
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
    from vq_voice_swap_b200 import sharding, synth

    L.load()
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    dsteps = args.diffusion_steps or cfg["steps"]
    lo = rank * B  # rank r owns global samples [r*B, (r+1)*B): noise is keyed by the GLOBAL index (sharding.keyed_noise)
    seed = 1234

    def noise_fn(i, like):
        return sharding.keyed_noise(seed, lo, B, i, T, like.device)

    clf = None
    if args.config == "vqvae":
        model = build_vqvae(dev, cfg["bc"])
        wave = sharding.keyed_noise(seed + 1, lo, B, -2, T, dev).clamp(-1, 1)
        labels = (torch.arange(lo, lo + B) % 8).to(dev)
        host_in = wave.cpu().pin_memory()

        def sample(x):
            codes = model.encode(x)
            return model.decode(codes, labels, steps=dsteps, constrain=True, noise_fn=noise_fn)
    else:
        model = build_model(dev, cfg["bc"])
        x_T = sharding.keyed_noise(seed, lo, B, -1, T, dev)
        host_in = x_T.cpu().pin_memory()
        cond_fn = None
        if args.config == "guided":
            clf = build_classifier(dev)
            labels = synth.integers("bench/guided/labels", (world * B,), 100)[lo:lo + B].to(dev)
            cond_fn = make_cond_fn(clf, labels)

        def sample(x):
            return model.diffusion.ddpm_sample(x, model.predictor, dsteps, cond_fn=cond_fn, noise_fn=noise_fn)

    dev_in = host_in.to(dev)
    gathered = [torch.empty(B, 1, T, device=dev) for _ in range(world)] if world > 1 else None

    def one_step(x):
        out = sample(x)
        if world > 1:
            dist.all_gather(gathered, out)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(n):
            fn()
        stop.record()
        barrier()
        ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    counts0 = L.launch_counts()
    for _ in range(args.warmup):
        one_step(dev_in)
    barrier()
    counts1 = L.launch_counts()
    with ClockSampler(local) as clocks:
        ms = timed(lambda: one_step(dev_in), args.steps)
    counts2 = L.launch_counts()
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---- end to end through the public API with HOST buffers --------------------------------
    host_out = torch.empty(B, 1, T).pin_memory()

    def e2e_step():
        x = host_in.to(dev, non_blocking=True)
        out = one_step(x)
        host_out.copy_(out, non_blocking=True)

    e2e_step()
    barrier()
    e2e_ms = timed(e2e_step, args.steps)
    e2e_value = world * B / (e2e_ms / args.steps / 1e3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel timing, roofline -----------------------------------------------------------
    if args.config == "vqvae":
        cond = model.vq.embed(model.encode(dev_in))
        plan, per_op, by_kind, umma = profile_kernels(model, dev_in, cond=cond, labels=labels)
    else:
        plan, per_op, by_kind, umma = profile_kernels(model, dev_in)
    peak, peak_src = _peaks()
    umma_ms = sum(m for m, _ in umma)
    umma_bytes = sum(b for _, b in umma)
    n_umma = max(len(umma), 1)
    achieved = (umma_bytes / n_umma) / ((umma_ms / n_umma) * 1e-3) / 1e9 if umma_ms > 0 else 0.0
    step_ms = sum(per_op)
    alg_bytes = B * (cfg["bytes_step"] * dsteps + cfg["bytes_once"])
    whole_path_gbs = alg_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "conv_umma_kernel", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "traffic": traffic_from_profiles(), "peak_source": peak_src,
        "launches_per_unet_step": len(umma), "avg_launch_ms": round(umma_ms / n_umma, 4),
        "avg_algorithmic_bytes_per_launch": round(umma_bytes / n_umma),
        "kernel_share_of_step": round(umma_ms / step_ms, 4) if step_ms else None,
        "whole_path_gbs": round(whole_path_gbs, 1), "whole_path_frac": round(whole_path_gbs / peak, 4),
        "fp32_equivalent_tflops": round(B * cfg["flops_step"] * dsteps / (ms_per_step * 1e-3) / 1e12, 1),
    }
    timed_launches = {k: counts2[k] - counts1[k] for k in counts2 if k != "memset" and counts2[k] != counts1[k]}
    line = {
        "metric": cfg["metric"], "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": PRECISION_DTYPE[args.precision],
        "data": "synthetic",
        "config": {"workload": cfg["workload"].format(bc=cfg["bc"], b=B, s=dsteps), "global_batch": world * B,
                   "precision": args.precision,
                   "l2": "activations (>= 0.5 GB per tensor) exceed the 126 MB L2; no flush needed",
                   "weights": "random-init architecture, zero-init tensors re-randomised (seeded)",
                   "noise": "device Philox keyed by (seed, global sample index, step): N-GPU output == 1-GPU output",
                   "backend": plan.backend, "parallelism": f"batch-sharded x{world}, one all_gather at the end"},
        "clocks": clocks.summary(),
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": B * T * 4, "d2h_bytes_per_step": B * T * 4},
        "gpu_launches": sum(timed_launches.values()), "gpu_launches_by_kernel": timed_launches,
        "roofline": roofline,
    }
    if args.config == "guided":
        line["guidance"] = guidance_share(model, clf, dev_in, labels)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args)
        if not args.no_eager:
            line["reference_eager_gpu"] = eager_gpu_sample(args, dev)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def guidance_share(model, clf, x, labels):
    """Config 5: device time of one predictor step vs one cond_fn evaluation (forward + input gradient of the classifier),
    and which library evaluates the latter."""
    cond_fn = make_cond_fn(clf, labels)
    ts = torch.full((x.shape[0],), 0.5, device=x.device)

    def ms_of(fn, n=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    t_pred = ms_of(lambda: model.predictor(x, ts))
    t_guid = ms_of(lambda: cond_fn(x, ts))
    from vq_voice_swap_b200 import classifier as C

    return {"predictor_ms": round(t_pred, 2), "cond_fn_ms": round(t_guid, 2), "cond_fn_share": round(t_guid / (t_pred + t_guid), 3),
            "cond_fn_engine": "ATen/cuDNN under autograd (VQVS_GUIDANCE=aten)" if C._use_aten() else C.GUIDANCE_ENGINE}


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (baseline/_ref, else the oracle port)
# ---------------------------------------------------------------------------------------------
class CpuArm:
    """One reverse-diffusion step (predictor + ddpm_previous) of the configured workload on host cores."""

    def __init__(self, args):
        from vq_voice_swap_b200 import synth
        from vq_voice_swap_b200.diffusion_model import DiffusionModel

        torch.set_num_threads(os.cpu_count())
        cfg = CONFIGS["uncond"]
        self.bc, self.total = cfg["bc"], args.diffusion_steps or cfg["steps"]
        shapes = synth.shapes_of(DiffusionModel("unet", self.bc))
        sd = synth.synth_state_dict(shapes, tag=f"bench{self.bc}")
        self.kind = "port"
        self.model = None
        if os.path.isdir(os.path.join(REF_DIR, "vq_voice_swap")) and not args.force_port:
            try:
                self.model = _import_reference().diffusion_model.DiffusionModel("unet", self.bc)
                self.model.load_state_dict(sd)
                self.model.eval()
                self.kind = "reference"
            except Exception as e:  # torchaudio etc. missing on the box: fall back to the port and say so
                self.note = f"baseline/_ref not importable ({type(e).__name__}: {e}); timed the oracle port"
                self.model = None
        if self.model is None:
            from oracle import hotpath as O  # the one other place bench.py executes oracle/

            self.O, self.sd = O, sd

    def x(self, batch):
        from vq_voice_swap_b200 import synth

        return synth.normal("bench/cpu/x", (batch, 1, T))

    def steps(self, x, n):
        grid = [(i + 1) / self.total for i in range(self.total)][::-1]
        t0 = time.perf_counter()
        with torch.no_grad():
            for t in grid[:n]:
                ts = torch.tensor([t] * x.shape[0])
                if self.kind == "reference":
                    eps = self.model.predictor(x, ts)
                    x = self.model.diffusion.ddpm_previous(x, ts, 1 / self.total, eps)
                else:
                    eps = self.O.unet_predictor(self.sd, x, ts)
                    x = self.O.ddpm_previous(self.O.make_alpha_bar("exp"), x, ts, 1 / self.total, eps, torch.randn_like(x))
        return time.perf_counter() - t0

    def samples_per_s(self, batch, timed):
        x = self.x(batch)
        self.steps(x, 1)  # warm-up
        dt = self.steps(x, timed)
        return batch / (dt / timed * self.total)

    def describe(self, batch, timed):
        what = ("unmodified reference modules from baseline/_ref (DiffusionModel.predictor + Diffusion.ddpm_previous)"
                if self.kind == "reference" else "oracle port (functional restatement on the same ATen CPU kernels)")
        return (f"{what}, torch CPU fp32, {torch.get_num_threads()} threads, unet{self.bc} batch {batch}: {timed} of {self.total} "
                f"diffusion steps timed after 1 warm-up, extrapolated linearly (every step does identical work)")


def _import_reference():
    """Import the unmodified reference package from baseline/_ref under its own name, without disturbing the drop-in
    namespace of this repository (both are called vq_voice_swap)."""
    import importlib

    saved = {k: v for k, v in sys.modules.items() if k == "vq_voice_swap" or k.startswith("vq_voice_swap.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_DIR)
    try:
        ref = importlib.import_module("vq_voice_swap")
        importlib.import_module("vq_voice_swap.diffusion_model")
        importlib.import_module("vq_voice_swap.models")
        assert os.path.abspath(ref.__file__).startswith(REF_DIR)
    finally:
        sys.path.remove(REF_DIR)
        mods = {k: v for k, v in sys.modules.items() if k == "vq_voice_swap" or k.startswith("vq_voice_swap.")}
        for k in mods:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref


def cpu_baseline_sample(args, batch=2, timed=2):
    arm = CpuArm(args)
    v = arm.samples_per_s(batch, timed)
    out = {"value": round(v, 5), "unit": UNIT, "cores": torch.get_num_threads(), "kind": arm.kind, "sample": arm.describe(batch, timed),
           "batch_1": round(arm.samples_per_s(1, 2), 5), "batch_4": round(arm.samples_per_s(4, 1), 5)}
    if getattr(arm, "note", None):
        out["note"] = arm.note
    return out


def eager_gpu_sample(args, dev, batch=16, timed=2):
    """Informational: the reference's modules in PyTorch eager ON THE B200 (cuDNN, TF32 convs by default) --
    the same-box, no-custom-kernels comparator.  Not the baseline of record (that is cpu_baseline)."""
    if not os.path.isdir(os.path.join(REF_DIR, "vq_voice_swap")):
        return {"unavailable": "baseline/_ref missing"}
    try:
        from vq_voice_swap_b200 import synth

        ref = _import_reference()
        cfg = CONFIGS["uncond"]
        total = args.diffusion_steps or cfg["steps"]
        model = ref.diffusion_model.DiffusionModel("unet", cfg["bc"])
        model.load_state_dict(synth.synth_state_dict(synth.shapes_of(model), tag=f"bench{cfg['bc']}"))
        model = model.to(dev).eval()
        x = torch.randn(batch, 1, T, device=dev)
        grid = [(i + 1) / total for i in range(total)][::-1]

        def run(n):
            xx = x
            with torch.no_grad():
                for t in grid[:n]:
                    ts = torch.tensor([t] * batch).to(dev)
                    xx = model.diffusion.ddpm_previous(xx, ts, 1 / total, model.predictor(xx, ts))
            return xx

        run(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(timed)
        b.record()
        torch.cuda.synchronize()
        per_sampler = a.elapsed_time(b) / 1e3 / timed * total
        del model
        torch.cuda.empty_cache()
        return {"value": round(batch / per_sampler, 3), "unit": UNIT, "batch": batch,
                "conv_precision": f"cudnn.conv.fp32_precision={torch.backends.cudnn.conv.fp32_precision}",
                "sample": f"reference modules (baseline/_ref) in torch eager on the GPU, {timed} of {total} steps timed, extrapolated"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if args.config != "uncond":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the headline config (uncond) only"}), flush=True)
        return
    batch, timed = 2, 1
    arm = CpuArm(args)
    x = arm.x(batch)
    for _ in range(args.warmup):
        arm.steps(x, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.steps(x, timed)
    dt = (time.perf_counter() - t0) / args.steps
    value = batch / (dt / timed * arm.total)
    cfg = CONFIGS["uncond"]
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": round(value, 5), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"unet{arm.bc} unconditional DDPM, {arm.total} steps, {T}-sample waveform "
                               "(BASELINE configs[1]), reference implementation on host CPU cores"},
        "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": torch.get_num_threads(), "kind": arm.kind,
                         "sample": "each bench step = " + arm.describe(batch, timed)},
        "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="uncond", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--diffusion-steps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the informational torch-eager-on-GPU comparator")
    ap.add_argument("--force-port", action="store_true", help="CPU arm: time the oracle port even if baseline/_ref exists")
    ap.add_argument("--precision", default="default", choices=["default", "fast128", "fast"],
                    help="operand-format policy of the predictor convs (engine.conv_precision): default = bf16x3, fp16 for "
                         "C_out >= 4*bc (unet64 forward 4.9e-5 vs the fp32 oracle); fast128 = fp16 from 2*bc (3.7e-4); fast = one fp16 "
                         "product everywhere (1.0e-3 per forward, 2.8e-4 on a 50-step sample: the TF32 class of the reference's own "
                         "GPU path).  Only the default is the headline configuration.")
    args = ap.parse_args()
    if PRECISION_FROM[args.precision]:
        os.environ["VQVS_F16_FROM"] = PRECISION_FROM[args.precision]
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
