"""The reference's UNMODIFIED scripts on this implementation.

The scripts are the byte-identical copies tools/fetch_reference.py placed under baseline/_ref/scripts (git-ignored; their
SHA-256 is checked against baseline/_ref/MANIFEST.json, which records identity with the reference checkout).  Each test
subprocesses `python -m vq_voice_swap_b200.run <script> ...` and reads the launcher's report: `vq_voice_swap` must
resolve into this repository, the WAV must exist, and libvqvs must have launched tcgen05 conv kernels."""
import hashlib
import json
import os
import subprocess
import sys
import wave

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
SCRIPTS = os.path.join(REF, "scripts")

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "MANIFEST.json")),
                               reason="baseline/_ref missing (python tools/fetch_reference.py)")


def _script(name):
    path = os.path.join(SCRIPTS, name)
    with open(os.path.join(REF, "MANIFEST.json")) as f:
        entry = json.load(f)["scripts/" + name]
    with open(path, "rb") as f:
        assert hashlib.sha256(f.read()).hexdigest() == entry["sha256"], f"{name} was edited after it was fetched"
    assert entry["identical_to_reference"], f"{name} is not the reference's file"
    return path


def _run(script, args, cwd, expect_ok=True, timeout=600):
    report = os.path.join(cwd, "report.json")
    env = dict(os.environ, VQVS_RUN_REPORT=report, PYTHONPATH="")
    env.pop("VQVS_BACKEND", None)
    p = subprocess.run([sys.executable, "-m", "vq_voice_swap_b200.run", script] + args, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=timeout)
    rep = json.load(open(report)) if os.path.exists(report) else None
    if expect_ok:
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return p, rep


def _wav(path):
    with wave.open(path, "rb") as f:
        assert f.getframerate() == 16000 and f.getnchannels() == 1 and f.getsampwidth() == 2
        return np.frombuffer(f.readframes(f.getnframes()), dtype="<i2")


def _save(model, tag, path):
    from vq_voice_swap_b200 import synth

    synth.load_synth(model, tag)
    model.save(path)


# ---------------------------------------------------------------------------
# no GPU: the launcher must redirect the imports and the path must refuse to run on the CPU
# ---------------------------------------------------------------------------
@needs_ref
@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_launcher_redirects_imports_and_refuses_cpu(tmp_path):
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    ckpt = str(tmp_path / "m.pt")
    _save(DiffusionModel("unet", 16), "dropin/cpu", ckpt)
    p, rep = _run(_script("sample_diffusion.py"), ["--checkpoint-path", ckpt, "--sample-steps", "2",
                                                    "--sample-path", str(tmp_path / "o.wav")], str(tmp_path), expect_ok=False)
    assert p.returncode != 0
    assert "no CPU fallback" in p.stderr
    assert rep["resolves_into_repo"], rep
    assert "baseline" not in rep["vq_voice_swap_file"]


@needs_ref
def test_plain_python_invocation_would_pick_the_reference_package():
    """Why the launcher exists: next to the scripts of a checkout sits the reference's own package, and Python puts the
    script's directory first on sys.path (baseline/_ref/scripts has no package, so emulate the checkout layout)."""
    code = ("import sys; sys.path.insert(0, %r); import vq_voice_swap, os; "
            "print(os.path.abspath(vq_voice_swap.__file__))" % REF)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/", env=dict(os.environ, PYTHONPATH=ROOT))
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip().startswith(REF)


# ---------------------------------------------------------------------------
# GPU: the scripts run end to end
# ---------------------------------------------------------------------------
@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("num_labels", [None, 4])
def test_sample_diffusion_unmodified(tmp_path, num_labels):
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    ckpt, out = str(tmp_path / "m.pt"), str(tmp_path / "sample.wav")
    _save(DiffusionModel("unet", 32, num_labels=num_labels), f"dropin/{num_labels}", ckpt)
    steps = 3
    _, rep = _run(_script("sample_diffusion.py"), ["--checkpoint-path", ckpt, "--sample-steps", str(steps), "--sample-path", out],
                  str(tmp_path))
    assert rep["resolves_into_repo"] and rep["libvqvs_loaded"], rep
    pcm = _wav(out)
    assert len(pcm) == 64000 and np.abs(pcm).max() > 0
    # 65 ResBlocks x 2 convs per predictor forward, every one on the tcgen05 kernel; the labelled model goes through
    # functools.partial(model.predictor, labels=...) (reference sample_diffusion.py:108-114) and must stay on the fused path
    assert rep["launches"]["conv_umma"] == 130 * steps, rep["launches"]
    assert rep["launches"]["conv_simt"] == 0
    assert rep["launches"]["conv_out"] == steps


@needs_ref
@pytest.mark.gpu
def test_sample_diffusion_many_samples_with_classifier(tmp_path):
    from vq_voice_swap_b200.classifier import Classifier
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    ckpt, clf, out = str(tmp_path / "m.pt"), str(tmp_path / "c.pt"), str(tmp_path / "samples")
    _save(DiffusionModel("unet", 32), "dropin/guided", ckpt)
    _save(Classifier(num_labels=5, base_channels=16), "dropin/clf", clf)
    _, rep = _run(_script("sample_diffusion.py"),
                  ["--checkpoint-path", ckpt, "--classifier-path", clf, "--classifier-scale", "0.5", "--target-class", "2",
                   "--sample-steps", "2", "--batch-size", "2", "--num-samples", "3", "--sample-path", out], str(tmp_path))
    assert rep["resolves_into_repo"] and rep["libvqvs_loaded"], rep
    files = sorted(os.listdir(out))
    assert files == ["sample_000000.wav", "sample_000001.wav", "sample_000002.wav"]
    for f in files:
        assert len(_wav(os.path.join(out, f))) == 64000
    assert rep["launches"]["conv_umma"] >= 130 * 2 * 2  # two batches x two steps of the predictor (+ the guidance model)


@needs_ref
@pytest.mark.gpu
def test_sample_vqvae_unmodified_check_vq(tmp_path):
    from vq_voice_swap_b200 import synth
    from vq_voice_swap_b200.audio_io import ChunkWriter
    from vq_voice_swap_b200.vq_vae import VQVAE

    ckpt, wav_in, wav_out = str(tmp_path / "vqvae.pt"), str(tmp_path / "in.wav"), str(tmp_path / "out.wav")
    _save(VQVAE(base_channels=32, pred_name="unet", enc_name="unet", num_labels=4), "dropin/vqvae", ckpt)
    w = ChunkWriter(wav_in, 16000)
    w.write((0.3 * synth.normal("dropin/wave", (64000,))).clamp(-1, 1).numpy())
    w.close()
    p, rep = _run(_script("sample_vqvae.py"), ["--label", "1", "--input-file", wav_in, "--sample-steps", "3", "--check-vq",
                                               ckpt, wav_out], str(tmp_path))
    assert rep["resolves_into_repo"] and rep["libvqvs_loaded"], rep
    assert "fraction of consistent VQ codes" in p.stdout
    assert len(_wav(wav_out)) == 64000
    # encoder (26 ResBlocks + head) twice (encode, --check-vq) + 3 decoder steps (65 ResBlocks + cond_proj each)
    assert rep["launches"]["conv_umma"] == 2 * (26 * 2 + 1) + 3 * (130 + 1), rep["launches"]
