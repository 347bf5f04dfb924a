"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the snapshot).

    python tools/fetch_reference.py [/root/reference]

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <reference>` -- the contract's recipe.
   The reference's setup.py declares `py_modules=["vq_voice_swap"]` for what is a *package*, so the wheel pip builds
   contains no code (1 KB, dist-info only); the outcome is recorded in baseline/_ref/INSTALL.json.
2. Because of that, the package directory and the three sampling scripts are copied byte for byte (never edited, never
   committed: baseline/_ref is git-ignored) and their SHA-256 digests are written to baseline/_ref/MANIFEST.json, which
   tests/test_scripts_dropin.py re-checks before running a script.
Nothing under baseline/_ref is imported by the product; it is the reference arm of bench.py and the script-level tests."""
import hashlib, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
SCRIPTS = ("sample_diffusion.py", "sample_vqvae.py", "sample_vqvae_uncond.py")


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def main(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "vq_voice_swap")):
        print(f"{src}: no reference checkout here; baseline/_ref left as it is")
        return 0
    os.makedirs(DEST, exist_ok=True)
    tmp = "/tmp/vqvs_ref_src"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(src, tmp)  # pip may write build artefacts next to setup.py; the mount is read-only
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade",
                          "--find-links", "/opt/wheelhouse", "--target", DEST, tmp], capture_output=True, text=True)
    installed_code = os.path.isdir(os.path.join(DEST, "vq_voice_swap")) and not os.path.exists(os.path.join(DEST, ".copied"))
    json.dump({"pip_returncode": pip.returncode, "pip_tail": (pip.stdout + pip.stderr)[-600:],
               "wheel_contained_package": installed_code,
               "note": "setup.py lists py_modules=['vq_voice_swap'] for a package directory: the wheel is empty, "
                       "so the package is copied verbatim instead"},
              open(os.path.join(DEST, "INSTALL.json"), "w"), indent=1)
    manifest = {}
    if not installed_code:
        pkg = os.path.join(DEST, "vq_voice_swap")
        shutil.rmtree(pkg, ignore_errors=True)
        shutil.copytree(os.path.join(src, "vq_voice_swap"), pkg, ignore=shutil.ignore_patterns("__pycache__"))
        open(os.path.join(DEST, ".copied"), "w").write("package copied verbatim from the reference checkout\n")
    os.makedirs(os.path.join(DEST, "scripts"), exist_ok=True)
    for s in SCRIPTS:
        shutil.copyfile(os.path.join(src, s), os.path.join(DEST, "scripts", s))
    for base, _, files in os.walk(DEST):
        for f in files:
            if f.endswith(".py"):
                p = os.path.join(base, f)
                rel = os.path.relpath(p, DEST)
                origin = os.path.join(src, rel[len("scripts/"):] if rel.startswith("scripts/") else rel)
                manifest[rel] = {"sha256": sha(p), "identical_to_reference": os.path.exists(origin) and sha(origin) == sha(p)}
    json.dump(manifest, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    bad = [k for k, v in manifest.items() if not v["identical_to_reference"]]
    print(f"baseline/_ref: {len(manifest)} files, {len(bad)} differ from the reference {bad[:3]}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:]))
