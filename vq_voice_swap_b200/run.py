"""Run one of the reference's UNMODIFIED scripts on the sm_100a implementation.

    python -m vq_voice_swap_b200.run /path/to/reference/sample_diffusion.py --checkpoint-path model.pt ...

`python /path/to/reference/sample_diffusion.py` cannot be redirected with PYTHONPATH alone: Python puts the SCRIPT's
directory first on sys.path, and in a reference checkout that directory holds the reference's own `vq_voice_swap`
package.  This launcher executes the script file with this repository's drop-in namespace (`vq_voice_swap/`, re-exports
of vq_voice_swap_b200) ahead of everything else and WITHOUT the script's directory on the path.  (Equivalent by hand:
`PYTHONSAFEPATH=1 PYTHONPATH=<this repo> python <script> ...` or `python -P`.)

With VQVS_RUN_REPORT=<file> a JSON report is written when the script ends: where `vq_voice_swap` was imported from,
whether libvqvs.so was loaded and how many kernels of each kind it launched (vqvs_launch_counts)."""
import json
import os
import runpy
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report(path: str, script: str, status):
    import vq_voice_swap

    from . import lib as L

    loaded = L._lib is not None
    out = {
        "script": os.path.abspath(script),
        "exit": status,
        "vq_voice_swap_file": os.path.abspath(vq_voice_swap.__file__),
        "resolves_into_repo": os.path.abspath(vq_voice_swap.__file__).startswith(REPO + os.sep),
        "libvqvs_loaded": loaded,
        "libvqvs_path": L.LIB_PATH if loaded else None,
        "launches": L.launch_counts() if loaded else {},
    }
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    script = argv[0]
    if not os.path.isfile(script):
        print(f"vq_voice_swap_b200.run: no such script: {script}", file=sys.stderr)
        return 2
    script_dir = os.path.dirname(os.path.abspath(script))
    # the drop-in namespace first; never the script's own directory (it may hold the reference's package)
    sys.path[:] = [REPO] + [p for p in sys.path if p and os.path.abspath(p) not in (REPO, script_dir)]
    sys.argv = [script] + argv[1:]
    status = 0
    try:
        runpy.run_path(script, run_name="__main__")
    except SystemExit as e:
        status = e.code if isinstance(e.code, int) else (0 if e.code is None else 1)
    except BaseException as e:
        status = f"{type(e).__name__}: {e}"
        raise
    finally:
        report = os.environ.get("VQVS_RUN_REPORT")
        if report:
            _report(report, script, status)
    return status if isinstance(status, int) else 1


if __name__ == "__main__":
    sys.exit(main())
