"""python -m mocodad_b200.dropin <script.py> [script arguments]   -- run a reference script unchanged on the B200 path."""
import os
import runpy
import sys

from . import overlay_paths


def main(argv=None) -> dict:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"mocodad_b200.dropin: no such script: {argv[0]}")
    repo_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    # overlay first, then the script's own directory (what `python script.py` would have put at sys.path[0]), then the rest;
    # any `models` / `utils` the interpreter already imported from elsewhere must not shadow the overlay
    for name in [m for m in sys.modules if m in ("models", "utils") or m.startswith(("models.", "utils."))]:
        del sys.modules[name]
    sys.path[:] = overlay_paths() + [os.path.dirname(script)] + [p for p in sys.path if p not in ("", os.getcwd())] + [repo_root]
    sys.argv = [script] + argv[1:]
    return runpy.run_path(script, run_name="__main__")   # the script's globals (model, out, ...) for callers that want them


if __name__ == "__main__":
    main()
