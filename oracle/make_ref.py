"""Copies the UNMODIFIED reference (pure Python, no build step) from /root/reference into the git-ignored oracle/_ref/ so
that it travels to the GPU box with the snapshot (like built .so files do) and can be executed there as

  * the reference arm of bench.py (`--impl reference`: the reference's own modules on the host cores, kind "reference"),
  * the PyTorch-eager-on-B200 baseline (`gpu_eager_baseline` in bench.py's line; SURVEY 8d calls it "the real bar"),
  * the unmodified train.py / evaluate.py callers in tests/test_gpu_real_callers.py (PYTHONPATH shadowing of models/,
    losses/, eval/ by the drop-in package).

Nothing under oracle/_ref/ is tracked (.gitignore) and the product never imports it.  /root/reference exists only in the
authoring container: __graft_entry__.build() runs this there; on the GPU box the copied files are used as they are.

    python oracle/make_ref.py            # idempotent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("WF_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
ITEMS = ["models", "losses", "eval", "datasets", "train.py", "evaluate.py", "main.py", "requirements.txt"]


def make_ref(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"make_ref: {SRC} not present (GPU box): using oracle/_ref as shipped" if os.path.isdir(DST)
                  else f"make_ref: neither {SRC} nor oracle/_ref exist")
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for it in ITEMS:
        s, d = os.path.join(SRC, it), os.path.join(DST, it)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            shutil.copyfile(s, d)
    for root, dirs, files in os.walk(DST):                   # the source tree is read-only; the copy must be removable
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    if verbose:
        print("make_ref: copied", ", ".join(ITEMS), "->", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
