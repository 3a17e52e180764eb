"""Copy recipe for the reference arm: place the UNMODIFIED reference files of the hot path under baseline/_ref/
(git-ignored, NOT gpurun-ignored: the copies travel to the GPU box with a snapshot, the repository history never holds
reference sources).  Run in the build container, where /root/reference exists:

    python tools/install_reference.py            # also called by __graft_entry__.build()

`pip install /root/reference` is not applicable: the reference has no packaging metadata (no setup.py / pyproject), it
is a directory of scripts.  bench.py (`--impl reference`, `cpu_baseline`, `reference_on_b200`) imports these files by
path, exactly as tests/golden/make_golden*.py import them from /root/reference."""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
# the files the training step lives in
FILES = ["modules/model.py", "modules/train.py",
         "tabular/modules/model.py", "tabular/modules/train.py",
         "celeba/module/model.py", "celeba/module/sagan.py", "celeba/module/train.py"]


def install(src=None, quiet=False):
    src = src or os.environ.get("CDG_REFERENCE", "/root/reference")
    if not os.path.isdir(src):
        return False
    for rel in FILES:
        a, b = os.path.join(src, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
            shutil.copyfile(a, b)
    if not quiet:
        print(f"reference files ({len(FILES)}) installed under {DEST}")
    return True


def installed():
    return all(os.path.exists(os.path.join(DEST, rel)) for rel in FILES)


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
