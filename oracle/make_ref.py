"""Recipe for oracle/_ref/: the UNMODIFIED reference loss file, staged for the GPU box.  TEST / BENCH
INFRASTRUCTURE - the product package never imports it.

The reference is pure Python: `src/open_clip/loss.py` needs only torch, so "building" the real
reference for this path is staging that one file where `bench.py --impl reference`, the `cpu_baseline` leg and
the `gpu_eager_baseline` leg can load it by path.  /root/reference does not exist on the GPU box; `oracle/_ref/`
is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the staged copy travels
with the snapshot like a built .so.

    python oracle/make_ref.py            # stage + verify against oracle/ref_loss.sha256
    python oracle/make_ref.py --update   # (re)write the checksum file after a reference update

`load_reference()` returns the staged module or None; callers fall back to the oracle port and say so
(`kind: "port"`).

`stage_tree()` additionally stages the reference's two Python packages (`src/open_clip`, `src/open_clip_train`:
pure Python + model-config JSON + the BPE vocabulary, 2.4 MB) as `oracle/_ref/src/` for the config-5 harness
(`scripts/train_step_harness.py`: the reference's unmodified `create_model` / `create_loss` / `train_one_epoch`
driven with a synthetic loader); `oracle/ref_tree.sha256` pins the files on the call path.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/open_clip/loss.py"
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "loss.py")
SUM = os.path.join(HERE, "ref_loss.sha256")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(update: bool = False) -> str | None:
    """Copy the reference loss file into oracle/_ref/ (when /root/reference exists) and verify its checksum.
    Returns the staged path, or None when neither the source nor a staged copy is available."""
    if os.path.exists(SRC):
        os.makedirs(DST_DIR, exist_ok=True)
        shutil.copyfile(SRC, DST)
        if update or not os.path.exists(SUM):
            with open(SUM, "w") as f:
                f.write(_sha(DST) + "  src/open_clip/loss.py\n")
    if not os.path.exists(DST):
        return None
    if os.path.exists(SUM):
        want = open(SUM).read().split()[0]
        got = _sha(DST)
        if got != want:
            raise RuntimeError(f"oracle/_ref/loss.py does not match oracle/ref_loss.sha256 ({got} != {want})")
    return DST


TREE_SRC = "/root/reference/src"
TREE_DST = os.path.join(DST_DIR, "src")
TREE_SUM = os.path.join(HERE, "ref_tree.sha256")
TREE_PINNED = ["open_clip/loss.py", "open_clip/factory.py", "open_clip/model.py", "open_clip/transformer.py",
               "open_clip_train/train.py", "open_clip_train/precision.py", "open_clip_train/distributed.py"]


def stage_tree(update: bool = False) -> str | None:
    """Copy src/open_clip and src/open_clip_train into oracle/_ref/src/ (when /root/reference exists) and verify
    the pinned files.  Returns the staged `src` directory (to put on sys.path) or None."""
    if os.path.isdir(TREE_SRC):
        for pkg in ("open_clip", "open_clip_train"):
            shutil.copytree(os.path.join(TREE_SRC, pkg), os.path.join(TREE_DST, pkg), dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        if update or not os.path.exists(TREE_SUM):
            with open(TREE_SUM, "w") as f:
                for rel in TREE_PINNED:
                    f.write(_sha(os.path.join(TREE_DST, rel)) + "  src/" + rel + "\n")
    if not os.path.isdir(os.path.join(TREE_DST, "open_clip_train")):
        return None
    if os.path.exists(TREE_SUM):
        for line in open(TREE_SUM):
            want, rel = line.split()
            got = _sha(os.path.join(DST_DIR, rel))
            if got != want:
                raise RuntimeError(f"oracle/_ref/{rel} does not match oracle/ref_tree.sha256")
    return TREE_DST


def load_reference():
    """The staged reference module (`ClipLossWithDINOEnhancements` etc.), or None if it was never staged."""
    path = stage()
    if path is None:
        return None
    name = "ref_open_clip_loss"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = stage(update="--update" in sys.argv)
    print("staged:", p, "sha256:", None if p is None else _sha(p))
    print("staged tree:", stage_tree(update="--update" in sys.argv))
