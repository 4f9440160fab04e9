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
