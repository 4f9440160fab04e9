"""oracle/_ref/loss.py (the staged, unmodified reference loss file that bench.py times) is the file the golden
fixtures were generated from: checksum + one fixture re-evaluated through it."""
import importlib.util
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT


def _make_ref():
    spec = importlib.util.spec_from_file_location("dsoft_make_ref", os.path.join(ROOT, "oracle", "make_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_staged_reference_matches_checksum_and_golden():
    mk = _make_ref()
    ref = mk.load_reference()  # raises on a checksum mismatch
    if ref is None:
        pytest.skip("oracle/_ref not staged (no /root/reference on this machine)")
    z = np.load(os.path.join(GOLDEN_DIR, "w1_noproj_text.npz"), allow_pickle=True)
    img, txt, dino = (torch.from_numpy(z[k]).double() for k in ("image", "text", "dino"))
    args = types.SimpleNamespace(use_projection=False, lambda_soft=0.5, soft_mode="kl_teacher", soft_dino_to_text=True,
                                 text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15, lambda_original=1.0,
                                 lambda_weighted=0.0)
    loss = ref.ClipLossWithDINOEnhancements()
    out = loss(img, txt, torch.tensor(float(z["scale"]), dtype=torch.float64), dino, args, output_dict=True)
    assert float(out["classic_loss"]) == pytest.approx(float(z["f64_r0_classic_loss"]), rel=1e-12)
