"""Drop-in proof at the reference's own call site: the UNMODIFIED `create_loss` (src/open_clip/factory.py:506-588)
and `train_one_epoch` (src/open_clip_train/train.py:145-586) run with this repo's loss class installed.

Runs where the reference is mounted (/root/reference; skipped elsewhere, e.g. on the GPU box).  On this CPU-only
box the kernels are replaced by the oracle-backed test double, so what is exercised is exactly the boundary:
constructor keywords, forward keywords, the `args` namespace produced by `make_effective_args`, the returned
dict that the loop `.item()`s, backward through a real optimizer step."""
import os
import sys
import types

import pytest
import torch

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference not mounted")


def _import_reference_train():
    # optional third-party modules the reference imports at module scope but this path never uses (SURVEY probe table)
    if "ftfy" not in sys.modules:
        m = types.ModuleType("ftfy")
        m.fix_text = lambda s: s
        sys.modules["ftfy"] = m
    if "matplotlib" not in sys.modules:
        mp = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mp.pyplot = plt
        sys.modules["matplotlib"] = mp
        sys.modules["matplotlib.pyplot"] = plt
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import open_clip  # noqa: F401
    import open_clip_train.train as T

    return T


class TinyClip(torch.nn.Module):
    def __init__(self, din, vocab, d):
        super().__init__()
        self.visual = torch.nn.Linear(din, d)
        self.tok = torch.nn.Embedding(vocab, d)
        self.logit_scale = torch.nn.Parameter(torch.tensor(2.659))  # ln(1/0.07), model.py:324-325

    def forward(self, images, texts):
        im = torch.nn.functional.normalize(self.visual(images.flatten(1)), dim=-1)
        tx = torch.nn.functional.normalize(self.tok(texts).mean(1), dim=-1)
        return {"image_features": im, "text_features": tx, "logit_scale": self.logit_scale.exp()}


class Loader(list):
    num_batches = 0
    num_samples = 0


class DataInfo:
    def __init__(self, loader):
        self.dataloader = loader

    def set_epoch(self, epoch):
        self.epoch = epoch


def test_reference_train_loop_runs_with_dropin_loss(pkg, oracle):
    from oracle_backend import OracleBackend

    T = _import_reference_train()
    pkg.install_into_open_clip()
    from open_clip.factory import create_loss

    torch.manual_seed(0)
    bs, nb, din, vocab, d, dd, n_table = 16, 3, 24, 50, 32, 48, 64
    args = types.SimpleNamespace(
        # create_loss (factory.py:506-588)
        distill=False, model="ViT-B-32", siglip=False, use_CyClip=False, use_coca=False, use_dino_general=True,
        local_loss=False, gather_with_grad=False, rank=0, world_size=1, horovod=False,
        # train_one_epoch
        device="cpu", precision="fp32", accum_freq=1, skip_scheduler=True, grad_clip_norm=None, batch_size=bs,
        log_every_n_steps=1, local_rank=0, use_mlflow=False, warmup=0, enable_warmup_dino_hyperparams=False,
        _precomputed_dino=torch.randn(n_table, dd) * 3, _dino_on_device=False,
        # DINO-Soft knobs (params.py:58-203; thesis sweep values sweep_manual.sh:30-46)
        use_projection=True, projection_type="mlp", lambda_soft=0.5, soft_mode="kl_teacher", soft_dino_to_text=True,
        text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15, lambda_weighted=0.0, lambda_original=1.0,
    )
    loss = create_loss(args)
    assert type(loss).__module__.startswith("dinosoft_b200"), type(loss)
    loss._backend = OracleBackend(oracle)  # CPU box: kernels replaced by the oracle test double

    model = TinyClip(din, vocab, d)
    loader = Loader((torch.randn(bs, din), torch.randint(0, vocab, (bs, 5)), torch.randint(0, n_table, (bs,)))
                    for _ in range(nb))
    loader.num_batches, loader.num_samples = nb, nb * bs
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    before = [p.detach().clone() for p in model.parameters()]

    logs = T.train_one_epoch(model, {"train": DataInfo(loader)}, loss, 0, opt, None, None, None, None, None, args)

    assert len(logs) == nb
    for rec in logs:
        for k in ("loss/total_loss", "loss/classic_loss", "loss/soft_loss", "loss/weighted_loss"):
            assert k in rec and rec[k] == rec[k], (k, rec)  # present and not NaN
        assert rec["loss/soft_loss"] > 0 and rec["loss/weighted_loss"] == 0.0
        assert rec["loss/total_loss"] == pytest.approx(rec["loss/classic_loss"] + 0.5 * rec["loss/soft_loss"], rel=1e-5)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, model.parameters()))
    assert loss.image_to_dino_proj is not None  # lazily created head (loss.py:214-238)
    assert len(list(loss.parameters())) == 4    # ... registered as a sub-module, so an optimizer can own it
