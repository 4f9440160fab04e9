"""world_size = 2 on CPU (gloo): the sharded host path of the module - packed all-gather, rank offsets,
LSE exchange call, gather_with_grad / soft_scope flags - against the reference's own 2-rank numbers
(tests/golden/w2_*.npz, produced by running the reference under gloo) and against the global oracle.
Compute is the oracle-backed test double; the collectives are the real torch.distributed calls."""
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN_DIR, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fname, soft_scope, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import importlib.util

    import dinosoft_b200 as pkg
    from oracle_backend import OracleBackend

    spec = importlib.util.spec_from_file_location("dinosoft_oracle", os.path.join(ROOT, "oracle", "dinosoft_oracle.py"))
    oracle = importlib.util.module_from_spec(spec)
    sys.modules["dinosoft_oracle"] = oracle
    spec.loader.exec_module(oracle)

    z = np.load(os.path.join(GOLDEN_DIR, fname))
    a = types.SimpleNamespace(**{k[4:]: z[k].item() for k in z.files if k.startswith("arg_")})
    img, txt, dino = (torch.from_numpy(z[k]).double() for k in ("image", "text", "dino"))
    b = img.shape[0] // world
    rows = slice(rank * b, (rank + 1) * b)
    m = pkg.ClipLossWithDINOEnhancements(local_loss=bool(z["local_loss"]), gather_with_grad=bool(z["gather_with_grad"]),
                                         rank=rank, world_size=world, soft_scope=soft_scope)
    m._backend = OracleBackend(oracle)
    if getattr(a, "use_projection", True):
        m.init_proj(img.shape[1], dino.shape[1], "cpu", getattr(a, "projection_type", "mlp"))
        m.image_to_dino_proj = m.image_to_dino_proj.double()
        sd = m.image_to_dino_proj.state_dict()
        mapping = {"0.weight": "w0", "0.bias": "b0", "2.weight": "w1", "2.bias": "b1"}
        m.image_to_dino_proj.load_state_dict({k: torch.from_numpy(z["head_" + mapping[k]]) for k in sd})
    im = img[rows].clone().requires_grad_(True)
    tx = txt[rows].clone().requires_grad_(True)
    sc = torch.tensor(float(z["scale"]), dtype=torch.float64, requires_grad=True)
    out = m(im, tx, sc, dino[rows], a, output_dict=True)
    out["total_loss"].backward()
    ret[rank] = dict(total=float(out["total_loss"].detach()), classic=float(out["classic_loss"].detach()),
                     soft=float(out["soft_loss"].detach()), d_image=im.grad.numpy(), d_text=tx.grad.numpy(),
                     d_scale=float(sc.grad), calls=list(m._backend.calls))
    dist.barrier()
    dist.destroy_process_group()


def _run(fname, soft_scope):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), fname, soft_scope, ret), nprocs=2, join=True)
    return [ret[r] for r in range(2)]


@pytest.mark.parametrize("fname,tol", [("w2_no_gather_grad.npz", 1e-6), ("w2_gather_grad.npz", 5e-3)])
def test_two_rank_reference_literal(fname, tol):
    """soft_scope='local' reproduces what the reference computes at world_size 2 (oracle (1))."""
    z = np.load(os.path.join(GOLDEN_DIR, fname))
    outs = _run(fname, "local")
    for r, o in enumerate(outs):
        assert o["calls"] == ["pack", "forward", "backward"]
        for k, key in (("total", "total_loss"), ("classic", "classic_loss"), ("soft", "soft_loss")):
            assert o[k] == pytest.approx(float(z[f"f64_r{r}_{key}"]), rel=tol, abs=1e-9), (r, k)
        for k in ("d_image", "d_text"):
            ref = z[f"f64_r{r}_{k}"]
            err = np.abs(o[k] - ref).max() / np.abs(ref).max()
            assert err < max(4 * tol, 1e-6), (r, k, err)
        assert o["d_scale"] == pytest.approx(float(z[f"f64_r{r}_d_logit_scale"]), rel=max(tol, 1e-6), abs=1e-9)


def test_two_rank_global_scope_matches_world1(oracle):
    """soft_scope='global' (north_star): the mean over ranks of every loss term equals the world_size-1 loss
    on the concatenated batch, and each rank's feature gradient is W x the global gradient's slice."""
    z = np.load(os.path.join(GOLDEN_DIR, "w2_gather_grad.npz"))
    outs = _run("w2_gather_grad.npz", "global")
    a = {k[4:]: z[k].item() for k in z.files if k.startswith("arg_")}
    img, txt, dino = (torch.from_numpy(z[k]).double() for k in ("image", "text", "dino"))
    head = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("head_")}
    cfg = oracle.OracleConfig(lambda_soft=a["lambda_soft"], soft_mode=a["soft_mode"], teacher_temp=a["teacher_temp"],
                              soft_dino_to_text=a["soft_dino_to_text"], text_lambda=a["text_lambda"],
                              text_student_temp=a["text_student_temp"], world_size=1, round_student_bf16=True)
    ref = oracle.loss_and_grads(img, txt, float(z["scale"]), dino, cfg, proj_params=head)["ranks"][0]
    mean_total = np.mean([o["total"] for o in outs])
    assert mean_total == pytest.approx(ref["total_loss"], rel=1e-6)  # the module's buffers are fp32
    b = img.shape[0] // 2
    for r, o in enumerate(outs):
        want = 2.0 * ref["d_image"][r * b:(r + 1) * b].numpy()
        assert np.abs(o["d_image"] - want).max() / np.abs(want).max() < 1e-6
        want = 2.0 * ref["d_text"][r * b:(r + 1) * b].numpy()
        assert np.abs(o["d_text"] - want).max() / np.abs(want).max() < 1e-6
    assert sum(o["d_scale"] for o in outs) == pytest.approx(2.0 * ref["d_logit_scale"], rel=1e-6)
