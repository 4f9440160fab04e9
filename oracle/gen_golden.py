"""Generate the golden fixtures under tests/golden/ by EXECUTING the reference implementation.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/gen_golden.py

The reference's ``src/open_clip/loss.py`` is loaded by file path (it only needs torch) and its
``ClipLossWithDINOEnhancements`` is run unmodified on seeded synthetic inputs, at world_size 1 (in
process) and world_size 2 (two spawned gloo processes).  Inputs, projection-head weights and every
output (loss terms, gradients) are stored as .npz so the oracle can be checked against them anywhere.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REF_LOSS = "/root/reference/src/open_clip/loss.py"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_open_clip_loss", REF_LOSS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_inputs(seed, B, D, Dd, clustered=True):
    """Clustered embeddings: bf16-representable values stored as fp32."""
    g = torch.Generator().manual_seed(seed)
    K = max(B // 8, 2)
    cid = torch.randint(0, K, (B,), generator=g)
    noise = 0.5 if clustered else 1.0

    def make(d):
        cent = torch.randn(K, d, generator=g)
        x = cent[cid] * (1.0 if clustered else 0.0) + noise * torch.randn(B, d, generator=g)
        return x

    img = torch.nn.functional.normalize(make(D), dim=-1)
    txt = torch.nn.functional.normalize(make(D), dim=-1)
    dino = make(Dd) * 3.0  # un-normalised, like real CLS tokens
    r = lambda x: x.to(torch.bfloat16).to(torch.float32)
    return r(img), r(txt), r(dino)


def head_params(module, projection_type, layernorm):
    p = {}
    if projection_type == "linear":
        p["w0"], p["b0"] = module.weight, module.bias
    else:
        p["w0"], p["b0"] = module[0].weight, module[0].bias
        p["w1"], p["b1"] = module[2].weight, module[2].bias
        if layernorm:
            p["ln_w"], p["ln_b"] = module[3].weight, module[3].bias
    return p


def run_reference(ref, rank, world, img, txt, dino, scale, args, local_loss, gather_with_grad, dtype, head_seed):
    """One rank's reference evaluation; returns dict of numpy outputs."""
    b = img.shape[0] // world
    rows = slice(rank * b, (rank + 1) * b)
    im = img[rows].to(dtype).clone().requires_grad_(True)
    tx = txt[rows].to(dtype).clone().requires_grad_(True)
    dn = None if dino is None else dino[rows].to(dtype)
    sc = torch.tensor(scale, dtype=dtype, requires_grad=True)
    loss = ref.ClipLossWithDINOEnhancements(
        local_loss=local_loss, gather_with_grad=gather_with_grad, cache_labels=False, rank=rank, world_size=world
    )
    out = {}
    use_proj = getattr(args, "use_projection", True) and dn is not None
    if use_proj:
        # build the lazily-created head deterministically (same weights on every rank) and in `dtype`
        torch.manual_seed(head_seed)
        loss.init_proj(img.shape[1], dino.shape[1], "cpu", getattr(args, "projection_type", "mlp"),
                       layernorm=getattr(args, "use_layernorm", False))
        loss.image_to_dino_proj = loss.image_to_dino_proj.to(dtype)
    res = loss(im, tx, sc, dn, args, output_dict=True)
    res["total_loss"].backward()
    for k in ("total_loss", "classic_loss", "soft_loss", "weighted_loss"):
        out[k] = np.asarray(float(res[k].detach()))
    out["d_image"] = im.grad.numpy().astype(np.float64)
    out["d_text"] = tx.grad.numpy().astype(np.float64)
    out["d_logit_scale"] = np.asarray(float(sc.grad))
    if use_proj:
        hp = head_params(loss.image_to_dino_proj, getattr(args, "projection_type", "mlp"),
                         getattr(args, "use_layernorm", False))
        for k, v in hp.items():
            out["head_" + k] = v.detach().numpy().astype(np.float64)
            if v.grad is not None:
                out["dhead_" + k] = v.grad.numpy().astype(np.float64)
    return out


def _worker(rank, world, port, payload, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    ref = load_reference()
    out = run_reference(ref, rank, world, **payload)
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


CASES = [
    # name, B, D, Dd, scale, world, flags, args
    dict(name="w1_noproj_text", B=96, D=64, Dd=128, scale=14.2857, world=1, clustered=True,
         args=dict(use_projection=False, lambda_soft=0.5, soft_mode="kl_teacher", soft_dino_to_text=True,
                   text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15)),
    dict(name="w1_mlp_text_scale100", B=64, D=64, Dd=96, scale=100.0, world=1, clustered=True,
         args=dict(use_projection=True, projection_type="mlp", lambda_soft=0.5, soft_mode="kl_teacher",
                   soft_dino_to_text=True, text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15)),
    dict(name="w1_linear_notext", B=80, D=64, Dd=64, scale=30.0, world=1, clustered=False,
         args=dict(use_projection=True, projection_type="linear", lambda_soft=0.25, soft_mode="kl_teacher",
                   soft_dino_to_text=False, teacher_temp=0.1)),
    dict(name="w1_mlp_layernorm", B=48, D=32, Dd=64, scale=20.0, world=1, clustered=True,
         args=dict(use_projection=True, projection_type="mlp", use_layernorm=True, lambda_soft=1.0,
                   soft_mode="kl_teacher", soft_dino_to_text=True, text_lambda=0.2, text_student_temp=0.05,
                   teacher_temp=0.15, lambda_original=0.7)),
    dict(name="w1_classic_only", B=72, D=64, Dd=64, scale=14.2857, world=1, clustered=True,
         args=dict(use_projection=True, lambda_soft=0.0, soft_mode="none")),
    dict(name="w1_scale_below_10", B=40, D=32, Dd=32, scale=4.0, world=1, clustered=False,
         args=dict(use_projection=False, lambda_soft=0.5, soft_mode="kl_teacher", teacher_temp=0.15)),
    dict(name="w1_residual", B=56, D=64, Dd=64, scale=25.0, world=1, clustered=True,
         args=dict(use_projection=True, projection_type="mlp", residual_projection=True, lambda_soft=0.5,
                   soft_mode="kl_teacher", soft_dino_to_text=True, text_lambda=0.3, text_student_temp=0.03,
                   teacher_temp=0.15)),
    dict(name="w1_residual_alpha", B=56, D=64, Dd=64, scale=25.0, world=1, clustered=True,
         args=dict(use_projection=True, projection_type="linear", residual_projection=True, residual_alpha=0.3,
                   lambda_soft=0.5, soft_mode="kl_teacher", teacher_temp=0.2)),
    # denominator-modulated CE branch (loss.py:416-471): alone, and with every other term switched on
    dict(name="w1_weighted", B=64, D=64, Dd=96, scale=14.2857, world=1, clustered=True,
         args=dict(use_projection=False, lambda_soft=0.0, soft_mode="none", lambda_weighted=0.5, rho=0.1,
                   c_clip=1.0, weight_text_symmetry=False)),
    dict(name="w1_weighted_sym_all_terms", B=80, D=64, Dd=96, scale=40.0, world=1, clustered=True,
         args=dict(use_projection=True, projection_type="mlp", lambda_soft=0.5, soft_mode="kl_teacher",
                   soft_dino_to_text=True, text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15,
                   lambda_weighted=0.7, rho=0.2, c_clip=0.5, weight_text_symmetry=True)),
    dict(name="w2_gather_grad", B=64, D=64, Dd=96, scale=14.2857, world=2, clustered=True,
         local_loss=True, gather_with_grad=True,
         args=dict(use_projection=True, projection_type="mlp", lambda_soft=0.5, soft_mode="kl_teacher",
                   soft_dino_to_text=True, text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15)),
    dict(name="w2_no_gather_grad", B=64, D=64, Dd=96, scale=50.0, world=2, clustered=True,
         local_loss=True, gather_with_grad=False,
         args=dict(use_projection=False, lambda_soft=0.5, soft_mode="kl_teacher", soft_dino_to_text=True,
                   text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15)),
]


# input / head seeds per fixture: the position the case had when its fixture was first generated
SEED_INDEX = {"w1_noproj_text": 0, "w1_mlp_text_scale100": 1, "w1_linear_notext": 2, "w1_mlp_layernorm": 3,
              "w1_classic_only": 4, "w1_scale_below_10": 5, "w1_residual": 6, "w1_residual_alpha": 7,
              "w2_gather_grad": 8, "w2_no_gather_grad": 9, "w1_weighted": 10, "w1_weighted_sym_all_terms": 11}


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    ref = load_reference()
    port = 29611
    only = set(sys.argv[1:])  # optional: fixture names to (re)generate; seeds depend on the case NAME, not order
    for case in CASES:
        if only and case["name"] not in only:
            continue
        ci = SEED_INDEX[case["name"]]
        img, txt, dino = synth_inputs(1234 + ci, case["B"], case["D"], case["Dd"], case["clustered"])
        args = types.SimpleNamespace(**case["args"])
        world = case["world"]
        save = dict(image=img.numpy().copy(), text=txt.numpy().copy(), dino=dino.numpy().copy(), scale=np.asarray(case["scale"]),
                    world=np.asarray(world), local_loss=np.asarray(case.get("local_loss", False)),
                    gather_with_grad=np.asarray(case.get("gather_with_grad", False)))
        for k, v in case["args"].items():
            save["arg_" + k] = np.asarray(v)
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            payload = dict(img=img.clone(), txt=txt.clone(), dino=dino.clone(), scale=case["scale"], args=args,
                           local_loss=case.get("local_loss", False),
                           gather_with_grad=case.get("gather_with_grad", False), dtype=dtype, head_seed=99 + ci)
            if world == 1:
                outs = [run_reference(ref, 0, 1, **payload)]
            else:
                mgr = mp.Manager()
                ret = mgr.dict()
                port += 1
                mp.spawn(_worker, args=(world, port, payload, ret), nprocs=world, join=True)
                outs = [ret[r] for r in range(world)]
            for r, o in enumerate(outs):
                for k, v in o.items():
                    if k.startswith("head_") and (tag != "f64" or r != 0):
                        continue  # weights once
                    save[f"{tag}_r{r}_{k}" if not k.startswith("head_") else k] = v
        path = os.path.join(OUT_DIR, case["name"] + ".npz")
        np.savez_compressed(path, **save)
        print(f"{case['name']:28s} total(f64,r0)={float(save['f64_r0_total_loss']):.6f} -> {os.path.relpath(path)}")


if __name__ == "__main__":
    if not os.path.exists(REF_LOSS):
        sys.exit("reference not mounted at /root/reference - fixtures can only be generated in the build container")
    main()
