"""BASELINE config 5: the reference's own training step with the B200 loss dropped in.

Drives the UNMODIFIED `open_clip.create_model` (ViT-B-32, random init), `open_clip.factory.create_loss` and
`open_clip_train.train.train_one_epoch` (train.py:145-586) from the staged reference tree (`oracle/_ref/src`,
see oracle/make_ref.py; /root/reference/src where it is mounted) with a synthetic loader: 224 px images, 77-token
captions, sample indices into a precomputed DINOv2 feature table (the reference trains on precomputed DINO CLS
features, main.py:693-741).  One process per GPU (torchrun for N > 1, DDP over NCCL as train.py expects).

  python scripts/train_step_harness.py --loss ours      --batch 256 --steps 8
  python scripts/train_step_harness.py --loss reference --batch 256 --steps 8
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \\
      scripts/train_step_harness.py --loss ours --batch 512

Prints ONE JSON line (rank 0): samples/s of the whole step (data already on the device, as the loss bench), the
per-step time, the last logged loss terms.  `--loss reference` keeps the reference's own loss class: the two runs
differ in nothing but the loss object that create_loss returns.  TEST / BENCH INFRASTRUCTURE - not product code.
"""
import argparse
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_src():
    for cand in (os.path.join(ROOT, "oracle", "_ref", "src"), "/root/reference/src"):
        if os.path.isdir(os.path.join(cand, "open_clip_train")):
            return cand
    return None


def import_reference():
    """open_clip + open_clip_train.train from the staged tree; optional third-party modules the reference imports at
    module scope but this path never uses are stubbed (SURVEY.md 8c probe table)."""
    src = reference_src()
    if src is None:
        raise SystemExit("reference tree not staged: run `python oracle/make_ref.py` where /root/reference exists")
    if "ftfy" not in sys.modules:
        try:
            import ftfy  # noqa: F401
        except ImportError:
            m = types.ModuleType("ftfy")
            m.fix_text = lambda s: s
            sys.modules["ftfy"] = m
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            mp = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mp.pyplot = plt
            sys.modules["matplotlib"] = mp
            sys.modules["matplotlib.pyplot"] = plt
    if src not in sys.path:
        sys.path.insert(0, src)
    import open_clip
    import open_clip_train.train as T

    return open_clip, T


class Loader(list):
    num_batches = 0
    num_samples = 0


class DataInfo:
    def __init__(self, loader):
        self.dataloader = loader

    def set_epoch(self, epoch):
        self.epoch = epoch


def build(a, dev, rank, world):
    open_clip, T = import_reference()
    import dinosoft_b200 as pkg

    if a.loss == "ours":
        pkg.install_into_open_clip()
    else:
        pkg.uninstall_from_open_clip()  # no-op unless an earlier run in this process installed the drop-in
    from open_clip.factory import create_loss

    torch.manual_seed(1234)  # same towers on every rank (DDP would broadcast rank 0's anyway)
    model = open_clip.create_model(a.model, precision="fp32", device=dev, output_dict=True)
    dd = a.dino_dim
    n_table = max(4 * a.batch * world, 1024)
    g = torch.Generator().manual_seed(7)
    table = (3.0 * torch.randn(n_table, dd, generator=g)).pin_memory()  # CPU table, as main.py:697-698 loads it
    args = types.SimpleNamespace(
        # create_loss (factory.py:506-588)
        distill=False, model=a.model, siglip=False, use_CyClip=False, use_coca=False, use_dino_general=True,
        local_loss=True, gather_with_grad=True, rank=rank, world_size=world, horovod=False,
        # train_one_epoch (train.py:145-586)
        device=str(dev), precision=a.precision, accum_freq=1, skip_scheduler=True, grad_clip_norm=None,
        batch_size=a.batch, log_every_n_steps=getattr(a, "log_every", 10 ** 9), local_rank=dev.index, use_mlflow=False, warmup=0,
        enable_warmup_dino_hyperparams=False, _precomputed_dino=table, _dino_on_device=False,
        distributed=world > 1, val_frequency=0, epochs=1, wandb=False, save_logs=False,
        # DINO-Soft knobs (params.py:58-203; thesis sweep values sweep_manual.sh:30-46)
        use_projection=True, projection_type="mlp", lambda_soft=0.5, soft_mode="kl_teacher", soft_dino_to_text=True,
        text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15, lambda_weighted=0.0, lambda_original=1.0,
    )
    loss = create_loss(args)
    # the lazily created projection head draws from the global RNG at the first forward (loss.py:223-238)
    torch.manual_seed(99)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index])
    g = torch.Generator().manual_seed(100 + rank)
    nb = a.steps
    vocab = 49408
    loader = Loader(
        (torch.randn(a.batch, 3, a.image_size, a.image_size, generator=g).to(dev),
         torch.randint(1, vocab, (a.batch, 77), generator=g).to(dev),
         torch.randint(0, n_table, (a.batch,), generator=g))
        for _ in range(nb))
    loader.num_batches, loader.num_samples = nb, nb * a.batch * world
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
    return T, model, loss, opt, {"train": DataInfo(loader)}, args


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loss", choices=["ours", "reference"], default="ours")
    ap.add_argument("--model", default="ViT-B-32")
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--precision", default="amp_bf16")
    ap.add_argument("--dino-dim", type=int, default=768)
    ap.add_argument("--image-size", type=int, default=224)
    a = ap.parse_args()

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("train_step_harness.py needs a CUDA device")
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    T, model, loss, opt, data, args = build(a, dev, rank, world)

    def epoch(ep):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.time()
        logs = T.train_one_epoch(model, data, loss, ep, opt, None, None, None, None, None, args)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return time.time() - t0, logs

    epoch(0)  # warm-up: lazy head creation, cuDNN / cuBLAS heuristics, plan creation
    sec, logs = epoch(1)
    t = torch.tensor([sec], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t)
    if rank == 0:
        last = logs[-1] if logs else {}
        print(json.dumps({
            "metric": "train_step_samples_per_sec", "value": a.batch * world * a.steps / sec, "unit": "samples/s",
            "n_gpus": world, "steps": a.steps, "ms_per_step": 1e3 * sec / a.steps, "loss_impl": a.loss,
            "loss_class": type(loss).__module__ + "." + type(loss).__name__,
            "config": {"workload": "BASELINE config 5: reference train_one_epoch, %s towers (random init), precomputed "
                                   "DINOv2 features (dim %d) from a CPU table, synthetic %d px images, %s"
                                   % (a.model, a.dino_dim, a.image_size, a.precision),
                       "per_gpu_batch": a.batch, "global_batch": a.batch * world},
            "last_losses": {k: v for k, v in last.items() if k.startswith("loss/")},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
