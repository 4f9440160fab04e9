"""Kernel-level timeline summary of one bench step (torch.profiler / CUPTI; no nsys in this image).

Launch like bench.py (plain python for one GPU, torchrun for several); rank 0 prints, per kernel name, the
number of launches and the summed device time of the profiled steps, plus the host time the step loop takes
to ENQUEUE a step (a step whose enqueue time approaches its device time is launch bound).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
      scripts/profile_step.py --steps 5
"""
import argparse
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32768)
    a = ap.parse_args()
    import torch.distributed as dist
    import dinosoft_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    b = a.batch // world
    img, txt, dino = bench.synth(1234 + rank, b, bench.D_CLIP, bench.D_DINO, dev)
    larg = types.SimpleNamespace(**bench.LOSS_ARGS)
    loss = pkg.ClipLossWithDINOEnhancements(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
    torch.manual_seed(99)
    loss.init_proj(bench.D_CLIP, bench.D_DINO, dev, "mlp")
    scale = torch.tensor(14.2857, device=dev, requires_grad=True)
    img.requires_grad_(True)
    txt.requires_grad_(True)

    def step():
        img.grad = txt.grad = scale.grad = None
        for p in loss.image_to_dino_proj.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = loss(img, txt, scale, dino, larg, output_dict=True)
        out["total_loss"].backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    t_enq = (time.perf_counter() - t0) / a.steps
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / a.steps
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        print(f"world={world} b={b}: host enqueue {t_enq * 1e3:.3f} ms/step, wall {t_dev * 1e3:.3f} ms/step")
        rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0
                and e.device_type == torch.autograd.DeviceType.CUDA]
        rows.sort(key=lambda r: -r[2])
        tot = sum(r[2] for r in rows)
        print(f"device time of all kernels: {tot / a.steps / 1e3:.3f} ms/step")
        for k, c, t in rows[:40]:
            print(f"{t / a.steps:9.1f} us/step  x{c / a.steps:5.1f}  {k[:100]}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
