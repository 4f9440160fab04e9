"""Isolated timing of the two collectives of one step (packed-feature all-gather, LSE all-gather) next to the
full forward / backward of the loss, at the bench workload.  Launch with torchrun like bench.py."""
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    import dinosoft_b200 as pkg

    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    B = 32768
    b = B // world
    row = 2 * 512 + 768 + 768
    g = torch.empty((B, row), dtype=torch.bfloat16, device=dev)
    lse = torch.empty((world, 5, b), dtype=torch.float32, device=dev)
    t_ag = timed(lambda: dist.all_gather_into_tensor(g.view(-1), g[rank * b:(rank + 1) * b].view(-1)))
    t_lse = timed(lambda: dist.all_gather_into_tensor(lse.view(-1), lse[rank].view(-1)))

    img, txt, dino = bench.synth(1234 + rank, b, bench.D_CLIP, bench.D_DINO, dev)
    larg = types.SimpleNamespace(**bench.LOSS_ARGS)
    loss = pkg.ClipLossWithDINOEnhancements(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
    torch.manual_seed(99)
    loss.init_proj(bench.D_CLIP, bench.D_DINO, dev, "mlp")
    scale = torch.tensor(14.2857, device=dev, requires_grad=True)
    img.requires_grad_(True); txt.requires_grad_(True)
    holder = {}

    def fwd():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            holder["o"] = loss(img, txt, scale, dino, larg, output_dict=True)

    def fwdbwd():
        img.grad = txt.grad = scale.grad = None
        for p in loss.image_to_dino_proj.parameters():
            p.grad = None
        fwd()
        holder["o"]["total_loss"].backward()

    t_step = timed(fwdbwd)
    with torch.no_grad():
        t_fwd_nograd = timed(fwd)
    if rank == 0:
        bytes_recv = (world - 1) * b * row * 2
        print(f"world={world}: feature all-gather {t_ag:.3f} ms ({bytes_recv / t_ag / 1e6:.0f} GB/s received per rank), "
              f"LSE all-gather {t_lse:.3f} ms, fwd+bwd {t_step:.3f} ms, no-grad forward {t_fwd_nograd:.3f} ms",
              flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
