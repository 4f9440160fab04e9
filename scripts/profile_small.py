"""Where does a small-batch step go?  (torch.profiler, B=4096: CPU-side vs GPU-side time per step)"""
import sys, time, types, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dinosoft_b200 as pkg
from bench import synth, LOSS_ARGS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda")
img, txt, dino = synth(1, B, 512, 768, dev)
args = types.SimpleNamespace(**LOSS_ARGS)
loss = pkg.ClipLossWithDINOEnhancements()
torch.manual_seed(0)
loss.init_proj(512, 768, dev, "mlp")
scale = torch.tensor(14.2857, device=dev, requires_grad=True)
img.requires_grad_(True); txt.requires_grad_(True)
params = list(loss.image_to_dino_proj.parameters())

def step():
    img.grad = txt.grad = scale.grad = None
    for p in params: p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = loss(img, txt, scale, dino, args, output_dict=True)
    out["total_loss"].backward()

for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
t_cpu = (time.perf_counter() - t0) / 50
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 50
print(f"B={B}: CPU enqueue time per step {t_cpu*1e3:.3f} ms, wall per step {t_all*1e3:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
