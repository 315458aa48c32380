"""Sharded fwd+bwd under torchrun with progress prints (debugging the CUDA-graph + NCCL path)."""
import os, sys, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vlp_b200
from vlp_b200 import functional as VF
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev); group = dist.group.WORLD
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 512; b = n // world
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
I = torch.nn.functional.normalize(a).to(torch.bfloat16)[rank*b:(rank+1)*b].contiguous()
T = torch.nn.functional.normalize(0.35*a+0.9368*c).to(torch.bfloat16)[rank*b:(rank+1)*b].contiguous()
ls = torch.tensor([math.log(1/0.07)], device=dev, requires_grad=True)
for k in range(8):
    Ii = I.detach().requires_grad_(True); Ti = T.detach().requires_grad_(True); ls.grad = None
    t0 = time.time()
    loss, _, _ = VF.fused_clip_loss_from_embeddings(Ii, Ti, ls, group=group)
    print(f"[rank {rank}] step {k} forward enqueued {time.time()-t0:.3f}s", flush=True)
    loss.backward()
    torch.cuda.synchronize()
    print(f"[rank {rank}] step {k} done loss {loss.item():.6f} |dI| {Ii.grad.float().norm().item():.4e} dl {ls.grad.item():.4e}", flush=True)
dist.barrier(); torch.cuda.synchronize()
print(f"[rank {rank}] OK", flush=True)
VF.release_graphs()
dist.destroy_process_group()
print(f'[rank {rank}] destroyed', flush=True)
