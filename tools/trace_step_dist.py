"""Kernel timeline of one sharded fwd+bwd step (rank 0), via torch.profiler. Run under torchrun."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vlp_b200
from vlp_b200 import functional as VF
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev); group = dist.group.WORLD
n, d = 32768, 512; b = n // world
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
I = torch.nn.functional.normalize(a).to(torch.bfloat16)[rank*b:(rank+1)*b].contiguous()
T = torch.nn.functional.normalize(0.35*a+0.9368*c).to(torch.bfloat16)[rank*b:(rank+1)*b].contiguous()
ls = torch.tensor([math.log(1/0.07)], device=dev, requires_grad=True)
def step():
    Ii = I.detach().requires_grad_(True); Ti = T.detach().requires_grad_(True); ls.grad = None
    loss, _, _ = VF.fused_clip_loss_from_embeddings(Ii, Ti, ls, group=group); loss.backward()
for _ in range(6): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    # print the middle step only
    starts = [i for i, e in enumerate(evs) if "lse_partial" in e.name]
    lo = starts[1]; hi = starts[2] if len(starts) > 2 else len(evs)
    base = evs[lo].time_range.start
    for e in evs[lo:hi]:
        if e.name.startswith("nccl:"): continue
        print(f"{(e.time_range.start-base):9.1f} us  +{e.time_range.elapsed_us():8.1f}  {e.name[:60]}")
    print("step span us:", evs[hi-1].time_range.end - base if hi-1 < len(evs) else None)
VF.release_graphs()
if world > 1: dist.destroy_process_group()
