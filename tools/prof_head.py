"""Dev: kernel-level time table (torch.profiler / CUPTI) of one full-head step: projection + L2-normalise
prologue, loss, and the whole backward, at N x 512 on one GPU.   python tools/prof_head.py [N]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vlp_b200  # noqa
from vlp_b200 import functional as VF
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d, F_I, F_T = 512, 512, 312
g = torch.Generator(device=dev).manual_seed(7)
f_img = torch.relu(torch.randn(n, F_I, generator=g, device=dev)).requires_grad_(True)
f_txt = torch.randn(n, F_T, generator=g, device=dev).requires_grad_(True)
w_img = (torch.randn(F_I, d, generator=g, device=dev) * F_I ** -0.5).requires_grad_(True)
w_txt = (torch.randn(F_T, d, generator=g, device=dev) * F_T ** -0.5).requires_grad_(True)
ls = torch.tensor([math.log(1 / 0.07)], device=dev, requires_grad=True)
def step():
    for t_ in (f_img, f_txt, w_img, w_txt, ls):
        t_.grad = None
    out = VF.fused_clip_loss(f_img, f_txt, w_img, w_txt, ls)
    out[0].backward()
for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{e.time_range.start - t0:9.1f} us  + {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")
print("step span us:", ev[-1].time_range.end - t0)
