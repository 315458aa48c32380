"""Minimal driver for ncu captures of the backward kernels: statistics once, then a few launches.
    python tools/prof_grad_both.py [N] [D] [both|pair|fwd]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vlp_b200  # noqa
from vlp_b200 import functional as VF
sys.path.insert(0, os.path.join(ROOT, "tools"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
what = sys.argv[3] if len(sys.argv) > 3 else "both"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
I = torch.nn.functional.normalize(a).to(torch.bfloat16)
T = torch.nn.functional.normalize(0.35 * a + 0.9368 * c).to(torch.bfloat16)
s = 1 / 0.07
for _ in range(4 if what == "fwd" else 1):
    rm, rl, rdiag, cm, cl = VF.lse_stats_fused(I, T, s, 0)
r, cst = VF.merge_stats(rm, rl, rdiag, s), VF.merge_stats(cm, cl, rdiag, s)
i16, t16 = VF.cast_bf16_to_f16(I), VF.cast_bf16_to_f16(T)
for _ in range(4):
    if what == "both":
        VF._grad_both(i16, t16, r[:3], cst[:3], s, 0, n, 1.0, 1.0, True)
    elif what == "pair":
        VF._grad(i16, t16, r[:3], cst[:3], s, 0, n, 1.0, 1.0, True)
torch.cuda.synchronize()
print("done")
