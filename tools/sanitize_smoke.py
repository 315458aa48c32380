"""Smoke-size pass over every kernel of the library, meant to run under compute-sanitizer
(one --tool per gpurun call):
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_smoke.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vlp_b200  # noqa
from vlp_b200 import functional as VF, retrieval as R
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
n, fi, ft, d = 600, 512, 312, 128
f_i = torch.relu(torch.randn(n, fi, generator=g, device=dev)).requires_grad_(True)
f_t = torch.randn(n, ft, generator=g, device=dev).requires_grad_(True)
w_i = (torch.randn(fi, d, generator=g, device=dev) * fi ** -0.5).requires_grad_(True)
w_t = (torch.randn(ft, d, generator=g, device=dev) * ft ** -0.5).requires_grad_(True)
ls = torch.tensor([math.log(1 / 0.07)], dtype=torch.float64, device=dev, requires_grad=True)
for single in (True, False):          # single-recompute backward, then the two-pass kernels
    VF.SINGLE_SWEEP = single
    for t in (f_i, f_t, w_i, w_t, ls):
        t.grad = None
    loss, il, tl, ie, te = VF.fused_clip_loss(f_i, f_t, w_i, w_t, ls)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and all(torch.isfinite(t.grad).all() for t in (f_i, f_t, w_i, w_t, ls))
    print("single-sweep" if single else "two-pass", "loss", float(loss))
rank = R.retrieval_ranks(ie, te)
top = R.retrieval_topk(ie, ie, 16)
torch.cuda.synchronize()
assert int(top[:, 0].eq(torch.arange(n, device=dev)).sum()) == n and int(rank.min()) >= 0
print("retrieval ok; SANITIZE SMOKE DONE")
