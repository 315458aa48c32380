"""Dev: minimal driver for ncu captures of the prologue GEMMs at the head's shapes (32768 x 512/312 -> 512):
forward projection (A K-major, B MN-major), d features (both K-major), dW (both MN-major, split-K)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vlp_b200  # noqa
from vlp_b200 import functional as VF
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = torch.Generator(device=dev).manual_seed(0)
for f in (512, 312):
    d = 512
    feat = torch.randn(n, f, generator=g, device=dev)
    w = torch.randn(f, d, generator=g, device=dev) * f ** -0.5
    du = torch.randn(n, d, generator=g, device=dev)
    for _ in range(3):
        VF._gemm_tf32(feat, w, n, d, f, 0, 0)      # u = feat W
        VF._gemm_tf32(du, w, n, f, d, 0, 1)        # d feat = du W^T
        VF._gemm_tf32(feat, du, f, d, n, 1, 0)     # dW = feat^T du
torch.cuda.synchronize()
print("done")
