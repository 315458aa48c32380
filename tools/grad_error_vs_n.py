"""Dev: gradient error of the fused head against the blocked fp64 reference (bench.parity_check) as a
function of the batch size, for the single-recompute and the two-pass backward."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import vlp_b200  # noqa
from vlp_b200 import functional as VF
import bench
dev = torch.device("cuda:0")
d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rho = float(sys.argv[2]) if len(sys.argv) > 2 else 0.35
for n in (2048, 4096, 8192, 16384, 32768):
    g = torch.Generator(device=dev).manual_seed(42)
    a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
    c = rho * a + math.sqrt(1 - rho ** 2) * c
    I = F.normalize(a).to(torch.bfloat16); T = F.normalize(c).to(torch.bfloat16)
    for single in (True, False):
        VF.SINGLE_SWEEP = single
        Ir = I.detach().requires_grad_(True); Tr = T.detach().requires_grad_(True)
        ls = torch.tensor([bench.LOGIT_SCALE], device=dev, requires_grad=True)
        loss, _, _ = VF.fused_clip_loss_from_embeddings(Ir, Tr, ls)
        loss.backward()
        torch.cuda.synchronize()
        p = bench.parity_check(torch, I, T, bench.LOGIT_SCALE, 0, 1, loss.detach(), Ir.grad, Tr.grad, ls.grad, n_sample=64)
        print(f"n={n:6d} {'single-sweep' if single else 'two-pass    '} loss {p['loss_rel_err']:.1e} dI {p['dI_rel_err']:.2e} "
              f"dT {p['dT_rel_err']:.2e} dls {p['dlogit_scale_rel_err']:.1e}", flush=True)
