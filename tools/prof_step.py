"""Two fused fwd+bwd steps at global batch N (default 32768) x 512 -- the command profiled by ncu."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vlp_b200
from vlp_b200 import functional as VF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
I = torch.nn.functional.normalize(a).to(torch.bfloat16)
T = torch.nn.functional.normalize(0.35 * a + 0.9368 * c).to(torch.bfloat16)
ls = torch.tensor([math.log(1 / 0.07)], device=dev, requires_grad=True)
for _ in range(steps):
    Ii = I.detach().requires_grad_(True); Ti = T.detach().requires_grad_(True)
    loss, _, _ = VF.fused_clip_loss_from_embeddings(Ii, Ti, ls)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
