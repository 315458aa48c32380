import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vlp_b200
from vlp_b200 import functional as VF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
I = torch.nn.functional.normalize(a).to(torch.bfloat16)
T = torch.nn.functional.normalize(0.35 * a + 0.9368 * c).to(torch.bfloat16)
for _ in range(3):
    VF.lse_stats_fused(I, T, 14.2857, 0)
torch.cuda.synchronize()
print("ok")
