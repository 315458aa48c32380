"""Developer check (GPU): fused loss fwd+bwd through vlp_b200.functional vs the fp64 closed form."""
import sys, os, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import vlp_b200
from vlp_b200 import functional as Fn
from oracle import clip_oracle as O

dev = torch.device("cuda:0")
bad = 0
for (n, d, ls, rho) in [(128, 64, 2.6593, 0.35), (256, 512, 2.6593, 0.35), (256, 512, 2.6593, 0.0), (256, 512, 5.0, 0.0), (256, 512, 5.0, 0.35),
                        (256, 512, math.log(50), 0.35), (256, 512, math.log(50), 0.0), (1, 64, 2.6593, 0.35), (2, 8, 2.6593, 0.35),
                        (200, 72, 2.6593, 0.35), (1000, 128, 3.0, 0.5), (2048, 256, 2.6593, 0.35), (4096, 512, 2.6593, 0.35), (4096, 512, 4.6, 0.35)]:
    I, T = O.make_embeddings(n, d, rho=rho, seed=42)
    ref = O.closed_form(I.numpy(), T.numpy(), ls)
    Ic = I.to(dev).requires_grad_(True); Tc = T.to(dev).requires_grad_(True)
    lsc = torch.tensor([ls], dtype=torch.float64, device=dev, requires_grad=True)
    loss, il, tl = Fn.fused_clip_loss_from_embeddings(Ic, Tc, lsc)
    loss.backward()
    torch.cuda.synchronize()
    e_loss = abs(loss.item() - ref["loss"]) / max(abs(ref["loss"]), 1e-30)
    e_il = abs(il.item() - ref["image_loss"]) / max(abs(ref["image_loss"]), 1e-30)
    e_dI = O.rel_err(Ic.grad.cpu().numpy(), ref["dI"]); e_dT = O.rel_err(Tc.grad.cpu().numpy(), ref["dT"])
    dl = lsc.grad.item()
    e_dl = abs(dl - ref["dlogit_scale"]) / max(abs(ref["dlogit_scale"]), 1e-30) if ref["dlogit_scale"] != 0 else abs(dl)
    ok = e_loss < 1e-4 and e_il < 1e-4 and e_dI < 1e-3 and e_dT < 1e-3 and e_dl < 1e-3
    bad += (not ok)
    print(f"n={n} d={d} ls={ls:.3f} rho={rho}: loss {loss.item():.6f} (ref {ref['loss']:.6f}) rel {e_loss:.2e} il {e_il:.1e} | dI {e_dI:.2e} dT {e_dT:.2e} dl {e_dl:.2e} ({dl:.3e}) {'OK' if ok else 'FAIL'}")
print("BWD CHECK", "PASS" if bad == 0 else f"FAIL ({bad})")

for n in (8192, 32768):
    d = 512
    I, T = O.make_embeddings(n, d, rho=0.35, seed=1)
    Ic = I.to(dev).to(torch.bfloat16).requires_grad_(True); Tc = T.to(dev).to(torch.bfloat16).requires_grad_(True)
    lsc = torch.tensor([2.6593], dtype=torch.float32, device=dev, requires_grad=True)
    def step():
        Ic.grad = None; Tc.grad = None; lsc.grad = None
        loss, _, _ = Fn.fused_clip_loss_from_embeddings(Ic, Tc, lsc)
        loss.backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"fwd+bwd n={n} d={d}: {ms:.3f} ms -> {n/ms*1e3/1e6:.2f} M pairs/s, algorithmic {6*n*n*d/ms/1e9:.1f} TFLOP/s ({6*n*n*d/ms/1e9/1645.6*100:.1f}% of 1645.6)")
