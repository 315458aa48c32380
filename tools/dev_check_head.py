"""Developer check (GPU): prologue (projection + normalise) and the full head incl. its backward."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import vlp_b200
from vlp_b200 import functional as Fn
from oracle import clip_oracle as O

dev = torch.device("cuda:0")
bad = 0
for (n, fi, ft, d, ls) in [(256, 512, 312, 512, 2.6593), (300, 512, 768, 128, 2.6593), (1024, 2048, 312, 256, 3.0), (4096, 512, 312, 512, 2.6593), (64, 512, 312, 32, 2.6593)]:
    f_i, f_t, w_i, w_t = O.make_features(n, fi, ft, d, seed=7)
    fic = f_i.to(dev).requires_grad_(True); ftc = f_t.to(dev).requires_grad_(True)
    wic = w_i.to(dev).requires_grad_(True); wtc = w_t.to(dev).requires_grad_(True)
    lsc = torch.tensor([ls], dtype=torch.float64, device=dev, requires_grad=True)
    loss, il, tl, ie, te = Fn.fused_clip_loss(fic, ftc, wic, wtc, lsc)
    loss.backward(); torch.cuda.synchronize()
    # 1. embeddings vs fp64 normalize
    Ei = torch.nn.functional.normalize(f_i.double() @ w_i.double()); Et = torch.nn.functional.normalize(f_t.double() @ w_t.double())
    e_emb = max((ie.detach().cpu().double() - Ei).abs().max().item(), (te.detach().cpu().double() - Et).abs().max().item())
    # 2. loss on the kernel's own bf16 embeddings
    ib = ie.detach().to(torch.bfloat16).float().cpu(); tb = te.detach().to(torch.bfloat16).float().cpu()
    ref = O.closed_form(ib.numpy(), tb.numpy(), ls)
    e_loss = abs(loss.item() - ref["loss"]) / abs(ref["loss"])
    # 3. straight-through backward oracle in fp64
    ui = (f_i.double() @ w_i.double()).numpy(); ut = (f_t.double() @ w_t.double()).numpy()
    dui = O.normalize_backward(ui, ref["dI"]); dut = O.normalize_backward(ut, ref["dT"])
    dwi = f_i.double().numpy().T @ dui; dwt = f_t.double().numpy().T @ dut
    dfi = dui @ w_i.double().numpy().T; dft = dut @ w_t.double().numpy().T
    errs = dict(dWi=O.rel_err(wic.grad.cpu().numpy(), dwi), dWt=O.rel_err(wtc.grad.cpu().numpy(), dwt),
                dfi=O.rel_err(fic.grad.cpu().numpy(), dfi), dft=O.rel_err(ftc.grad.cpu().numpy(), dft),
                dl=abs(lsc.grad.item() - ref["dlogit_scale"]) / max(abs(ref["dlogit_scale"]), 1e-30))
    ok = e_emb < 2e-3 and e_loss < 1e-4 and all(v < 2e-3 for v in errs.values())
    bad += (not ok)
    print(f"n={n} Fi={fi} Ft={ft} d={d}: emb maxabs {e_emb:.2e} | loss rel {e_loss:.2e} | " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()) + (" OK" if ok else " FAIL"))
print("HEAD CHECK", "PASS" if bad == 0 else f"FAIL ({bad})")
