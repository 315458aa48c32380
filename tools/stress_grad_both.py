"""Dev: repeat the single-recompute backward on one input and compare the runs element by element
(bit-reproducibility), and against the two-pass kernel (gross errors = stale tiles)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vlp_b200  # noqa
from vlp_b200 import functional as VF
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
g = torch.Generator(device=dev).manual_seed(7)
a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
I = torch.nn.functional.normalize(a).to(torch.bfloat16)
T = torch.nn.functional.normalize(0.35 * a + 0.9368 * c).to(torch.bfloat16)
del a, c
s = 1 / 0.07
rm, rl, rdiag, cm, cl = VF.lse_stats_fused(I, T, s, 0)
r, cs = VF.merge_stats(rm, rl, rdiag, s)[:3], VF.merge_stats(cm, cl, rdiag, s)[:3]
i16, t16 = VF.cast_bf16_to_f16(I), VF.cast_bf16_to_f16(T)
dI2, _ = VF._grad(i16, t16, r, cs, s, 0, n, 1.0, 1.0, False)
dT2, _ = VF._grad(t16, i16, cs, r, s, 0, n, 1.0, 1.0, False)
ref = None
bad = 0
for k in range(reps):
    # some unrelated traffic before the launch, like a real step (forward sweep + casts)
    VF.lse_stats_fused(I, T, s, 0)
    dI, dT, ds = VF._grad_both(i16, t16, r, cs, s, 0, n, 1.0, 1.0, True)
    torch.cuda.synchronize()
    eI = ((dI.double() - dI2.double()).norm() / dI2.double().norm()).item()
    eT = ((dT.double() - dT2.double()).norm() / dT2.double().norm()).item()
    msg = f"run {k}: vs two-pass dI {eI:.2e} dT {eT:.2e}"
    if ref is not None:
        for name, x, y in (("dI", dI, ref[0]), ("dT", dT, ref[1])):
            ne = (x != y)
            cnt = int(ne.sum())
            if cnt:
                bad += 1
                rows = ne.any(dim=1).nonzero().flatten()
                md = (x - y).abs().max().item()
                rel = ((x - y).abs() / y.abs().clamp_min(1e-30))[ne].max().item()
                msg += f" | {name} differs in {cnt} elems, {len(rows)} rows (first {rows[:6].tolist()}, blocks {sorted(set((rows // 128).tolist()))[:8]}), max abs {md:.3e} max rel {rel:.3e}"
        if ds.item() != ref[2].item():
            msg += f" | ds differs {ds.item()} vs {ref[2].item()}"
    else:
        ref = (dI.clone(), dT.clone(), ds.clone())
    print(msg, flush=True)
print("STRESS", "FAILED" if bad else "PASSED")
