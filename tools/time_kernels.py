"""Times the two hot kernels alone (C ABI) at N x D; prints ms and algorithmic/executed TFLOP/s."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vlp_b200
from vlp_b200 import functional as VF
dev = torch.device("cuda:0")
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8192, 32768]
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
s = 14.2857
for n in sizes:
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
    I = torch.nn.functional.normalize(a).to(torch.bfloat16)
    T = torch.nn.functional.normalize(0.35 * a + 0.9368 * c).to(torch.bfloat16)
    def timeit(fn, reps=5):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_f = timeit(lambda: VF.lse_stats(I, T, s, 0))
    t_ff = timeit(lambda: VF.lse_stats_fused(I, T, s, 0))
    rm, rl, rd = VF.lse_stats(I, T, s, 0); cm, cl, cd = VF.lse_stats(T, I, s, 0)
    rs = VF.merge_stats(rm, rl, rd, s)[:3]; cs = VF.merge_stats(cm, cl, cd, s)[:3]
    i16 = VF.cast_bf16_to_f16(I); t16 = VF.cast_bf16_to_f16(T)
    t_g = timeit(lambda: VF._grad(i16, t16, rs, cs, s, 0, n, 1.0, 1.0, True))
    fl = 2.0 * n * n * d
    print(f"N={n} D={d}: fused_fwd {t_ff:.3f} ms ({fl/t_ff/1e9:.0f} TF/s) lse_fwd {t_f:.3f} ms ({fl/t_f/1e9:.0f} TF/s)  grad {t_g:.3f} ms (alg {fl/t_g/1e9:.0f} / exec {2*fl/t_g/1e9:.0f} TF/s)"
          f"  => step est {t_ff+2*t_g:.3f} ms = {6*n*n*d/(t_ff+2*t_g)/1e9/1645.6*100:.1f}% of peak")
