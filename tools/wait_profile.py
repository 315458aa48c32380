"""Where does the backward pipeline wait?  Builds csrc/*.cu with -DVLP_PROFILE_WAITS into
tools/libvlpclip_prof.so (the shipped library is untouched), runs one dI pass at N x D and prints,
per role of the SM pair, the share of the kernel's cycles spent blocked on each barrier.

    python tools/wait_profile.py build          (CPU box: cross-compile)
    python tools/wait_profile.py [N] [D]        (B200)
"""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vlp_b200  # noqa: E402,F401
from vlp_b200 import _build, _lib  # noqa: E402

PROF_LIB = os.path.join(ROOT, "tools", "libvlpclip_prof.so")
NAMES = ["P.tma  wait ring slot free", "P.mma  wait X staged", "P.mma  wait S buffer free",
         "P.mma  wait Y stage full", "P.smx  wait X block released", "P.smx  wait S tile ready",
         "P.smx  wait G slot free", "P.smx  wait staging barrier", "C.tma  wait ring slot free",
         "C.mma  wait accumulator flushed", "C.mma  wait G tile arrived", "C.mma  wait Y stage full",
         "P      kernel cycles", "C      kernel cycles"]


def build():
    cmd = [_build._nvcc()] + _build.NVCC_FLAGS + ["-DVLP_PROFILE_WAITS", "-o", PROF_LIB] + _build.sources()
    subprocess.run(cmd, check=True)
    print("built", PROF_LIB)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        return build()
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    _build.LIB_PATH = PROF_LIB
    _build.needs_build = lambda: False
    from vlp_b200 import functional as VF
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(n, d, generator=g, device=dev)
    c = torch.randn(n, d, generator=g, device=dev)
    I = torch.nn.functional.normalize(a).to(torch.bfloat16)
    T = torch.nn.functional.normalize(0.35 * a + 0.9368 * c).to(torch.bfloat16)
    s = math.exp(math.log(1 / 0.07))
    rm, rl, rdiag, cm, cl = VF.lse_stats_fused(I, T, s, 0)
    r_stats = VF.merge_stats(rm, rl, rdiag, s)[:3]
    c_stats = VF.merge_stats(cm, cl, rdiag, s)[:3]
    i16, t16 = VF.cast_bf16_to_f16(I), VF.cast_bf16_to_f16(T)
    for _ in range(2):
        VF._grad(i16, t16, r_stats, c_stats, s, 0, n, 1.0, 1.0, True)
    # [74][16] counters of the two-pass backward (grad_pair_kernel); the rest of the buffer is unused
    prof_all = torch.zeros(74 * 16 + 148 * 8, dtype=torch.int64, device=dev)
    prof = prof_all[:74 * 16].view(74, 16)
    lib.vlpclip_dev_set_wait_profile(prof_all.data_ptr())
    lib.vlpclip_time_grad_kernel(1)
    VF._grad(i16, t16, r_stats, c_stats, s, 0, n, 1.0, 1.0, True)
    ms = lib.vlpclip_last_grad_kernel_ms()
    torch.cuda.synchronize()
    lib.vlpclip_dev_set_wait_profile(None)
    p = prof.double().cpu()
    used = p[:, 12] > 0
    tiles = (n // 128) ** 2 / max(int(used.sum()), 1)
    print(f"N={n} D={d}: kernel {ms:.3f} ms, {int(used.sum())} SM pairs, {tiles:.0f} tiles per pair")
    tot_p, tot_c = p[used, 12].mean().item(), p[used, 13].mean().item()
    print(f"cycles per tile: producer {tot_p / tiles:.0f}, consumer {tot_c / tiles:.0f}")
    for i, name in enumerate(NAMES[:12]):
        tot = tot_p if name.startswith("P") else tot_c
        v = p[used, i].mean().item()
        print(f"  {name:34s} {v / tiles:8.0f} cyc/tile  {100 * v / tot:5.1f} %")


if __name__ == "__main__":
    main()
