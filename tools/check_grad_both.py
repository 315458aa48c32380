"""Dev check of the single-recompute backward (vlpclip_grad_both) on a B200.

    python tools/check_grad_both.py [N] [D]

1. parity on small / ragged shapes: dI, dT, dscale against the two-pass kernel (vlpclip_grad) and a torch
   fp32 reference; 2. bit-reproducibility (two runs); 3. kernel time at N x D next to the two-pass time."""
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vlp_b200  # noqa: E402,F401
from vlp_b200 import _build, _lib  # noqa: E402

PROF = "--prof" in sys.argv
if PROF:   # a second build of the library with per-role blocked-cycle counters
    sys.argv.remove("--prof")
    import subprocess
    prof_lib = os.path.join(ROOT, "tools", "libvlpclip_gbprof.so")
    srcs = _build.sources()
    deps = srcs + [os.path.join(_build.CSRC, f) for f in os.listdir(_build.CSRC) if f.endswith(".cuh")]
    if not os.path.exists(prof_lib) or any(os.path.getmtime(f) > os.path.getmtime(prof_lib) for f in deps):
        subprocess.run([_build._nvcc()] + _build.NVCC_FLAGS + ["-DVLP_PROFILE_WAITS", "-o", prof_lib] + srcs, check=True)
    if "--build-only" in sys.argv:
        sys.exit(0)
    _build.LIB_PATH = prof_lib
    _build.needs_build = lambda: False
TIME_ONLY = "--time-only" in sys.argv      # skip the parity part (diagnostic builds give wrong results)
if TIME_ONLY:
    sys.argv.remove("--time-only")
if "--lib" in sys.argv:                    # a library built by hand (tools/diag_build.sh)
    i = sys.argv.index("--lib")
    _build.LIB_PATH = os.path.abspath(sys.argv[i + 1])
    _build.needs_build = lambda: False
    del sys.argv[i:i + 2]
    PROF = PROF or "prof" in os.path.basename(_build.LIB_PATH)
from vlp_b200 import functional as VF  # noqa: E402

PROD_NAMES = {0: "P.tma wait ring stage free", 1: "P.mma wait X staged", 2: "P.mma wait S buffer free",
              3: "P.mma wait Y stage full", 4: "P.smx wait S tile ready", 6: "P.smx wait G slot stored",
              5: "P.smx tmem ld + release S (sect.)", 7: "P.smx proxy fence + arrive (sect.)",
              13: "P.smx softmax arithmetic (sect.)", 14: "P.mma issue region, 4 stages (sect.)",
              8: "P.st  wait G staged", 9: "P.st  poll ring slot fetched", 10: "P.st  store read smem",
              11: "P.st  store complete", 12: "P.st  publish"}
DI_NAMES = {0: "I.tma wait G slot free", 1: "I.tma fetch G (poll + issue)", 2: "I.tma wait ring stage free",
            3: "I.mma wait accumulator flushed", 4: "I.mma wait G tile landed", 5: "I.mma wait Y stage full",
            6: "I.epi wait accumulator full", 7: "I.epi flush"}
DT_NAMES = {0: "T.tma wait G slot free", 1: "T.tma fetch G (poll + issue)", 2: "T.tma wait ring stage free",
            3: "T.mma wait accumulator flushed", 4: "T.mma wait G tile landed", 5: "T.mma wait I stage full",
            6: "T.epi wait accumulator full", 7: "T.epi poll column turn", 8: "T.epi flush (TMA reduce-add)",
            9: "T.epi flush (TMA store)", 10: "T.epi flush (read-back, final rows)"}


def wait_profile(lib, run, n_tiles_total):
    prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda:0")
    lib.vlpclip_dev_set_wait_profile(prof.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.vlpclip_dev_set_wait_profile(None)
    b = prof.view(148, 16).double().cpu()
    n_sms = lib.vlpclip_sm_count()
    npc = n_sms // 3
    env = int(os.environ.get("VLP_B200_GB_NP", "0"))
    if 0 < env < npc:
        npc = env
    tiles = n_tiles_total / npc
    prod, di, dt = b[0:npc], b[npc:2 * npc], b[2 * npc:n_sms]
    print(f"  cycles per tile-step ({npc} producers, {tiles:.0f} tiles each): producer {prod[:, 15].mean() / tiles:.0f}, "
          f"dI consumer {di[:, 15].mean() / tiles:.0f}, dT consumer {dt[:, 15].mean() / tiles:.0f}")
    for names, blk in ((PROD_NAMES, prod), (DI_NAMES, di), (DT_NAMES, dt)):
        for i, nm in names.items():
            print(f"    {nm:38s} {blk[:, i].mean() / tiles:8.0f} cyc/tile   (max CTA {blk[:, i].max() / tiles:8.0f})")


def make(n, d, dev, rho=0.35, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = torch.randn(n, d, generator=g, device=dev)
    c = torch.randn(n, d, generator=g, device=dev)
    return (torch.nn.functional.normalize(a).to(torch.bfloat16),
            torch.nn.functional.normalize(rho * a + math.sqrt(1 - rho * rho) * c).to(torch.bfloat16))


def reference(I, T, s, lo=0):
    i = I.float().requires_grad_(True)
    t = T.float().requires_grad_(True)
    sc = torch.tensor(s, device=I.device, requires_grad=True)
    logits = (i @ t.T) * sc
    n = T.shape[0]
    rows = torch.arange(I.shape[0], device=I.device) + lo
    # rectangular block of the global loss: rows lo..lo+n_loc of the row direction, all columns partially
    loss_r = torch.nn.functional.cross_entropy(logits, rows, reduction="sum") / n
    return loss_r, i, t, sc, logits


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def stats(I, T, s):
    rm, rl, rdiag, cm, cl = VF.lse_stats_fused(I, T, s, 0)
    return VF.merge_stats(rm, rl, rdiag, s), VF.merge_stats(cm, cl, rdiag, s)


def parity(n, d, dev, s):
    I, T = make(n, d, dev)
    r, c = stats(I, T, s)
    i16, t16 = VF.cast_bf16_to_f16(I), VF.cast_bf16_to_f16(T)
    dI2, ds2 = VF._grad(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True)
    dT2, _ = VF._grad(t16, i16, c[:3], r[:3], s, 0, n, 1.0, 1.0, False)
    dI, dT, ds = VF._grad_both(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True)
    dIb, dTb, dsb = VF._grad_both(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True)
    torch.cuda.synchronize()
    i = I.float().requires_grad_(True)
    t = T.float().requires_grad_(True)
    sc = torch.tensor(s, device=dev, requires_grad=True)
    logits = (i @ t.T) * sc
    lab = torch.arange(n, device=dev)
    loss = 0.5 * (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.T, lab))
    loss.backward()
    e = (rel(dI, i.grad), rel(dT, t.grad), abs(ds.item() - sc.grad.item()) / max(abs(sc.grad.item()), 1e-30),
         rel(dI, dI2), rel(dT, dT2))
    repro = torch.equal(dI, dIb) and torch.equal(dT, dTb) and torch.equal(ds, dsb)
    ok = max(e[:3]) < 1e-3 and repro
    print(f"parity n={n:5d} d={d:3d}: dI {e[0]:.2e} dT {e[1]:.2e} ds {e[2]:.2e} | vs two-pass dI {e[3]:.2e} dT {e[4]:.2e}"
          f" | bit-reproducible {repro}  {'PASS' if ok else 'FAIL'}", flush=True)
    return ok


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    dev = torch.device("cuda:0")
    lib = _lib.load()
    s = 1 / 0.07
    ok = True
    for nn, dd in [(128, 64), (100, 72), (256, 512), (1000, 128), (2048, 256), (4096, 512), (6500, 512), (3000, 768)]:
        if TIME_ONLY:
            break
        ok = parity(nn, dd, dev, s) and ok
    I, T = make(n, d, dev)
    r, c = stats(I, T, s)
    i16, t16 = VF.cast_bf16_to_f16(I), VF.cast_bf16_to_f16(T)
    lib.vlpclip_time_grad_kernel(1)
    for _ in range(2):
        VF._grad_both(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True)
    best = 1e9
    for _ in range(5):
        VF._grad_both(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True)
        best = min(best, lib.vlpclip_last_grad_kernel_ms())
    two = 0.0
    for x, y, xs, ys in ((i16, t16, r, c), (t16, i16, c, r)):
        b = 1e9
        for _ in range(3):
            VF._grad(x, y, xs[:3], ys[:3], s, 0, n, 1.0, 1.0, True)
            b = min(b, lib.vlpclip_last_grad_kernel_ms())
        two += b
    lib.vlpclip_time_grad_kernel(0)
    flops = 6.0 * n * n * d
    print(f"N={n} D={d}: single-recompute backward {best:.3f} ms ({flops / best / 1e9:.0f} TF/s executed), "
          f"two-pass {two:.3f} ms ({8.0 * n * n * d / two / 1e9:.0f} TF/s executed)", flush=True)
    if PROF:
        wait_profile(lib, lambda: VF._grad_both(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True),
                     ((n + 127) // 128) ** 2)
    for R in (4096, 8192, 16384):
        if R >= n or TIME_ONLY:
            continue
        Il, il16 = I[:R], i16[:R]
        rm, rl, rdiag, cm, cl = VF.lse_stats_fused(Il, T, s, 0)
        cdiag = torch.zeros(n, dtype=torch.float32, device=dev)
        cdiag[:R] = rdiag
        rr, cc = VF.merge_stats(rm, rl, rdiag, s)[:3], VF.merge_stats(cm, cl, cdiag, s)[:3]
        lib.vlpclip_time_grad_kernel(1)
        b = 1e9
        for _ in range(3):
            VF._grad_both(il16, t16, rr, cc, s, 0, n, 1.0, 1.0, True)
            b = min(b, lib.vlpclip_last_grad_kernel_ms())
        lib.vlpclip_time_grad_kernel(0)
        print(f"rank shape {R} x {n} ({n // R} GPUs): single-recompute backward {b * 1e3:.0f} us", flush=True)
    print("CHECK", "PASSED" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
