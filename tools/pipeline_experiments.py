"""Which stage bounds the forward / backward pipelines?  One gpurun call answers it.

Builds the staged kernel versions csrc/next/{lse_fwd,grad_bwd}.cu (+ csrc/prologue.cu) several times
with experiment switches (see the comment blocks at the top of those files) into
tools/variants/libvlpclip_<name>.so -- the shipped library and its sources csrc/*.cu are never
touched -- and, on a B200, runs every variant in its own process (a hung mock cannot take the
others down): kernel times of the forward sweep (rows only / fused columns) and of one backward
pass, plus the blocked-cycle profile of every pipeline role (all variants carry
-DVLP_PROFILE_WAITS, and -DVLP_WAIT_WATCHDOG: a barrier wait of more than ~1 s traps instead of
hanging the GPU).  Variants marked "real" are also checked against a torch fp32 reference;
"mock" variants compute garbage by construction and are timing-only.

    python tools/pipeline_experiments.py build                 (cross-compile all variants; compile check)
    python tools/pipeline_experiments.py run [N] [D]           (B200; builds what is missing -- tools/variants/
                                                                does not travel with gpurun -- and writes
                                                                gpurun_out/pipeline_experiments.txt)
    python tools/pipeline_experiments.py one <name> [N] [D]    (B200; one variant, this process)
    VARIANTS=base,bwd_quad python tools/pipeline_experiments.py run   (a subset)
"""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vlp_b200  # noqa: E402,F401
from vlp_b200 import _build, _lib  # noqa: E402

VAR_DIR = os.path.join(ROOT, "tools", "variants")
# name -> (kind, -D switches, what the number means)
VARIANTS = {
    "base": ("real", [], "staged kernels without switches (= shipped algorithms) + wait counters"),
    "bwd_pingpong": ("real", ["VLP_BWD_PINGPONG"],
                     "softmax warp groups alternate whole tiles (2 tile times per tile and group)"),
    "fwd_pingpong": ("real", ["VLP_FWD_PINGPONG"],
                     "forward softmax warps as two groups of 8 alternating whole tiles (two 32-column passes each)"),
    "fwd_pair_pingpong": ("real", ["VLP_FWD_PAIR", "VLP_FWD_PINGPONG"], "both forward variants together"),
    "fwd_pair": ("real", ["VLP_FWD_PAIR"],
                 "forward on CTA pairs with cta_group::2 MMAs: each SM stages half of every Y tile"),
    "bwd_quad": ("real", ["VLP_BWD_QUAD"],
                 "backward on clusters of 4 CTAs with cta_group::2 MMAs: every SM stages half of each Y tile "
                 "(its cycles/tile are per 4 SMs: equal throughput = HALF the pair kernel's figure)"),
    "bwd_quad_push4": ("real", ["VLP_BWD_QUAD", "VLP_PUSH_SPLIT=4"], "quad backward + split push"),
    "half_y_fwd": ("mock", ["VLP_EXP_HALF_Y_F"], "forward streams half of Y per SM (cta_group::2 traffic)"),
    "half_y_bwd_p": ("mock", ["VLP_EXP_HALF_Y_P"], "backward producer streams half of Y"),
    "half_y_bwd_c": ("mock", ["VLP_EXP_HALF_Y_C"], "backward consumer streams half of Y"),
    "half_y_bwd_pc": ("mock", ["VLP_EXP_HALF_Y_P", "VLP_EXP_HALF_Y_C"], "both backward roles stream half of Y"),
    "fine_rings": ("real", ["VLP_P_KB_PER_STAGE=1", "VLP_C_Q_PER_STAGE=32", "VLP_FWD_KB_PER_STAGE=1"],
                   "16 KB ring stages everywhere (8 / 8 / 12 stages): fewer bytes pinned under the MMAs"),
    "fine_p": ("real", ["VLP_P_KB_PER_STAGE=1"], "16 KB stages in the backward producer's ring only (8 stages)"),
    "fine_c": ("real", ["VLP_C_Q_PER_STAGE=32"], "16 KB stages in the backward consumer's ring only (8 stages)"),
    "fine_rings_pingpong": ("real", ["VLP_P_KB_PER_STAGE=1", "VLP_C_Q_PER_STAGE=32", "VLP_FWD_KB_PER_STAGE=1",
                                     "VLP_BWD_PINGPONG"], "both real variants together"),
    "push4": ("real", ["VLP_PUSH_SPLIT=4"], "G tile pushed to the consumer as 4 concurrent 8 KB bulk copies"),
    "g3": ("real", ["VLP_G_SLOTS=3"], "three G slots (consumer ring 3 x 32 KB): looser producer/consumer coupling"),
    "g3_push4_fine": ("real", ["VLP_G_SLOTS=3", "VLP_PUSH_SPLIT=4", "VLP_P_KB_PER_STAGE=1", "VLP_C_Q_PER_STAGE=32",
                               "VLP_FWD_KB_PER_STAGE=1"], "three G slots + split push + 16 KB ring stages"),
    "epi8": ("real", ["VLP_EPI_WARPS=8"], "all 8 non-issuing consumer warps flush the accumulator (half the columns each)"),
    "x_unroll": ("real", ["VLP_X_UNROLL"], "X block of a segment / work item fetched with all loads in flight (d = 512)"),
    "x_unroll_epi8": ("real", ["VLP_X_UNROLL", "VLP_EPI_WARPS=8"], "both per-segment savings together"),
    "no_smx": ("mock", ["VLP_EXP_NO_SMX", "VLP_EXP_NO_SMX_F"], "softmax arithmetic removed (fwd + bwd)"),
    "bwd_decouple": ("mock", ["VLP_EXP_DECOUPLE"], "no G hand-off: each backward role at its own pace"),
    "bwd_decouple_half_y": ("mock", ["VLP_EXP_DECOUPLE", "VLP_EXP_HALF_Y_P", "VLP_EXP_HALF_Y_C"],
                            "own pace + half of Y: the pure MMA-issue bound of each role"),
}
BWD_NAMES = ["P.tma  wait ring slot free", "P.mma  wait X staged", "P.mma  wait S buffer free",
             "P.mma  wait Y stage full", "P.smx  wait X block released", "P.smx  wait S tile ready",
             "P.smx  wait G slot free", "P.smx  wait staging barrier", "C.tma  wait ring slot free",
             "C.mma  wait accumulator flushed", "C.mma  wait G tile arrived", "C.mma  wait Y stage full"]
RECT_ROWS = (4096, 8192)   # rows per rank at 8 / 4 GPUs of the 32768-pair job
FWD_NAMES = ["F.tma  wait ring slot free", "F.mma  wait X staged", "F.mma  wait S buffer free",
             "F.mma  wait Y stage full", "F.smx  wait X block released", "F.smx  wait S tile ready",
             "F.smx  wait column barrier"]


def lib_of(name):
    return os.path.join(VAR_DIR, f"libvlpclip_{name}.so")


def next_sources():
    nxt = os.path.join(_build.CSRC, "next")
    return [os.path.join(nxt, "lse_fwd.cu"), os.path.join(nxt, "grad_bwd.cu"),
            os.path.join(_build.CSRC, "prologue.cu")]


def build(only_missing=False, names=None):
    """Cross-compile every variant (one nvcc process per core)."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(VAR_DIR, exist_ok=True)

    def one_build(name):
        defs = VARIANTS[name][1]
        cmd = [_build._nvcc()] + _build.NVCC_FLAGS + ["-DVLP_PROFILE_WAITS", "-DVLP_WAIT_WATCHDOG"] + [f"-D{d}" for d in defs] + \
              ["-o", lib_of(name)] + next_sources()
        subprocess.run(cmd, check=True)
        return name

    todo = [n for n in (names or VARIANTS) if n in VARIANTS and not (only_missing and os.path.exists(lib_of(n)))]
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo) or 1, os.cpu_count() or 8))) as ex:
        for name in ex.map(one_build, todo):
            print("built", lib_of(name), flush=True)


def _reference(I, T, s):
    """torch fp32 restatement of the head on the bf16-rounded embeddings (loss + dI, dT, ds)."""
    import torch
    i = I.float().requires_grad_(True)
    t = T.float().requires_grad_(True)
    sc = torch.tensor(s, device=I.device, requires_grad=True)
    logits = (i @ t.T) * sc
    lab = torch.arange(I.shape[0], device=I.device)
    loss = 0.5 * (torch.nn.functional.cross_entropy(logits, lab) +
                  torch.nn.functional.cross_entropy(logits.T, lab))
    loss.backward()
    return loss.detach(), i.grad, t.grad, sc.grad


def one(name, n, d):
    import torch
    _build.LIB_PATH = lib_of(name)
    _build.needs_build = lambda: False
    from vlp_b200 import functional as VF
    lib = _lib.load()
    kind = VARIANTS[name][0]
    dev = torch.device("cuda:0")
    s = math.exp(math.log(1 / 0.07))

    def make(nn):
        g = torch.Generator(device=dev).manual_seed(0)
        a = torch.randn(nn, d, generator=g, device=dev)
        c = torch.randn(nn, d, generator=g, device=dev)
        return (torch.nn.functional.normalize(a).to(torch.bfloat16),
                torch.nn.functional.normalize(0.35 * a + 0.9368 * c).to(torch.bfloat16))

    def stats(I, T):
        rm, rl, rdiag, cm, cl = VF.lse_stats_fused(I, T, s, 0)
        return VF.merge_stats(rm, rl, rdiag, s), VF.merge_stats(cm, cl, rdiag, s)

    if kind == "real":   # parity on a ragged small problem first
        nn = 1000
        I, T = make(nn)
        r, c = stats(I, T)
        loss = 0.5 * (r[3].double().mean() + c[3].double().mean())
        i16, t16 = VF.cast_bf16_to_f16(I), VF.cast_bf16_to_f16(T)
        dI, ds1 = VF._grad(i16, t16, r[:3], c[:3], s, 0, nn, 1.0, 1.0, True)
        dT, ds2 = VF._grad(t16, i16, c[:3], r[:3], s, 0, nn, 1.0, 1.0, True)
        torch.cuda.synchronize()
        ref_loss, ref_dI, ref_dT, ref_ds = _reference(I, T, s)
        rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()  # noqa: E731
        e_loss = abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
        e = (e_loss, rel(dI, ref_dI), rel(dT, ref_dT), abs(ds1.item() - ref_ds.item()) / abs(ref_ds.item()),
             abs(ds2.item() - ref_ds.item()) / abs(ref_ds.item()))
        ok = e[0] < 1e-4 and max(e[1:]) < 1e-3
        print(f"[{name}] parity n={nn}: loss {e[0]:.2e} dI {e[1]:.2e} dT {e[2]:.2e} ds {e[3]:.2e}/{e[4]:.2e}"
              f"  {'PASS' if ok else 'FAIL'}")

    I, T = make(n)
    prof = torch.zeros(74 * 16 + 148 * 8, dtype=torch.int64, device=dev)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    ms_rows = timed(lambda: VF.lse_stats(I, T, s, 0))
    ms_fused = timed(lambda: VF.lse_stats_fused(I, T, s, 0))
    r, c = stats(I, T)
    i16, t16 = VF.cast_bf16_to_f16(I), VF.cast_bf16_to_f16(T)
    for _ in range(2):
        VF._grad(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True)
    lib.vlpclip_time_grad_kernel(1)
    ms_grad = 1e9
    for _ in range(3):
        VF._grad(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True)
        ms_grad = min(ms_grad, lib.vlpclip_last_grad_kernel_ms())
    lib.vlpclip_time_grad_kernel(0)
    # one profiled launch of each kernel
    lib.vlpclip_dev_set_wait_profile(prof.data_ptr())
    VF.lse_stats_fused(I, T, s, 0)
    VF._grad(i16, t16, r[:3], c[:3], s, 0, n, 1.0, 1.0, True)
    torch.cuda.synchronize()
    lib.vlpclip_dev_set_wait_profile(None)
    print(f"[{name}] N={n} D={d}: fwd rows-only {ms_rows:.3f} ms, fwd fused {ms_fused:.3f} ms, "
          f"grad kernel {ms_grad:.3f} ms   ({VARIANTS[name][2]})")
    tiles_total = ((n + 127) // 128) ** 2
    b = prof[:74 * 16].view(74, 16).double().cpu()
    used = b[:, 12] > 0
    if used.any():
        tiles = tiles_total / int(used.sum())
        tp, tc = b[used, 12].mean().item(), b[used, 13].mean().item()
        print(f"[{name}]   bwd: {int(used.sum())} clusters, cycles/tile: producer {tp / tiles:.0f}, "
              f"consumer {tc / tiles:.0f}")
        for i, nm in enumerate(BWD_NAMES):
            print(f"[{name}]     {nm:34s} {b[used, i].mean().item() / tiles:8.0f} cyc/tile")
    f = prof[74 * 16:].view(148, 8).double().cpu()
    used = f[:, 7] > 0
    if used.any():
        tiles = tiles_total / int(used.sum())
        print(f"[{name}]   fwd (fused) cycles/tile (incl. wave quantisation): {f[used, 7].mean().item() / tiles:.0f}")
        for i, nm in enumerate(FWD_NAMES):
            print(f"[{name}]     {nm:34s} {f[used, i].mean().item() / tiles:8.0f} cyc/tile")

    # per-rank shapes of the sharded run on ONE GPU: R local rows against all n columns (rank 0's
    # block), i.e. the three sweeps one rank of an (n / R)-GPU job executes
    for R in RECT_ROWS:
        if R >= n:
            continue
        Il, il16 = I[:R], i16[:R]
        ms_f = timed(lambda: VF.lse_stats_fused(Il, T, s, 0))
        rm, rl, rdiag, cm, cl = VF.lse_stats_fused(Il, T, s, 0)
        cdiag = torch.zeros(n, dtype=torch.float32, device=dev)
        cdiag[:R] = rdiag
        rr, cc = VF.merge_stats(rm, rl, rdiag, s)[:3], VF.merge_stats(cm, cl, cdiag, s)[:3]
        lib.vlpclip_time_grad_kernel(1)
        ms_di = ms_dt = 1e9
        for _ in range(3):
            VF._grad(il16, t16, rr, cc, s, 0, n, 1.0, 1.0, True)
            ms_di = min(ms_di, lib.vlpclip_last_grad_kernel_ms())
            VF._grad(t16, il16, cc, rr, s, 0, n, 1.0, 1.0, True)
            ms_dt = min(ms_dt, lib.vlpclip_last_grad_kernel_ms())
        lib.vlpclip_time_grad_kernel(0)
        print(f"[{name}] rank shape {R} x {n} ({n // R} GPUs): fwd fused {ms_f * 1e3:.0f} us, "
              f"dI kernel {ms_di * 1e3:.0f} us, dT kernel {ms_dt * 1e3:.0f} us")


def run(n, d):
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    log = open(os.path.join(out_dir, "pipeline_experiments.txt"), "w")
    only = [v for v in os.environ.get("VARIANTS", "").split(",") if v]   # optional subset, e.g. VARIANTS=base,bwd_quad
    build(only_missing=True, names=only or None)   # tools/variants/ is gpurun-ignored: built on the box
    for name in (only or VARIANTS):
        if name not in VARIANTS:
            sys.stdout.write(f"[{name}] unknown variant\n")
            continue
        if not os.path.exists(lib_of(name)):
            line = f"[{name}] library missing: run `python tools/pipeline_experiments.py build` first\n"
        else:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", name, str(n), str(d)],
                                   capture_output=True, text=True, timeout=90)
                line = r.stdout + ("" if r.returncode == 0 else f"[{name}] rc={r.returncode}\n{r.stderr[-2000:]}\n")
            except subprocess.TimeoutExpired:
                line = f"[{name}] TIMEOUT (hung kernel?)\n"
        sys.stdout.write(line)
        log.write(line)
        log.flush()


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "run"
    if cmd == "build":
        build()
    elif cmd == "one":
        one(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 32768,
            int(sys.argv[4]) if len(sys.argv) > 4 else 512)
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 32768, int(sys.argv[3]) if len(sys.argv) > 3 else 512)
