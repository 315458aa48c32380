#!/usr/bin/env python
"""bench.py -- CLIP loss fwd+bwd throughput of the fused B200 head (and of the reference CPU path).

    python bench.py --gpus N --steps K --warmup W            # our arm (1 rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # reference torch ops on the host cores

One "step" = one forward + backward of the symmetric InfoNCE head on one synthetic batch of
L2-normalised bf16 embeddings (gradients w.r.t. both embedding streams and logit_scale).
Workload (N = 1 GPU): the north-star configuration, global batch 32768 x dim 512.  With G GPUs the
same global batch is row-sharded (32768 / G rows per rank, strong scaling); every rank still meets
all 32768 columns (all-gather of the text embeddings, all-gather of the column statistics,
reduce-scatter of dT).

Rank 0 prints ONE JSON line (see the driver contract in the task description):
  value      whole-job pairs/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public API with HOST (pinned) inputs: per step one H2D copy of
             that step's embeddings (prefetched on a side stream) and a D2H read of the loss
  roofline   dominant kernel (grad_both_kernel, the single-recompute backward): algorithmic
             4 (N/G) N D flops per launch / its measured duration vs the measured dense bf16 peak of
             MEASURED_PEAKS.json; `roofline_forward` is the same for the forward sweep
  full_head  the same step INCLUDING the projection + L2-normalise prologue and its backward
             (F_alg = 6 N^2 D + 6 N (F_i + F_t) D)
  parity     self-check OUTSIDE the timed region: loss against a blocked fp64 evaluation of the
             reference formula on the global batch, dI / dT of sampled rows and d logit_scale against
             fp64; the run exits non-zero when a bound (1e-4 loss, 1e-3 gradients) is exceeded
  cpu_baseline  the oracle port (reference torch ops, host cores) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CLIP loss fwd+bwd image-text pairs/sec"
UNIT = "pairs/s"
LOGIT_SCALE = math.log(1 / 0.07)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
PRIMING_STEPS = 3   # untimed steps before the W warm-up steps (see main): allocator steady state


class ClockSampler:
    """SM clock + throttle reasons sampled every ~10 ms DURING the timed region (NVML in a thread;
    falls back to one `nvidia-smi` query when pynvml is unavailable)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = False
        self._thread = None
        self.active = False      # polling runs from the warm-up on ...
        self.timed = False       # ... samples are kept only while the timed region runs

    def _run(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.gpu_index
        if vis:
            try:
                idx = int(vis.split(",")[self.gpu_index])
            except Exception:
                pass
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        while not self._stop:
            if not self.active:
                time.sleep(0.002)
                continue
            try:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.timed:      # (queries during the warm-up only take the first-call costs of NVML)
                    self.samples.append(mhz)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        import threading
        try:
            import pynvml  # noqa: F401
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self._thread is not None:
            self._stop = True
            self._thread.join(timeout=2)
        if not self.samples:
            try:
                r = subprocess.run(["nvidia-smi", f"--id={self.gpu_index}",
                                    "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=10).stdout.strip().split(", ")
                out["sm_mhz"], out["sm_max_mhz"], out["samples"] = float(r[0]), float(r[1]), 1
            except Exception:
                pass
            return out
        sm = sorted(self.samples)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_min_mhz"] = sm[0]
        out["sm_max_mhz"] = self.max_mhz
        out["samples"] = len(sm)
        out["reasons"] = sorted(self.reasons)
        return out


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_sample(n: int, d: int, block: int, steps: int, warmup: int, seed: int = 42):
    """Reference torch ops (oracle/clip_oracle.py, a restatement of VisionLanguageModule.py:456-459,
    533-552) on a row slab of the workload: `block` pairs meet all n columns.

    The full n x n problem needs ~5 live n x n fp64 buffers (43 GB at n = 32768), so each step runs
    a slab that does exactly block/n of the full step's arithmetic: S slab = s * I_blk @ T_all^T
    (fp32 GEMM, fp64 logits as in the reference), both soft-max passes over the slab, and autograd
    back to dI_blk and dT_all.  pairs/s = block / t_slab.
    """
    import torch
    import torch.nn.functional as F
    from oracle import clip_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(seed)
    t_all = F.normalize(torch.randn(n, d, generator=g)).to(torch.bfloat16).float().requires_grad_(True)
    i_blk = F.normalize(torch.randn(block, d, generator=g)).to(torch.bfloat16).float().requires_grad_(True)
    ls = torch.tensor([LOGIT_SCALE], dtype=torch.float64, requires_grad=True)   # reference :111
    labels = torch.arange(block)

    def step():
        for t in (t_all, i_blk, ls):
            t.grad = None
        scale = torch.clamp(ls.exp(), max=O.LOGIT_SCALE_MAX)                 # :456-457
        logits = (i_blk @ t_all.T) * scale                                   # :459 (slab of it)
        image_loss = F.cross_entropy(logits, labels, reduction="sum")        # :550 on the slab rows
        # column direction: same element-wise work on the same slab (log-softmax along dim 0)
        text_part = -torch.log_softmax(logits, dim=0)[labels, labels].sum()  # :551 (slab share)
        loss = (image_loss + text_part) / (2 * n)                            # :552
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    return {"value": block / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "ms_per_step": dt * 1e3,
            "sample": (f"row slab of the workload: {block} pairs x all {n} columns, d={d} "
                       f"(= {block}/{n} of one full fwd+bwd step; reference torch ops, fp32 GEMM + "
                       "fp64 soft-max as in VisionLanguageModule.py:456-459,550-552)")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, d = args.n, args.d
    block = args.cpu_block
    res = cpu_reference_sample(n, d, block, max(1, args.steps), max(0, args.warmup))
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32/f64",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(n, d, max(1, args.gpus)),
            "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"],
                             "kind": res["kind"], "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)
    return 0


def workload_config(n, d, world):
    """The `config` object of the JSON line: identical for our arm and the reference arm."""
    b = n // world
    return {"workload": f"fused contrastive loss fwd+bwd, global batch {n} x dim {d}",
            "global_batch": n, "dim": d, "rows_per_gpu": b, "logit_scale": LOGIT_SCALE,
            "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
            "l2": "two input sets rotated; inputs+operand copies+gradients = "
                  f"{(2 * n * d * 4 + 2 * 2 * n * d * 2 + 2 * n * d * 4) / 1e6:.0f} MB per step > 126 MB L2",
            "untimed_steps": f"{PRIMING_STEPS} allocator-priming steps + the W warm-up steps, same loop as the timed K"}


# --------------------------------------------------------------------------------------------
# self-check: blocked fp64 evaluation of the reference formula on the global batch
# --------------------------------------------------------------------------------------------
def parity_check(torch, I_loc, T_loc, ls_value, rank, world, loss, dI_loc, dT_loc, dls, n_sample=32):
    """loss = (CE(S) + CE(S^T)) / 2 with S = min(e^l, 100) I T^T  (VisionLanguageModule.py:456-459,
    550-552) evaluated in fp64, block by block, on the embeddings the kernels consumed; gradients
    from the closed form G = (P_row + P_col - 2 Id) / (2 N): rows of dI = s G T and dT = s G^T I
    for `n_sample` rows of this rank, d logit_scale = s * sum(G o C) in full."""
    dist = torch.distributed if world > 1 else None
    dev = I_loc.device
    b, d = I_loc.shape
    n = b * world
    lo = rank * b
    if world > 1:
        I_all = torch.empty(n, d, dtype=I_loc.dtype, device=dev)
        T_all = torch.empty(n, d, dtype=T_loc.dtype, device=dev)
        dist.all_gather_into_tensor(I_all, I_loc.contiguous())
        dist.all_gather_into_tensor(T_all, T_loc.contiguous())
    else:
        I_all, T_all = I_loc, T_loc
    e = math.exp(ls_value)
    s = min(e, 100.0)
    Td = T_all.double()
    blk = 1024
    row_lse = torch.empty(b, dtype=torch.float64, device=dev)
    diag = torch.empty(b, dtype=torch.float64, device=dev)
    col_m = torch.full((n,), -float("inf"), dtype=torch.float64, device=dev)
    col_s = torch.zeros(n, dtype=torch.float64, device=dev)
    for r0 in range(0, b, blk):
        S = s * (I_loc[r0:r0 + blk].double() @ Td.T)
        row_lse[r0:r0 + blk] = torch.logsumexp(S, dim=1)
        idx = torch.arange(r0, min(b, r0 + blk), device=dev)
        diag[r0:r0 + blk] = S[idx - r0, lo + idx]
        m_new = torch.maximum(col_m, S.max(dim=0).values)
        col_s = col_s * torch.exp(col_m - m_new) + torch.exp(S - m_new[None, :]).sum(dim=0)
        col_m = m_new
        del S
    if world > 1:
        m_glob = col_m.clone()
        dist.all_reduce(m_glob, op=dist.ReduceOp.MAX)
        col_s = col_s * torch.exp(col_m - m_glob)
        dist.all_reduce(col_s)
        col_m = m_glob
        row_lse_all = torch.empty(n, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(row_lse_all, row_lse)
    else:
        row_lse_all = row_lse
    col_lse = col_m + torch.log(col_s)
    sums = torch.stack([(row_lse - diag).sum(), (col_lse[lo:lo + b] - diag).sum()])
    if world > 1:
        dist.all_reduce(sums)
    il, tl = (sums / n).tolist()
    ref_loss = 0.5 * (il + tl)
    # d logit_scale: second blocked pass with the global column statistics
    ds = torch.zeros((), dtype=torch.float64, device=dev)
    for r0 in range(0, b, blk):
        C = I_loc[r0:r0 + blk].double() @ Td.T
        S = s * C
        G = torch.exp(S - row_lse[r0:r0 + blk, None]) + torch.exp(S - col_lse[None, :])
        idx = torch.arange(r0, min(b, r0 + blk), device=dev)
        G[idx - r0, lo + idx] -= 2.0
        ds += (G * C).sum() / (2.0 * n)
        del C, S, G
    if world > 1:
        dist.all_reduce(ds)
    ref_dls = float(ds) * (e if e <= 100.0 else 0.0)
    # sampled gradient rows of this rank
    g = torch.Generator().manual_seed(1234 + rank)
    rows = torch.randperm(b, generator=g)[:min(n_sample, b)].to(dev)
    S = s * (I_loc[rows].double() @ Td.T)                          # rows of S owned by the images
    G = (torch.exp(S - row_lse[rows, None]) + torch.exp(S - col_lse[None, :]))
    G[torch.arange(len(rows), device=dev), lo + rows] -= 2.0
    ref_dI = s * (G / (2.0 * n)) @ Td
    St = s * (T_loc[rows].double() @ I_all.double().T)            # columns of S owned by the texts
    Gt = (torch.exp(St - row_lse_all[None, :]) + torch.exp(St - col_lse[lo + rows, None]))
    Gt[torch.arange(len(rows), device=dev), lo + rows] -= 2.0
    ref_dT = s * (Gt / (2.0 * n)) @ I_all.double()
    rel = lambda a, r_: float((a.double() - r_).norm() / r_.norm().clamp_min(1e-300))  # noqa: E731
    errs = torch.tensor([abs(float(loss) - ref_loss) / abs(ref_loss), rel(dI_loc[rows], ref_dI),
                         rel(dT_loc[rows], ref_dT)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    e_loss, e_di, e_dt = errs.tolist()
    e_dls = abs(float(dls) - ref_dls) / max(abs(ref_dls), 1e-300) if ref_dls != 0.0 else abs(float(dls))
    ok = e_loss <= 1e-4 and e_di <= 1e-3 and e_dt <= 1e-3 and e_dls <= 1e-3
    return {"ok": bool(ok), "loss": float(loss), "loss_ref_fp64": ref_loss, "loss_rel_err": e_loss,
            "dI_rel_err": e_di, "dT_rel_err": e_dt, "dlogit_scale_rel_err": e_dls,
            "sampled_rows_per_rank": int(len(rows)), "bounds": {"loss": 1e-4, "grads": 1e-3},
            "reference": "blocked fp64 evaluation of VisionLanguageModule.py:456-459,550-552 and its closed-form "
                         "gradient on the global batch (max over ranks)"}


def cpu_full_step(timeout_s=300):
    """BASELINE.json configs[0]: the reference's ResNet34 + TinyBERT CLIP step (stock torch ops,
    batch 32, synthetic 224 x 224 radiographs + 32-token captions) on THIS box's host cores, in the
    same run (tools/full_step.py cpu, a subprocess so that its thread settings stay its own)."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "full_step.py"), "cpu", "--batch", "32",
                            "--dim", "128", "--steps", "2"], capture_output=True, text=True, timeout=timeout_s)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (r.stderr or "no output")[-200:]}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"[:200]}


def torch_gpu_baseline(torch, I, T, ls_value, steps=2):
    """Stock PyTorch on the same B200: the reference's own ops (fp32 GEMM, fp64 logits, two
    F.cross_entropy, autograd) on the whole batch -- the on-box bar SURVEY.md section 2 names."""
    import torch.nn.functional as F
    n = I.shape[0]
    try:
        i = I.float().requires_grad_(True)
        t = T.float().requires_grad_(True)
        ls = torch.tensor([ls_value], dtype=torch.float64, device=I.device, requires_grad=True)
        labels = torch.arange(n, device=I.device)

        def step():
            i.grad = t.grad = ls.grad = None
            logits = (i @ t.T) * torch.clamp(ls.exp(), max=100)
            loss = (F.cross_entropy(logits, labels) + F.cross_entropy(logits.T, labels)) / 2
            loss.backward()
            return loss
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        peak = torch.cuda.max_memory_allocated() / 2 ** 30
        return {"value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "peak_mem_gib": peak,
                "what": "reference torch ops (fp32 GEMM, fp64 logits, 2 x F.cross_entropy, autograd) on the same GPU, "
                        f"batch {n}"}
    except RuntimeError as exc:     # out of memory at this batch size
        return {"value": None, "error": str(exc)[:200]}
    finally:
        torch.cuda.empty_cache()


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.nn.functional as F
    import vlp_b200  # noqa: F401
    from vlp_b200 import _lib, functional as VF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fused head has no CPU path "
                         "(use --impl reference for the host baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib = _lib.load()
    n, d = args.n, args.d
    if n % world != 0:
        raise SystemExit(f"global batch {n} not divisible by {world} ranks")
    b = n // world

    # two rotating input sets (same distribution, different seeds); every rank builds its slice.
    # fp32 leaf tensors holding bf16-representable values -- the parity surface of the tests: the
    # kernels consume them exactly, and autograd returns fp32 gradients (bf16 leaves would get bf16
    # gradients, whose rounding alone is 1.7e-3 normwise: outside the 1e-3 bound)
    def make_inputs(seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        a = torch.randn(n, d, generator=g, device=dev)
        c = torch.randn(n, d, generator=g, device=dev)
        c = 0.35 * a + math.sqrt(1 - 0.35 ** 2) * c
        I = F.normalize(a).to(torch.bfloat16)[rank * b:(rank + 1) * b].float().contiguous()
        T = F.normalize(c).to(torch.bfloat16)[rank * b:(rank + 1) * b].float().contiguous()
        return I, T

    sets = [make_inputs(42), make_inputs(43)]
    ls = torch.tensor([LOGIT_SCALE], dtype=torch.float32, device=dev, requires_grad=True)

    def step(I, T):
        I = I.detach().requires_grad_(True)
        T = T.detach().requires_grad_(True)
        ls.grad = None
        loss, _, _ = VF.fused_clip_loss_from_embeddings(I, T, ls, group=group)
        loss.backward()
        return loss, I.grad, T.grad

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    # (the NVML thread is started BEFORE the warm-up: nvmlInit contends with kernel launches for
    #  driver locks and must not fall into the timed region; it only samples while `active`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # the warm-up runs exactly like the timed loop (same live tensors, same 2-step run-ahead), so that
    # the caching allocator has reached its steady state: a warm-up that drops the returned gradients
    # leaves a cudaMalloc of 2 x 64 MB for the second timed step (seen as one 5-11 ms step)
    # (3 priming steps come before the W warm-up steps: the fourth step of a process is the first one the
    #  host enqueues two steps ahead of the GPU with every buffer of the pattern alive, and the allocator
    #  answers it with fresh cudaMallocs -- 5-18 ms when it falls into the timed region, as it did with W = 3)
    wmarks = []
    sampler.active = True
    for w in range(PRIMING_STEPS + max(3, args.warmup)):
        loss, dI_last, dT_last = step(*sets[w % 2])
        wmarks.append(torch.cuda.Event())
        wmarks[-1].record()
        if w >= 2:
            wmarks[w - 2].synchronize()
    sync_all()
    sampler.timed = True
    launches0 = lib.vlpclip_launch_count() + VF.GRAPH_REPLAYED_LAUNCHES
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    # the host stays at most 2 steps ahead of the GPU (as a training loop that reads its loss
    # would): an unbounded run-ahead makes the caching allocator cudaMalloc new blocks inside the
    # timed region
    marks = []
    e0.record()
    for k in range(args.steps):
        loss, dI_last, dT_last = step(*sets[k % 2])
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
        if k >= 2:
            marks[k - 2].synchronize()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    if os.environ.get("VLP_BENCH_STEP_TIMES"):   # dev: per-step device times of the timed region
        ts = [e0.elapsed_time(marks[0])] + [marks[i - 1].elapsed_time(marks[i]) for i in range(1, len(marks))]
        print("step ms:", " ".join(f"{t:.2f}" for t in ts), file=sys.stderr)
    last_set = (args.steps - 1) % 2
    dls_last = ls.grad.detach().clone()
    sampler.timed = False
    sampler.active = False
    launches = lib.vlpclip_launch_count() + VF.GRAPH_REPLAYED_LAUNCHES - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n / (ms_per_step * 1e-3)

    # ---- self-check of the last timed step (outside the timed region) ----
    parity = parity_check(torch, sets[last_set][0], sets[last_set][1], LOGIT_SCALE, rank, world,
                          loss.detach(), dI_last, dT_last, dls_last)

    # ---- dominant kernels alone (through the C ABI, on the stream they run on) ----
    I, T = sets[0][0].to(torch.bfloat16), sets[0][1].to(torch.bfloat16)
    if world > 1:
        T_all = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
        torch.distributed.all_gather_into_tensor(T_all, T)
    else:
        T_all = T
    scale = math.exp(LOGIT_SCALE)
    reps = max(3, min(10, args.steps))
    # forward sweep (lse_partial_kernel + its two merge kernels)
    for _ in range(2):
        rm, rl, rdiag, cm, cl = VF.lse_stats_fused(I, T_all, scale, -rank * b)
    torch.cuda.synchronize()
    k0 = torch.cuda.Event(enable_timing=True)
    k1 = torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(reps):
        VF.lse_stats_fused(I, T_all, scale, -rank * b)
    k1.record()
    torch.cuda.synchronize()
    fwd_ms = k0.elapsed_time(k1) / reps
    r_stats = VF.merge_stats(rm, rl, rdiag, scale)[:3]
    if world > 1:
        cm_all = torch.stack(_all_gather_list(cm, world))
        cl_all = torch.stack(_all_gather_list(cl, world))
        cdiag = torch.cat(_all_gather_list(rdiag.contiguous(), world))
        c_stats = VF.merge_stats(cm_all, cl_all, cdiag, scale)[:3]
    else:
        c_stats = VF.merge_stats(cm, cl, rdiag, scale)[:3]
    i16 = VF.cast_bf16_to_f16(I)
    t16 = VF.cast_bf16_to_f16(T_all)
    grad_call = lambda: VF._grad_both(i16, t16, r_stats, c_stats, scale, -rank * b, n, 1.0, 1.0, True)  # noqa: E731
    for _ in range(2):
        grad_call()
    torch.cuda.synchronize()
    k0.record()
    for _ in range(reps):
        grad_call()
    k1.record()
    torch.cuda.synchronize()
    grad_call_ms = k0.elapsed_time(k1) / reps       # kernel + its helper launches
    # the dominant kernel alone: CUDA events recorded by the library around the grad_both_kernel
    # launch itself, on the stream it runs on
    grad_ms = grad_call_ms
    lib.vlpclip_time_grad_kernel(1)
    acc, cnt = 0.0, 0
    for _ in range(reps):
        grad_call()
        ms_k = float(lib.vlpclip_last_grad_kernel_ms())
        if ms_k > 0:
            acc, cnt = acc + ms_k, cnt + 1
    lib.vlpclip_time_grad_kernel(0)
    if cnt:
        grad_ms = acc / cnt

    # ---- the whole head: projection + L2-normalise prologue, loss, and all of its backward ----
    F_I, F_T = 512, 312        # ResNet34 / TinyBERT feature widths (VisionLanguageModule.py:102-109)
    gfe = torch.Generator(device=dev).manual_seed(7 + rank)
    f_img = torch.relu(torch.randn(b, F_I, generator=gfe, device=dev)).requires_grad_(True)
    f_txt = torch.randn(b, F_T, generator=gfe, device=dev).requires_grad_(True)
    gw = torch.Generator(device=dev).manual_seed(11)
    w_img = (torch.randn(F_I, d, generator=gw, device=dev) * F_I ** -0.5).requires_grad_(True)
    w_txt = (torch.randn(F_T, d, generator=gw, device=dev) * F_T ** -0.5).requires_grad_(True)

    def head_step():
        for t_ in (f_img, f_txt, w_img, w_txt, ls):
            t_.grad = None
        out = VF.fused_clip_loss(f_img, f_txt, w_img, w_txt, ls, group=group)
        out[0].backward()
        return out[0]

    for _ in range(3):
        head_step()
    sync_all()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hsteps = max(3, min(10, args.steps))
    h0.record()
    for k in range(hsteps):
        head_step()
    h1.record()
    sync_all()
    head_ms = h0.elapsed_time(h1) / hsteps
    if world > 1:
        t = torch.tensor([head_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        head_ms = float(t.item())

    # ---- weak-scaling point: 4096 rows per GPU, global batch 4096 * G ----
    bw = 4096
    gws = torch.Generator(device=dev).manual_seed(100 + rank)
    Iw = F.normalize(torch.randn(bw, d, generator=gws, device=dev)).to(torch.bfloat16).float()
    Tw = F.normalize(torch.randn(bw, d, generator=gws, device=dev)).to(torch.bfloat16).float()
    for _ in range(3):
        step(Iw, Tw)
    sync_all()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    for _ in range(10):
        step(Iw, Tw)
    w1.record()
    sync_all()
    weak_ms = w0.elapsed_time(w1) / 10
    if world > 1:
        t = torch.tensor([weak_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        weak_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host embeddings -> H2D (prefetched on a side stream) -> loss D2H ----
    host = [(I_.cpu().pin_memory(), T_.cpu().pin_memory()) for (I_, T_) in sets]
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [(torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def prefetch(k):
        slot = k % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            dev_bufs[slot][0].copy_(host[k % 2][0], non_blocking=True)
            dev_bufs[slot][1].copy_(host[k % 2][1], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(steps):
        for s_ in range(2):
            consumed[s_].record(torch.cuda.current_stream())
        prefetch(0)
        for k in range(steps):
            if k + 1 < steps:
                prefetch(k + 1)
            slot = k % 2
            torch.cuda.current_stream().wait_event(ready[slot])
            loss, _, _ = step(*dev_bufs[slot])
            consumed[slot].record(torch.cuda.current_stream())
            loss_host.copy_(loss.detach().reshape(1), non_blocking=False)   # D2H read of the result
        return float(loss_host.item())

    e2e_loop(PRIMING_STEPS + 3)
    sync_all()
    t0 = time.perf_counter()
    g0 = torch.cuda.Event(enable_timing=True)
    g1 = torch.cuda.Event(enable_timing=True)
    g0.record()
    e2e_loop(args.steps)
    g1.record()
    sync_all()
    e2e_ms = g0.elapsed_time(g1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = n / (e2e_ms / args.steps * 1e-3)
    h2d = 2 * b * d * 4 * world          # fp32 image + text embeddings of the global batch
    d2h = 4 * world

    VF.release_graphs()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return 0 if parity["ok"] else 3

    peaks = measured_peaks()
    f_alg_step = 6.0 * n * n * d                    # S, dI, dT GEMMs (recompute not counted)
    f_alg_grad = 4.0 * (n / world) * n * d          # one grad_both launch on this rank: dI and dT
    f_alg_fwd = 2.0 * (n / world) * n * d           # one forward sweep on this rank: S
    grad_tflops = f_alg_grad / (grad_ms * 1e-3) / 1e12
    fwd_tflops = f_alg_fwd / (fwd_ms * 1e-3) / 1e12
    step_tflops = f_alg_step / (ms_per_step * 1e-3) / 1e12 / world   # per GPU
    f_alg_head = f_alg_step + 6.0 * n * (F_I + F_T) * d
    head_tflops = f_alg_head / (head_ms * 1e-3) / 1e12 / world
    traffic, traffic_source = None, "not measured in this run (ncu is not part of the bench)"
    fwd_traffic, fwd_traffic_source = None, traffic_source

    def _ncu_traffic(name):
        prof = os.path.join(ROOT, "profiles", name)
        m = json.load(open(prof))["metrics"]
        return (float(m["dram__bytes_read.sum"]["value"]) + float(m["dram__bytes_write.sum"]["value"])) * 1e6

    if world == 1 and n == 32768 and d == 512:
        src = "profiles/%s: one ncu --set full capture of this kernel on this workload (committed, not this run)"
        try:
            traffic = _ncu_traffic("r02_ncu_grad_both_final.json")
            traffic_source = src % "r02_ncu_grad_both_final.json"
        except Exception:
            traffic = None
        try:
            fwd_traffic = _ncu_traffic("r02_ncu_lse_fwd_final.json")
            fwd_traffic_source = src % "r02_ncu_lse_fwd_final.json"
        except Exception:
            fwd_traffic = None
    cpu = None
    tgb = None
    cfs = None
    if world == 1 and not args.skip_cpu:
        cpu = cpu_reference_sample(n, d, args.cpu_block, 2, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        tgb = torch_gpu_baseline(torch, sets[0][0], sets[0][1], LOGIT_SCALE)
        cfs = cpu_full_step()
    n_weak = bw * world
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(n, d, world),
        "pct_of_bf16_peak": 100.0 * step_tflops / peaks["bf16_tflops"],
        "algorithmic_tflops_per_gpu": step_tflops,
        "executed_flops_per_step": 8.0 * n * n * d,
        "loss": parity["loss"],
        "parity": parity,
        "roofline": {"bound": "tensor", "kernel": "grad_both_kernel",
                     "achieved": grad_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": grad_tflops / peaks["bf16_tflops"],
                     "frac_of_sustained": grad_tflops / peaks["bf16_tflops_sustained"],
                     "frac_executed": 1.5 * grad_tflops / peaks["bf16_tflops"],
                     "frac_executed_of_sustained": 1.5 * grad_tflops / peaks["bf16_tflops_sustained"],
                     "peak_source": peaks["source"] + " (burst cuBLAS bf16)",
                     "ms_per_launch": grad_ms, "ms_per_call_with_helpers": grad_call_ms,
                     "algorithmic_flops_per_launch": f_alg_grad,
                     "executed_flops_per_launch": 1.5 * f_alg_grad,
                     "traffic": traffic, "traffic_source": traffic_source},
        "roofline_forward": {"bound": "tensor", "kernel": "lse_partial_kernel (+ 2 merge kernels)",
                             "achieved": fwd_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": fwd_tflops / peaks["bf16_tflops"],
                             "frac_of_sustained": fwd_tflops / peaks["bf16_tflops_sustained"],
                             "ms_per_call": fwd_ms, "algorithmic_flops_per_launch": f_alg_fwd,
                             "traffic": fwd_traffic, "traffic_source": fwd_traffic_source},
        "full_head": {"what": "projection + L2-normalise prologue (image 512 -> d, text 312 -> d), loss, and the "
                              "whole backward (dW, d features, d logit_scale) inside the CUDA-event bracket",
                      "value": n / (head_ms * 1e-3), "unit": UNIT, "ms_per_step": head_ms,
                      "algorithmic_flops_per_step": f_alg_head,
                      "pct_of_bf16_peak": 100.0 * head_tflops / peaks["bf16_tflops"]},
        "weak_scaling_point": {"rows_per_gpu": bw, "global_batch": n_weak, "ms_per_step": weak_ms,
                               "value": n_weak / (weak_ms * 1e-3), "unit": UNIT,
                               "algorithmic_tflops_per_gpu": 6.0 * n_weak * n_weak * d / (weak_ms * 1e-3) / 1e12 / world,
                               "note": "efficiency E(G) = this figure at G GPUs / the same figure at 1 GPU (SURVEY 8(d))"},
        "cpu_baseline": cpu,
        "torch_gpu_baseline": tgb,
        "cpu_full_step": cfs,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps,
                "note": "per step: H2D of the fp32 embeddings from pinned memory (prefetched one step ahead), "
                        "D2H of the loss; the two fp32 gradient tensors stay on the device, where their "
                        "consumers (projection / encoder backward) live"},
        "gpu_launches": int(launches),
    }
    _emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()
    if not parity["ok"]:
        sys.stderr.write("bench.py: PARITY CHECK FAILED: " + json.dumps(parity) + "\n")
        return 3
    return 0


def _all_gather_list(t, world):
    import torch
    out = [torch.empty_like(t) for _ in range(world)]
    torch.distributed.all_gather(out, t.contiguous())
    return out


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON
    line on stdout, so everything else is sent to stderr and the line goes to the saved fd."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    text = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, text)
    else:
        os.write(_REAL_STDOUT, text)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=32768, help="global batch (pairs)")
    ap.add_argument("--d", type=int, default=512, help="embedding dim")
    ap.add_argument("--cpu-block", type=int, default=1024, help="rows of the CPU reference slab")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        # (a step is a bounded slab of the workload, ~0.25 s on 16 cores: any --steps up to a few
        #  hundred ends within minutes)
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
