"""Whole pre-training step (ResNet34 + TinyBERT + CLIP head) -- BASELINE.json configs[0] and configs[3].

The encoders are NOT part of the fused path (they stay stock PyTorch); this tool only answers
"what share of a real step is the head, and what does the fused head buy end to end".

    python tools/full_step.py cpu [--batch 32] [--dim 128] [--steps 3]
        configs[0]: the reference's step on the host cores with stock torch ops everywhere
        (synthetic 224x224 grayscale radiographs replicated to 3 channels as PretrainDataModule.py:168
        does, 32-token captions, AdamW 5e-5, seed 42).  A reported baseline, nothing of ours runs.

    python tools/full_step.py gpu [--batch-per-gpu 1024] [--dim 512] [--steps 10] [--head fused|torch]
        configs[3] (one rank per GPU under torchrun, or a single GPU): bf16-autocast encoders,
        replica gradients averaged with one flat all-reduce (DDP semantics), the fused head on the global batch
        (--head fused, the drop-in VisionLanguageModule's kernels) or the stock torch loss on all-gathered
        embeddings (--head torch, the control).  Prints pairs/s and the head's share of the step.

One JSON line per run on stdout.
"""
from __future__ import annotations

import argparse
import functools
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("VLP_B200_RANDOM_INIT", "1")   # no network: TinyBERT from its config


def synthetic_batch(bsz, device, seq=32, res=224, seed=42):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(bsz, 1, res, res, generator=g).repeat(1, 3, 1, 1)
    ids = torch.randint(0, 30522, (bsz, seq), generator=g)
    return {"x-ray": x.to(device),
            "caption_tokenized": {"input_ids": ids.to(device),
                                  "token_type_ids": torch.zeros(bsz, seq, dtype=torch.long, device=device),
                                  "attention_mask": torch.ones(bsz, seq, dtype=torch.long, device=device)},
            "label": torch.zeros(bsz, dtype=torch.long, device=device), "caption": ["c"] * bsz}


def torch_head(image_features, text_features, w_img, w_txt, logit_scale):
    """The reference's head, op for op (VisionLanguageModule.py:448-459, 533-552)."""
    import torch
    import torch.nn.functional as F
    i = F.normalize(image_features @ w_img)
    t = F.normalize(text_features @ w_txt)
    s = torch.clamp(logit_scale.exp(), max=100)
    logits = (i @ t.T) * s
    labels = torch.arange(len(logits), device=logits.device)
    return (F.cross_entropy(logits, labels) + F.cross_entropy(logits.T, labels)) / 2


def build_encoders(dim):
    import torch
    from vlp_b200.module import ImageEncoder, TextEncoder
    torch.manual_seed(42)
    img, txt = ImageEncoder("resnet34"), TextEncoder("tinybert")
    w_img = torch.nn.Parameter(torch.randn(512, dim) * 512 ** -0.5)
    w_txt = torch.nn.Parameter(torch.randn(312, dim) * 312 ** -0.5)
    ls = torch.nn.Parameter(torch.tensor([math.log(1 / 0.07)]))   # fp64 like the reference (:111)
    return img, txt, w_img, w_txt, ls


def run_cpu(args):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device("cpu")
    img, txt, w_img, w_txt, ls = build_encoders(args.dim)
    params = list(img.parameters()) + list(txt.parameters()) + [w_img, w_txt, ls]
    opt = torch.optim.AdamW(params, lr=5e-5)
    batch = synthetic_batch(args.batch, dev)
    t_step, t_head = [], []
    for step in range(args.steps + 1):
        t0 = time.perf_counter()
        f_i = img(batch["x-ray"])
        f_t = txt(**batch["caption_tokenized"])
        t1 = time.perf_counter()
        loss = torch_head(f_i, f_t, w_img, w_txt, ls)
        t2 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        t3 = time.perf_counter()
        if step > 0:   # first step = warm-up
            t_step.append(t3 - t0)
            t_head.append(t2 - t1)
    s = sorted(t_step)[len(t_step) // 2]
    print(json.dumps({"config": "ResNet34+TinyBERT CLIP step on CPU (stock torch, reference ops)",
                      "batch": args.batch, "dim": args.dim, "cores": cores, "s_per_step": s,
                      "pairs_per_s": args.batch / s,
                      "head_forward_share": sorted(t_head)[len(t_head) // 2] / s,
                      "loss": float(loss.detach()), "steps": args.steps, "data": "synthetic"}))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import vlp_b200  # noqa: F401
    from vlp_b200 import functional as VF
    from vlp_b200.module import VisionLanguageModule
    torch.manual_seed(42)
    m = VisionLanguageModule(image_model="resnet34", text_encoder_model="tinybert",
                             optimizer=functools.partial(torch.optim.AdamW, lr=5e-5), deduplicate=False,
                             masked_loss=False, image_embedding_dim=512, text_embedding_dim=312,
                             embedding_dim=args.dim).to(dev)
    m = m.to(memory_format=torch.channels_last)
    opt = m.configure_optimizers()["optimizer"]
    bsz = args.batch_per_gpu
    batch = synthetic_batch(bsz, dev, seed=42 + rank)
    batch["x-ray"] = batch["x-ray"].contiguous(memory_format=torch.channels_last)

    def gathered(t):
        if world == 1:
            return t
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.detach())
        parts[rank] = t          # keep the local shard in the graph (DDP averages the rest)
        return torch.cat(parts)

    def sync_grads():
        """Average the replicas' gradients (what DDP does), as one flat all-reduce."""
        if world == 1:
            return
        ps = [p for p in m.parameters() if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1).float() for p in ps])
        dist.all_reduce(flat)
        flat /= world
        o = 0
        for p in ps:
            n = p.grad.numel()
            p.grad.copy_(flat[o:o + n].view_as(p.grad))
            o += n

    def step(timers=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timers is not None else None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f_i = m.image_encoder(batch["x-ray"])
            f_t = m.text_encoder(**batch["caption_tokenized"])
        f_i, f_t = f_i.float(), f_t.float()
        if ev:
            ev[0].record()
        if args.head == "fused":
            # the drop-in module's own head: fused projection + normalise, fused loss on the global batch
            i_emb, i_bf16, i_f16 = VF.project_normalize(f_i, m.image_projection)
            t_emb, t_bf16, t_f16 = VF.project_normalize(f_t, m.text_projection)
            group = dist.group.WORLD if world > 1 else None
            loss, _, _ = VF.fused_clip_loss_from_embeddings(i_emb, t_emb, m.logit_scale, group=group,
                                                            grad_scale=float(world),
                                                            _operands=(i_bf16, t_bf16, i_f16, t_f16))
        else:
            import torch.nn.functional as F
            i = gathered(F.normalize(f_i @ m.image_projection))
            t = gathered(F.normalize(f_t @ m.text_projection))
            s = torch.clamp(m.logit_scale.exp(), max=100)
            logits = (i @ t.T) * s
            labels = torch.arange(len(logits), device=dev)
            loss = (F.cross_entropy(logits, labels) + F.cross_entropy(logits.T, labels)) / 2
            loss = loss * world   # each rank back-propagates its shard's share; sync_grads averages
        if ev:
            ev[1].record()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        sync_grads()
        opt.step()
        if ev:
            ev[2].record()
            timers.append(ev)
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    timers = []
    if world > 1:
        dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step(timers)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    head_ms = sum(e[0].elapsed_time(e[1]) for e in timers) / len(timers)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "full VLP step ResNet34+TinyBERT bf16", "head": args.head,
                          "n_gpus": world, "global_batch": bsz * world, "dim": args.dim,
                          "ms_per_step": float(t.item()), "pairs_per_s": bsz * world / (float(t.item()) * 1e-3),
                          "head_forward_ms": head_ms, "loss": float(loss.detach()),
                          "steps": args.steps, "data": "synthetic"}))
    if world > 1:
        VF.release_graphs()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["cpu", "gpu"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--batch-per-gpu", type=int, default=1024)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--head", choices=["fused", "torch"], default="fused")
    args = ap.parse_args()
    if args.dim is None:
        args.dim = 128 if args.mode == "cpu" else 512
    if args.steps is None:
        args.steps = 3 if args.mode == "cpu" else 10
    (run_cpu if args.mode == "cpu" else run_gpu)(args)


if __name__ == "__main__":
    main()
