"""GPU, >= 2 devices, NCCL: the row-sharded fused loss (all-gather / stat merge / reduce-scatter)
against the single-process fp64 oracle on the global batch (BASELINE config 3: 4096 x 512)."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, ls, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import vlp_b200  # noqa: F401
        from vlp_b200 import functional as VF
        from oracle import clip_oracle as O
        I, T = O.make_embeddings(n, d, rho=0.35, seed=42)      # global batch built from the seed
        b = n // world
        Il = I[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
        Tl = T[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
        lsc = torch.tensor([ls], dtype=torch.float64, device=dev, requires_grad=True)
        # 4 steps: the first runs eagerly, the second captures fwd+bwd (kernels + NCCL collectives)
        # into a CUDA graph, the rest replay it -- results must not change
        first = None
        for _ in range(4):
            Il.grad = Tl.grad = lsc.grad = None
            loss, il, tl = VF.fused_clip_loss_from_embeddings(Il, Tl, lsc, group=dist.group.WORLD)
            loss.backward()
            torch.cuda.synchronize()
            cur = (loss.item(), Il.grad.clone(), Tl.grad.clone(), lsc.grad.item())
            if first is None:
                first = cur
            else:
                assert cur[0] == first[0] and cur[3] == first[3]
                assert torch.equal(cur[1], first[1]) and torch.equal(cur[2], first[2])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss.item(), image_loss=il.item(),
                 text_loss=tl.item(), dI=Il.grad.cpu().numpy(), dT=Tl.grad.cpu().numpy(),
                 dl=lsc.grad.item(), used_peer_windows=int(any(w.ok for w in VF._PEER_WINDOWS.values())))
        # gradient averaging (DDP): grad_scale = world multiplies the row-sharded gradients only;
        # d logit_scale is already the all-reduced global total on every rank and must stay as it is
        Il.grad = Tl.grad = lsc.grad = None
        loss_s, _, _ = VF.fused_clip_loss_from_embeddings(Il, Tl, lsc, group=dist.group.WORLD,
                                                          grad_scale=float(world))
        loss_s.backward()
        torch.cuda.synchronize()
        assert abs(lsc.grad.item() - first[3]) <= 1e-6 * abs(first[3])
        assert (Il.grad - world * first[1]).norm() <= 1e-6 * (world * first[1]).norm()
        assert (Tl.grad - world * first[2]).norm() <= 1e-6 * (world * first[2]).norm()
        if torch.cuda.device_count() > 1 and rank == 0:
            # device guard: operands on this rank's GPU while another device is current
            other = (rank + 1) % torch.cuda.device_count()
            with torch.cuda.device(other):
                a = Il.detach().clone().requires_grad_(True)
                l_other, _, _ = VF.fused_clip_loss_from_embeddings(a, Tl.detach(), lsc.detach())
                l_other.backward()
            torch.cuda.synchronize(dev)
            assert torch.isfinite(l_other) and torch.isfinite(a.grad).all()
        # the same step with the two-pass backward (dT kernel with the fused reduce-scatter, then the
        # dI kernel) instead of the single-recompute kernel: identical up to fp32 summation order
        VF.SINGLE_SWEEP, VF._GRAPH_MODE = False, "0"
        Il.grad = Tl.grad = lsc.grad = None
        loss1, _, _ = VF.fused_clip_loss_from_embeddings(Il, Tl, lsc, group=dist.group.WORLD)
        loss1.backward()
        torch.cuda.synchronize()
        # (two different kernels: their fp16 G entries round differently here and there -- each flip is
        #  one fp16 ulp, 4.9e-4 of that entry -- so the bound is a cross-implementation sanity bound,
        #  well inside the 1e-3 the oracle comparison below allows; small shards see the largest values)
        XTOL = 3e-4
        e_t = ((Tl.grad - first[2]).norm() / first[2].norm()).item()
        e_i = ((Il.grad - first[1]).norm() / first[1].norm()).item()
        assert abs(loss1.item() - first[0]) <= 1e-6 * abs(first[0])
        assert e_t <= XTOL and e_i <= XTOL, f"two-pass vs single-recompute: dT {e_t:.2e} dI {e_i:.2e}"
        VF.SINGLE_SWEEP = True
        # the same step with NCCL reduce-scatter instead of the fused NVLink stores: identical up
        # to the fp32 summation order of the partials
        VF.PEER_RS_MODE, VF._GRAPH_MODE = "0", "0"
        Il.grad = Tl.grad = lsc.grad = None
        loss2, _, _ = VF.fused_clip_loss_from_embeddings(Il, Tl, lsc, group=dist.group.WORLD)
        loss2.backward()
        torch.cuda.synchronize()
        e_t = ((Tl.grad - first[2]).norm() / first[2].norm()).item()
        e_i = ((Il.grad - first[1]).norm() / first[1].norm()).item()
        assert abs(loss2.item() - first[0]) <= 1e-6 * abs(first[0])
        assert e_t <= XTOL and e_i <= XTOL, f"NCCL reduce-scatter vs fused: dT {e_t:.2e} dI {e_i:.2e}"
    finally:
        try:
            VF.release_graphs()
        except Exception:
            pass
        dist.destroy_process_group()


# BASELINE config 3 (4096 x 512 over 2 / 4 / 8 GPUs) plus ragged shards (rows per rank not a multiple
# of the 128-row tile, K not a multiple of 64)
@pytest.mark.parametrize("world,n,d,ls", [(2, 4096, 512, math.log(1 / 0.07)), (4, 4096, 512, math.log(1 / 0.07)),
                                          (8, 4096, 512, math.log(1 / 0.07)), (2, 600, 72, 3.0),
                                          (4, 1000, 72, 3.0), (8, 2408, 136, math.log(50.0))])
def test_sharded_global_batch_matches_oracle(tmp_path, world, n, d, ls):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from oracle import clip_oracle as O
    mp.spawn(_worker, args=(world, _free_port(), n, d, ls, str(tmp_path)), nprocs=world, join=True)
    I, T = O.make_embeddings(n, d, rho=0.35, seed=42)
    ref = O.closed_form(I.numpy(), T.numpy(), ls)
    b = n // world
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        if all(torch.cuda.can_device_access_peer(a, b_) for a in range(world) for b_ in range(world)
               if a != b_):
            assert int(got["used_peer_windows"]) == 1, "fused reduce-scatter fell back to NCCL"
        assert abs(float(got["loss"]) - ref["loss"]) < 1e-4 * ref["loss"]
        assert abs(float(got["image_loss"]) - ref["image_loss"]) < 1e-4 * ref["image_loss"]
        assert O.rel_err(got["dI"], ref["dI"][r * b:(r + 1) * b]) < 1e-3
        assert O.rel_err(got["dT"], ref["dT"][r * b:(r + 1) * b]) < 1e-3
        assert abs(float(got["dl"]) - ref["dlogit_scale"]) < 1e-3 * abs(ref["dlogit_scale"])
