"""Host-side plan of the (optionally row-sharded) fused CLIP loss.

One process per GPU; rank r owns rows [r*b, (r+1)*b) of both embedding streams (b = N / world).
The plan is written against a small ``ops`` interface so that the very same collective
orchestration runs (a) in production on the CUDA kernels (``functional.CudaOps``) and (b) in the
CPU test-suite on a numpy restatement of the kernel contracts under the ``gloo`` backend
(``tests/test_distributed_cpu.py``).  Only the per-rank tile work differs; the exchange steps are:

forward   all_gather(T_loc)                       -> T_all            (b*D bf16 per rank), or
          ``ops.push_gather``: the fp16 cast of T_loc stores its rows into every peer's window
          (NVLink) and a 1-element all_reduce closes the exchange
          all_gather([col ref | col l | own diag])  -> column statistics ((2N + b) fp32 per rank)
          all_reduce([sum row_loss, sum col_loss])                     (2 fp32)
backward  dT rows stored into the owner's peer window from the kernel epilogue (NVLink, fused
          reduce-scatter; ``ops.grad_scatter`` / ``ops.scatter_finish``), or, without peer
          windows, reduce_scatter(dT_all partial) -> dT_loc          (N*D fp32 per rank)
          all_reduce(dscale)                                           (1 fp32; doubles as the
          barrier between the peer stores and the slot sum)

``ops`` contract (shapes: x [n_rows, D], y [n_cols, D]):
    lse_stats_fused(x, y, scale, diag_shift) -> (row_max, row_l, diag, col_ref, col_l)   [optional]
        one sweep: the row statistics below plus, per column j of S, an upper reference
        col_ref[j] >= max_i <x_i, y_j> and col_l[j] = sum over i != positive of
        exp(scale*<x_i,y_j> - scale*col_ref[j]).
    lse_stats(x, y, scale, diag_shift) -> (row_max, row_l, diag)
        row_max[i] = max_j <x_i, y_j> (positive pair included), row_l[i] = sum over j != positive
        of exp(scale*<x_i,y_j> - scale*row_max[i]) (up to the fp32 rounding the merge undoes),
        diag[i] = <x_i, y_{i-diag_shift}> or 0 when that column does not exist.
    merge_stats(part_max [P,n], part_l [P,n], diag [n], scale) -> (max, lg2l, q, row_loss)
    loss_sums(row_loss, col_loss) -> 2-vector
    loss_finish(sums, n_global) -> (loss, image_loss, text_loss) 3-vector                  [optional]
    grad(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale) -> (dx, ds)
    to_backward_operand(x) -> operand copy used by grad (fp16 on the GPU)
  optional, fused with their collectives over peer windows (NVLink on the GPU; emulated with gloo
  collectives by ``tests/kernel_contract_ops.WindowContractOps``):
    push_gather(i_loc, t_loc, group, world, rank) -> (i_operand, t_all_operand) | None
        gathers the text shard in the backward's operand format; ends with a collective all ranks pass
    peer_window(group, world, rank, rows, d, device) -> window | None
    grad_scatter(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale,
                 window) -> (parity, ds)      rows of dX stored into slot `rank` of their owner's window
    scatter_finish(window, parity, out_mul, out_dtype, device) -> dX rows owned by this rank
        (sum of the slots in rank order, times out_mul); call after a collective all ranks pass
"""
from __future__ import annotations

import torch


def _dist():
    import torch.distributed as dist
    return dist


def group_info(group):
    if group is None:
        return 1, 0
    dist = _dist()
    return dist.get_world_size(group), dist.get_rank(group)


def all_gather_rows(t: torch.Tensor, group, world: int) -> torch.Tensor:
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    _dist().all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def forward_plan(ops, i_loc, t_loc, scale: float, group=None, exact_columns: bool = False, ids_loc=None):
    """Returns a dict with the three losses (device scalars) and everything backward needs.

    When ``ops`` offers ``lse_stats_fused`` (and ``exact_columns`` is False) the row and the column
    statistics come out of ONE sweep over the logits; otherwise the columns get their own sweep."""
    world, rank = group_info(group)
    n_loc = i_loc.shape[0]
    n_glob = n_loc * world
    lo = rank * n_loc
    pushed = None
    if world > 1 and hasattr(ops, "push_gather"):
        # the cast to the backward's operand format stores every row into all peers' windows: the
        # all-gather rides on a kernel that has to run anyway, and one gathered copy serves both sweeps
        pushed = ops.push_gather(i_loc, t_loc, group, world, rank)
    if pushed is not None:
        i_loc, t_all = pushed
    else:
        t_all = all_gather_rows(t_loc, group, world) if world > 1 else t_loc

    fused = (not exact_columns) and hasattr(ops, "lse_stats_fused")
    ids = None
    if ids_loc is not None:
        # duplicate-caption mask: ids of the local rows and of ALL columns (one small all-gather)
        if not fused:
            raise NotImplementedError("caption ids need the fused statistics sweep")
        ids_all = all_gather_rows(ids_loc, group, world) if world > 1 else ids_loc
        ids = (ids_loc, ids_all)
    if fused and ids is not None:
        r_max_p, r_l_p, r_diag, c_max_p, c_l_p = ops.lse_stats_fused(i_loc, t_all, scale, -lo, ids[0], ids[1])
        c_diag_own = r_diag
    elif fused:
        # one sweep: rows of S owned by the local images (complete) + partial column statistics
        r_max_p, r_l_p, r_diag, c_max_p, c_l_p = ops.lse_stats_fused(i_loc, t_all, scale, -lo)
        c_diag_own = r_diag                     # S_jj seen from column j is the same logit
    else:
        # rows of S owned by the local images: complete after one sweep over T_all
        r_max_p, r_l_p, r_diag = ops.lse_stats(i_loc, t_all, scale, -lo)
        # columns (= rows of S^T owned by the texts): partial over the local images
        c_max_p, c_l_p, c_diag_full = ops.lse_stats(t_all, i_loc, scale, lo)
        c_diag_own = c_diag_full[lo:lo + n_loc]
    r_max, r_lg, r_q, r_loss = ops.merge_stats(r_max_p, r_l_p, r_diag, scale)
    if world > 1:
        # one all-gather for the three small messages: [col ref (N) | col l (N) | own diag (b)]
        n_all = c_max_p.shape[0]
        packed = torch.cat([c_max_p, c_l_p, c_diag_own.to(c_max_p.dtype)]).unsqueeze(0)
        gathered = all_gather_rows(packed, group, world)            # [world, 2N + b]
        c_max_p = gathered[:, :n_all].contiguous()
        c_l_p = gathered[:, n_all:2 * n_all].contiguous()
        c_diag = gathered[:, 2 * n_all:].reshape(-1).contiguous()
    else:
        c_diag = c_diag_own
    c_max, c_lg, c_q, c_loss = ops.merge_stats(c_max_p, c_l_p, c_diag, scale)
    sums = ops.loss_sums(r_loss, c_loss[lo:lo + n_loc])
    if world > 1:
        _dist().all_reduce(sums, group=group)
    if hasattr(ops, "loss_finish"):
        l3 = ops.loss_finish(sums, n_glob)                     # (loss, image_loss, text_loss)
    else:
        losses = sums / float(n_glob)
        l3 = torch.stack([(losses[0] + losses[1]) * 0.5, losses[0], losses[1]])   # reference :552
    return {"loss": l3[0], "image_loss": l3[1], "text_loss": l3[2], "losses": l3,
            "t_all": t_all, "bwd_operands": pushed, "r_stats": (r_max, r_lg, r_q), "c_stats": (c_max, c_lg, c_q),
            "world": world, "rank": rank, "n_loc": n_loc, "n_glob": n_glob, "ids": ids}


def backward_plan(ops, i_loc_op, t_all_op, r_stats, c_stats, scale: float, n_loc: int, n_glob: int,
                  rank: int, world: int, group=None, w_row: float = 1.0, w_col: float = 1.0,
                  need_i: bool = True, need_t: bool = True, need_scale: bool = True,
                  out_mul=None, out_dtypes=(torch.float32, torch.float32), tail_barrier=False,
                  single_sweep=True, ids=None):
    """(dI_loc, dT_loc, dscale) of the global loss; operands are the backward copies.

    ``out_mul`` (device scalar) multiplies dI and dT (not dscale); where a gradient is final on this
    rank the kernel epilogue applies it and writes ``out_dtypes`` directly."""
    lo = rank * n_loc
    d_i = d_t = ds = None
    pending = None
    window = None
    if need_t and world > 1 and hasattr(ops, "peer_window"):
        window = ops.peer_window(group, world, rank, n_loc, t_all_op.shape[1], t_all_op.device)
    one_sweep = (single_sweep or ids is not None) and need_i and need_t and hasattr(ops, "grad_both") \
        and (world == 1 or window is not None)
    if ids is not None and not one_sweep:
        raise NotImplementedError("the duplicate-caption mask is implemented in the single-recompute backward only "
                                  "(needs gradients for both embedding streams and, when sharded, peer windows)")
    if one_sweep:
        # ONE sweep over the logit tiles feeds both accumulations (8 N^2 D executed flops per step
        # instead of 10 N^2 D); with peer windows the dT rows leave through the fused reduce-scatter
        extra = {} if ids is None else {"row_ids": ids[0], "col_ids": ids[1]}
        d_i, second, ds = ops.grad_both(i_loc_op, t_all_op, r_stats, c_stats, scale, -lo, n_glob, w_row,
                                        w_col, need_scale, out_mul, out_dtypes, window, **extra)
        if window is None:
            d_t = second
        else:
            parity = second
        need_i = False        # done
        if world == 1:
            return d_i, d_t, ds
    elif window is not None:
        # fused reduce-scatter: the dT kernel stores every finished row block straight into the
        # owning rank's window over NVLink (no [N, D] partial, no NCCL kernel competing for SMs)
        parity, ds = ops.grad_scatter(t_all_op, i_loc_op, c_stats, r_stats, scale, lo, n_glob, w_col,
                                      w_row, need_scale, window)
    elif need_t and world > 1:
        # dT_all partial over the local images first, so that its fp32 reduce-scatter (the largest
        # message of the step, N*D*4 bytes per rank) runs on a side stream under the dI kernel
        d_t_all, ds = ops.grad(t_all_op, i_loc_op, c_stats, r_stats, scale, lo, n_glob, w_col, w_row,
                               need_scale, None, torch.float32)
        d_t = torch.empty((n_loc,) + tuple(d_t_all.shape[1:]), dtype=d_t_all.dtype,
                          device=d_t_all.device)
        if d_t_all.is_cuda and (need_i or need_scale):
            main = torch.cuda.current_stream()
            side = _side_stream(d_t_all.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                _dist().reduce_scatter_tensor(d_t, d_t_all, group=group)
            d_t_all.record_stream(side)
            d_t.record_stream(side)
            pending = side
        else:
            _dist().reduce_scatter_tensor(d_t, d_t_all.contiguous(), group=group)
    if (need_i or (need_scale and ds is None)) and d_i is None:
        d_i, ds_i = ops.grad(i_loc_op, t_all_op, r_stats, c_stats, scale, -lo, n_glob, w_row, w_col,
                             need_scale and ds is None, out_mul, out_dtypes[0])
        if ds is None:
            ds = ds_i
        if not need_i:
            d_i = None
    if need_t and world == 1:
        d_t, ds_t = ops.grad(t_all_op, i_loc_op, c_stats, r_stats, scale, lo, n_glob, w_col, w_row,
                             need_scale and ds is None, out_mul, out_dtypes[1])
        if ds is None:
            ds = ds_t
    if pending is not None:
        torch.cuda.current_stream().wait_stream(pending)
    if need_t and world > 1 and window is None and out_mul is not None:
        d_t = d_t * out_mul
    if need_scale and world > 1:
        _dist().all_reduce(ds, group=group)
    if (window is not None or tail_barrier) and world > 1:
        # ``tail_barrier``: the operands live in peer windows the next step's gather overwrites --
        # no rank may get there before every rank has finished reading them
        if not need_scale:      # any collective every rank passes after its kernels will do
            _dist().all_reduce(torch.zeros(1, device=t_all_op.device), group=group)
    if window is not None:
        # every peer's stores have landed: sum the slots of the local window in rank order
        d_t = ops.scatter_finish(window, parity, out_mul, out_dtypes[1], t_all_op.device)
    return d_i, d_t, ds


_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]
