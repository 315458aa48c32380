"""Functional core of the fused CLIP head (host side, thin Python over the C ABI).

Mirrors the arithmetic of the reference's ``VisionLanguageModule.forward`` (lines 448-459) and
``_compute_loss`` (lines 533-552) but never builds the N x N logit matrix:

    fused_clip_loss_from_embeddings(I, T, logit_scale)        embedding-level entry (parity surface)
    fused_clip_loss(f_img, f_txt, W_img, W_txt, logit_scale)  projection + normalise + loss
    project_normalize(features, W)                            prologue only (returns embeddings)

All compute runs in ``csrc/libvlpclip.so`` (hand-written sm_100a kernels).  Tensors must live on a
B200; anything else raises -- there is no CPU / eager fallback.  PyTorch only provides memory,
streams, autograd bookkeeping and (for the sharded variant) ``torch.distributed`` collectives.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch

from . import _lib, sharded

LOGIT_SCALE_MAX = 100.0  # reference: torch.clamp(logit_scale.exp(), max=100)  (:457)
# VLP_B200_SINGLE_SWEEP=0: two grad_pair_kernel passes (dI, dT; S recomputed per direction) instead
# of the single-recompute vlpclip_grad_both kernel
SINGLE_SWEEP = os.environ.get("VLP_B200_SINGLE_SWEEP", "1") != "0"
# VLP_B200_EXACT_COLUMNS=1: give the column statistics their own sweep (exact column maxima)
# instead of fusing them into the row sweep (see include/vlpclip.h: vlpclip_lse_fwd_fused)
EXACT_COLUMNS = os.environ.get("VLP_B200_EXACT_COLUMNS", "0") == "1"


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
_SM_RESERVED = False


def _reserve_sms_for_collectives(n_free: int = int(os.environ.get("VLP_B200_NCCL_SMS", "4"))):
    """Sharded mode: keep a few SMs out of the persistent kernels so NCCL can overlap with them."""
    global _SM_RESERVED
    if not _SM_RESERVED:
        lib = _lib.load()
        lib.vlpclip_set_sm_limit(max(2, lib.vlpclip_sm_count() - n_free))
        _SM_RESERVED = True


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _device_guard(*tensors):
    """Context manager making the tensors' device current: the C side launches on the current
    device / stream (cudaGetDevice), so operands on cuda:1 while cuda:0 is current would launch on
    the wrong GPU.  All tensor operands must share one device."""
    dev = None
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise ValueError(f"operands live on different devices ({dev} and {t.device})")
    if dev is None:
        import contextlib
        return contextlib.nullcontext()
    return torch.cuda.device(dev)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} lives on {t.device}: the fused CLIP head only runs on a CUDA "
                           "sm_100 device (no CPU fallback)")


def _as_2d_contig(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D [batch, dim], got shape {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def as_scale_tensor(scale, device) -> torch.Tensor:
    """fp32 device scalar [1] holding s; accepts a python number or a tensor (no host sync for the
    latter)."""
    if isinstance(scale, torch.Tensor):
        return scale.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()
    return torch.full((1,), float(scale), dtype=torch.float32, device=device)


def scale_from_logit_scale(logit_scale: torch.Tensor):
    """(s, ds/dl) = (min(e^l, 100), e^l or 0) as fp32 device scalars, one launch, no host sync
    (reference :456-457; evaluated in double like the reference's float64 parameter)."""
    lib = _lib.load()
    ls = logit_scale.detach().reshape(1)
    if ls.dtype not in (torch.float32, torch.float64):
        ls = ls.double()
    ls = ls.contiguous()
    out = torch.empty(2, dtype=torch.float32, device=ls.device)
    rc = lib.vlpclip_scale_prep(ls.data_ptr(), 1 if ls.dtype == torch.float64 else 0,
                                out.data_ptr(), out.data_ptr() + 4, _stream())
    _lib.check(rc, "scale_prep")
    return out[0:1], out[1:2]


def cast_bf16_to_f16(src: torch.Tensor) -> torch.Tensor:
    """fp16 copy of bf16 embeddings (operands of the backward GEMMs)."""
    _require_cuda(src, "src")
    assert src.dtype == torch.bfloat16 and src.is_contiguous()
    dst = torch.empty_like(src, dtype=torch.float16)
    lib = _lib.load()
    _lib.check(lib.vlpclip_cast_bf16_to_f16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()),
               "cast_bf16_to_f16")
    return dst


def cast_f32_operands(src: torch.Tensor, want_f16: bool = True):
    """(bf16, fp16 | None) operand copies of fp32 embeddings in one pass (the fp16 copy is the image
    of the bf16-rounded values, so forward and backward see identical operands)."""
    _require_cuda(src, "src")
    assert src.dtype == torch.float32 and src.is_contiguous() and src.numel() % 4 == 0
    b = torch.empty_like(src, dtype=torch.bfloat16)
    h = torch.empty_like(src, dtype=torch.float16) if want_f16 else None
    lib = _lib.load()
    _lib.check(lib.vlpclip_cast_f32_operands(src.data_ptr(), b.data_ptr(), h.data_ptr() if want_f16 else None,
                                             src.numel(), _stream()), "cast_f32_operands")
    return b, h


def lse_stats(x_bf16: torch.Tensor, y_bf16: torch.Tensor, scale: float, diag_shift: int = 0):
    """Row statistics of ``scale * x @ y.T``: (row_max [n], row_l [n], diag [n]).

    ``row_l`` leaves out the positive pair (column ``i - diag_shift``); ``diag`` holds its raw
    cosine (0 for rows without a partner column)."""
    _require_cuda(x_bf16, "x")
    _require_cuda(y_bf16, "y")
    if x_bf16.dtype != y_bf16.dtype or x_bf16.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("lse_stats expects two bf16 (or two fp16) operands")
    x = _as_2d_contig(x_bf16, "x")
    y = _as_2d_contig(y_bf16, "y")
    if x.shape[1] != y.shape[1]:
        raise ValueError(f"embedding dims differ: {x.shape[1]} vs {y.shape[1]}")
    n_rows, d = x.shape
    n_cols = y.shape[0]
    lib = _lib.load()
    dev = x.device
    sc = as_scale_tensor(scale, dev)
    row_max = torch.empty(n_rows, dtype=torch.float32, device=dev)
    row_l = torch.empty(n_rows, dtype=torch.float32, device=dev)
    diag = torch.zeros(n_rows, dtype=torch.float32, device=dev)
    nbytes = lib.vlpclip_lse_workspace_bytes(n_rows, n_cols, d)
    ws = _ws(nbytes, dev)
    fn = lib.vlpclip_lse_fwd_f16 if x.dtype == torch.float16 else lib.vlpclip_lse_fwd
    rc = fn(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), n_rows, n_cols, d,
            sc.data_ptr(), int(diag_shift), row_max.data_ptr(), row_l.data_ptr(),
            diag.data_ptr(), ws.data_ptr(), nbytes, _stream())
    _lib.check(rc, "lse_fwd")
    return row_max, row_l, diag


def _as_ids(ids: torch.Tensor, n: int, device, name: str) -> torch.Tensor:
    if ids.numel() != n:
        raise ValueError(f"{name}: expected {n} caption ids, got {ids.numel()}")
    return ids.detach().to(device=device, dtype=torch.int32).reshape(n).contiguous()


def lse_stats_fused(x_bf16: torch.Tensor, y_bf16: torch.Tensor, scale: float, diag_shift: int = 0,
                    row_ids: Optional[torch.Tensor] = None, col_ids: Optional[torch.Tensor] = None):
    """One sweep over ``scale * x @ y.T``: row statistics (as ``lse_stats``) plus the statistics of
    every column over the given rows: (row_max, row_l, diag, col_ref, col_l).  ``row_ids`` /
    ``col_ids`` (caption ids >= 0): logits whose row and column ids agree are excluded from both
    directions unless they are the positive pair (duplicate-caption mask)."""
    _require_cuda(x_bf16, "x")
    _require_cuda(y_bf16, "y")
    if x_bf16.dtype != y_bf16.dtype or x_bf16.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("lse_stats_fused expects two bf16 (or two fp16) operands")
    x = _as_2d_contig(x_bf16, "x")
    y = _as_2d_contig(y_bf16, "y")
    if x.shape[1] != y.shape[1]:
        raise ValueError(f"embedding dims differ: {x.shape[1]} vs {y.shape[1]}")
    n_rows, d = x.shape
    n_cols = y.shape[0]
    lib = _lib.load()
    dev = x.device
    sc = as_scale_tensor(scale, dev)
    row_max = torch.empty(n_rows, dtype=torch.float32, device=dev)
    row_l = torch.empty(n_rows, dtype=torch.float32, device=dev)
    diag = torch.zeros(n_rows, dtype=torch.float32, device=dev)
    col_max = torch.empty(n_cols, dtype=torch.float32, device=dev)
    col_l = torch.empty(n_cols, dtype=torch.float32, device=dev)
    nbytes = lib.vlpclip_lse_fused_workspace_bytes(n_rows, n_cols, d)
    ws = _ws(nbytes, dev)
    if (row_ids is None) != (col_ids is None):
        raise ValueError("row_ids and col_ids must be given together")
    if row_ids is not None:
        rid, cid = _as_ids(row_ids, n_rows, dev, "row_ids"), _as_ids(col_ids, n_cols, dev, "col_ids")
        rc = lib.vlpclip_lse_fwd_fused_masked(
            x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), n_rows, n_cols, d,
            1 if x.dtype == torch.float16 else 0, sc.data_ptr(), int(diag_shift), rid.data_ptr(),
            cid.data_ptr(), row_max.data_ptr(), row_l.data_ptr(), diag.data_ptr(), col_max.data_ptr(),
            col_l.data_ptr(), ws.data_ptr(), nbytes, _stream())
        _lib.check(rc, "lse_fwd_fused_masked")
        return row_max, row_l, diag, col_max, col_l
    fn = lib.vlpclip_lse_fwd_fused_f16 if x.dtype == torch.float16 else lib.vlpclip_lse_fwd_fused
    rc = fn(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), n_rows, n_cols, d, sc.data_ptr(),
            int(diag_shift), row_max.data_ptr(), row_l.data_ptr(), diag.data_ptr(),
            col_max.data_ptr(), col_l.data_ptr(), ws.data_ptr(), nbytes, _stream())
    _lib.check(rc, "lse_fwd_fused")
    return row_max, row_l, diag, col_max, col_l


def merge_stats(part_max: torch.Tensor, part_l: torch.Tensor, diag: Optional[torch.Tensor],
                scale: float, want_lse: bool = False):
    """Merge [P, n] partial statistics and fold the positive-pair logits back in.

    Returns (max [n], lg2l [n], q [n], row_loss [n] | None, lse [n] | None)."""
    if part_max.dim() == 1:
        part_max = part_max.unsqueeze(0)
        part_l = part_l.unsqueeze(0)
    part_max = part_max.contiguous()
    part_l = part_l.contiguous()
    nparts, n = part_max.shape
    dev = part_max.device
    lib = _lib.load()
    sc = as_scale_tensor(scale, dev)
    out_max = torch.empty(n, dtype=torch.float32, device=dev)
    out_lg = torch.empty(n, dtype=torch.float32, device=dev)
    out_q = torch.empty(n, dtype=torch.float32, device=dev)
    out_loss = torch.empty(n, dtype=torch.float32, device=dev) if diag is not None else None
    lse = torch.empty(n, dtype=torch.float32, device=dev) if want_lse else None
    rc = lib.vlpclip_lse_merge(part_max.data_ptr(), part_l.data_ptr(),
                               diag.data_ptr() if diag is not None else None, nparts, n,
                               sc.data_ptr(), lse.data_ptr() if want_lse else None,
                               out_max.data_ptr(), None, out_lg.data_ptr(), out_q.data_ptr(),
                               out_loss.data_ptr() if out_loss is not None else None, _stream())
    _lib.check(rc, "lse_merge")
    return out_max, out_lg, out_q, out_loss, lse


def _loss_sums(row_loss: torch.Tensor, col_loss: torch.Tensor) -> torch.Tensor:
    """[sum_i row_loss, sum_i col_loss] as a 2-vector (fp32, device; fixed summation order)."""
    lib = _lib.load()
    out2 = torch.empty(2, dtype=torch.float32, device=row_loss.device)
    rc = lib.vlpclip_loss_reduce(row_loss.data_ptr(), col_loss.data_ptr(), row_loss.numel(),
                                 out2.data_ptr(), _stream())
    _lib.check(rc, "loss_reduce")
    return out2


def _loss_finish(sums: torch.Tensor, n_global: int) -> torch.Tensor:
    lib = _lib.load()
    out3 = torch.empty(3, dtype=torch.float32, device=sums.device)
    rc = lib.vlpclip_loss_finish(sums.data_ptr(), int(n_global), out3.data_ptr(), _stream())
    _lib.check(rc, "loss_finish")
    return out3


def _grad(x_f16, y_f16, x_stats, y_stats, scale: float, diag_shift: int, n_global: int,
          w_row: float = 1.0, w_col: float = 1.0, want_dscale: bool = False,
          out_mul: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float32):
    """x_stats / y_stats = (max, lg2l, q) of the rows of S owned by X / Y.

    ``out_mul`` (fp32 device scalar) is folded into dX by the kernel epilogue; ``out_dtype`` may be
    float32 or bfloat16."""
    lib = _lib.load()
    n_rows, d = x_f16.shape
    n_cols = y_f16.shape[0]
    dev = x_f16.device
    sc = as_scale_tensor(scale, dev)
    assert out_dtype in (torch.float32, torch.bfloat16)
    dx = torch.empty(n_rows, d, dtype=out_dtype, device=dev)
    ds = torch.zeros(1, dtype=torch.float32, device=dev) if want_dscale else None
    nbytes = lib.vlpclip_grad_workspace_bytes(n_rows, n_cols, d)
    ws = _ws(nbytes, dev)
    rc = lib.vlpclip_grad(x_f16.data_ptr(), x_f16.stride(0), y_f16.data_ptr(), y_f16.stride(0),
                          x_stats[0].data_ptr(), x_stats[1].data_ptr(), x_stats[2].data_ptr(),
                          y_stats[0].data_ptr(), y_stats[1].data_ptr(), y_stats[2].data_ptr(),
                          n_rows, n_cols, d, sc.data_ptr(), int(diag_shift), int(n_global),
                          float(w_row), float(w_col),
                          out_mul.data_ptr() if out_mul is not None else None,
                          1 if out_dtype == torch.bfloat16 else 0, dx.data_ptr(),
                          ds.data_ptr() if want_dscale else None, ws.data_ptr(), nbytes, _stream())
    _lib.check(rc, "grad")
    return dx, ds


def _grad_both(x_f16, y_f16, x_stats, y_stats, scale, diag_shift: int, n_global: int,
               w_row: float = 1.0, w_col: float = 1.0, want_dscale: bool = False,
               out_mul: Optional[torch.Tensor] = None, out_dtypes=(torch.float32, torch.float32),
               window: "Optional[PeerWindow]" = None, row_ids: Optional[torch.Tensor] = None,
               col_ids: Optional[torch.Tensor] = None):
    """Single-recompute backward: (dX, dY | parity, dscale) from ONE sweep over the logit tiles.
    ``row_ids`` / ``col_ids``: duplicate-caption mask (the statistics must come from the masked forward).

    Without ``window`` dY is returned as a tensor; with a peer window the final dY rows are stored
    into the owning ranks' windows (fused reduce-scatter) and the window parity is returned instead
    (finish with ``_scatter_finish`` after a collective)."""
    lib = _lib.load()
    n_rows, d = x_f16.shape
    n_cols = y_f16.shape[0]
    dev = x_f16.device
    sc = as_scale_tensor(scale, dev)
    assert out_dtypes[0] in (torch.float32, torch.bfloat16) and out_dtypes[1] in (torch.float32, torch.bfloat16)
    dx = torch.empty(n_rows, d, dtype=out_dtypes[0], device=dev)
    ds = torch.zeros(1, dtype=torch.float32, device=dev) if want_dscale else None
    nbytes = lib.vlpclip_grad_both_workspace_bytes(n_rows, n_cols, d)
    ws = _ws(nbytes, dev)
    if window is None:
        dy = torch.empty(n_cols, d, dtype=out_dtypes[1], device=dev)
        owners, n_owners, rows_per_owner, second = None, 0, 0, dy
    else:
        assert n_cols == window.rows * window.world and d == window.d
        dy = None
        parity = window.next_parity()
        owners, n_owners, rows_per_owner, second = window.owner_rows(parity), window.world, window.rows, parity
    common = (x_f16.data_ptr(), x_f16.stride(0), y_f16.data_ptr(), y_f16.stride(0),
              x_stats[0].data_ptr(), x_stats[1].data_ptr(), x_stats[2].data_ptr(),
              y_stats[0].data_ptr(), y_stats[1].data_ptr(), y_stats[2].data_ptr(),
              n_rows, n_cols, d, sc.data_ptr(), int(diag_shift), int(n_global), float(w_row), float(w_col),
              out_mul.data_ptr() if out_mul is not None else None,
              1 if out_dtypes[0] == torch.bfloat16 else 0, dx.data_ptr(),
              1 if out_dtypes[1] == torch.bfloat16 else 0, dy.data_ptr() if dy is not None else None,
              owners, n_owners, rows_per_owner, ds.data_ptr() if want_dscale else None)
    if row_ids is not None:
        rid, cid = _as_ids(row_ids, n_rows, dev, "row_ids"), _as_ids(col_ids, n_cols, dev, "col_ids")
        rc = lib.vlpclip_grad_both_masked(*common, rid.data_ptr(), cid.data_ptr(), ws.data_ptr(), nbytes, _stream())
    else:
        rc = lib.vlpclip_grad_both(*common, ws.data_ptr(), nbytes, _stream())
    _lib.check(rc, "grad_both")
    return dx, second, ds


# ----------------------------------------------------------------------------------------------
# fused reduce-scatter over NVLink peer memory
# ----------------------------------------------------------------------------------------------
# VLP_B200_PEER_RS=auto|1|0: the sharded backward stores dT rows from the accumulator epilogue
# straight into the owning rank's window (CUDA IPC peer memory) instead of calling NCCL
# reduce-scatter on a [N, D] fp32 partial.  auto = use it when every rank could map every window.
PEER_RS_MODE = os.environ.get("VLP_B200_PEER_RS", "auto").lower()
# VLP_B200_PUSH_GATHER=auto|1|0: all-gather the text shard by storing its fp16 operand copy into every
# peer's window from the cast kernel (the forward then sweeps the fp16 copies); needs the windows.
PUSH_GATHER_MODE = os.environ.get("VLP_B200_PUSH_GATHER", "auto").lower()
MAX_PEER_RANKS = 8
_PEER_WINDOWS = {}


class _DeviceMemory:
    """Zero-copy torch view of library-owned device memory (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2, "strides": None}


class PeerWindow:
    """Exchange window on every rank, mapped into every peer:
    fp32 [2 parities][world slots][rows, d] (reduce-scatter slots) followed by
    fp16 [2 parities][world * rows, d] (gathered text operand).

    Rank s writes its partial of the rows owned by rank o into slot s of o's window; after a
    collective that all ranks pass, o sums its slots in slot order (``vlpclip_slot_sum``).  Two
    parities alternate between calls so that a rank still summing call k is never overwritten
    by a peer already storing call k+1."""

    def __init__(self, group, world: int, rank: int, rows: int, d: int, device):
        import ctypes
        dist = sharded._dist()
        lib = _lib.load()
        self.group, self.world, self.rank = group, world, rank
        self.rows, self.d = rows, d
        self.slot_bytes = rows * d * 4
        self.rs_bytes = 2 * world * self.slot_bytes
        self.gather_bytes = world * rows * d * 2          # one parity of the gathered fp16 operand
        self.local = None
        self.peers = [None] * world
        self.parity = 0
        self.gather_parity = 0
        self.device = device
        ok = 1
        handle = (ctypes.c_ubyte * 64)()
        ptr = ctypes.c_void_p()
        if lib.vlpclip_peer_alloc(self.rs_bytes + 2 * self.gather_bytes, ctypes.byref(ptr),
                                  handle) == 0:
            self.local = ptr.value
        else:
            ok = 0
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle) if ok else None, group=group)
        if ok and all(h is not None for h in handles):
            for r, h in enumerate(handles):
                if r == rank:
                    self.peers[r] = self.local
                    continue
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                q = ctypes.c_void_p()
                if lib.vlpclip_peer_open(buf, ctypes.byref(q)) != 0:
                    ok = 0
                    break
                self.peers[r] = q.value
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)     # also a barrier
        self.ok = bool(flag.item())
        self._gathered = None
        if self.ok:      # views are created here, never inside a CUDA-graph capture
            self._gathered = [self._view(0), self._view(1)]
        if not self.ok:
            self.error = lib.vlpclip_last_error().decode(errors="replace")
            self.close(barrier=False)

    def next_parity(self) -> int:
        self.parity ^= 1
        return self.parity

    def owner_rows(self, parity: int):
        """Host array: for every owner o, the address of THIS rank's slot inside o's window."""
        import ctypes
        off = (parity * self.world + self.rank) * self.slot_bytes
        return (ctypes.c_void_p * self.world)(*[p + off for p in self.peers])

    def slots(self, parity: int) -> int:
        return self.local + parity * self.world * self.slot_bytes

    def next_gather_parity(self) -> int:
        self.gather_parity ^= 1
        return self.gather_parity

    def gather_dsts(self, parity: int):
        """Host array: for every rank r, where THIS rank's rows live inside r's gathered operand."""
        import ctypes
        off = self.rs_bytes + parity * self.gather_bytes + self.rank * self.rows * self.d * 2
        return (ctypes.c_void_p * self.world)(*[p + off for p in self.peers])

    def _view(self, parity: int) -> torch.Tensor:
        mem = _DeviceMemory(self.local + self.rs_bytes + parity * self.gather_bytes,
                            (self.world * self.rows, self.d), "<f2")
        t = torch.as_tensor(mem, device=self.device)
        t._vlp_owner = mem
        return t

    def gathered(self, parity: int) -> torch.Tensor:
        """[world * rows, d] fp16 view of the local gathered operand (no copy)."""
        return self._gathered[parity]

    def close(self, barrier: bool = True):
        lib = _lib.load()
        for r, q in enumerate(self.peers):
            if q is not None and r != self.rank:
                lib.vlpclip_peer_close(q)
        self.peers = [None] * self.world
        if barrier:      # nobody may free a window a peer still has mapped
            try:
                sharded._dist().barrier(group=self.group)
            except Exception:
                barrier = False
        if self.local is not None and barrier:
            lib.vlpclip_peer_free(self.local)
        self.local = None      # (without a barrier the memory is left to process exit)


def peer_window(group, world: int, rank: int, rows: int, d: int, device) -> Optional[PeerWindow]:
    """The window for this (group, shard shape), or None when peer mode is off / unavailable.
    Collective: every rank of ``group`` must call it at the same point (first use of a shape)."""
    if PEER_RS_MODE == "0" or world < 2 or world > MAX_PEER_RANKS or device.type != "cuda":
        return None
    if torch.cuda.is_current_stream_capturing():
        w = _PEER_WINDOWS.get((id(group), rows, d, device.index))
        return w if (w is not None and w.ok) else None
    key = (id(group), rows, d, device.index)
    w = _PEER_WINDOWS.get(key)
    if w is None:
        w = _PEER_WINDOWS[key] = PeerWindow(group, world, rank, rows, d, device)
        if not w.ok:
            msg = f"vlp_b200: NVLink peer windows unavailable ({w.error}); using NCCL reduce-scatter"
            if PEER_RS_MODE == "1":
                raise RuntimeError(msg)
            if rank == 0:
                import warnings
                warnings.warn(msg)
    return w if w.ok else None


def _push_gather(i_loc, t_loc, group, world: int, rank: int):
    """(I_loc fp16, T_all fp16) with T_all assembled by NVLink peer stores from the cast kernel, or
    None when peer windows are off.  Ends with a collective every rank passes, so all shards have
    landed when it returns (stream order)."""
    if PUSH_GATHER_MODE == "0" or t_loc.dtype != torch.bfloat16 or i_loc.dtype != torch.bfloat16:
        return None
    window = peer_window(group, world, rank, t_loc.shape[0], t_loc.shape[1], t_loc.device)
    if window is None:
        if PUSH_GATHER_MODE == "1":
            raise RuntimeError("VLP_B200_PUSH_GATHER=1 but peer windows are unavailable")
        return None
    lib = _lib.load()
    t_loc = t_loc.contiguous()
    parity = window.next_gather_parity()
    rc = lib.vlpclip_cast_push_f16(t_loc.data_ptr(), t_loc.numel(), window.gather_dsts(parity),
                                   world, _stream())
    _lib.check(rc, "cast_push_f16")
    i_f16 = cast_bf16_to_f16(i_loc)        # (runs while the peers' stores are in flight)
    sharded._dist().all_reduce(torch.zeros(1, device=t_loc.device), group=group)
    return i_f16, window.gathered(parity)


def release_peer_windows() -> None:
    if _PEER_WINDOWS:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for w in _PEER_WINDOWS.values():
            if w.ok and w.local is not None:
                w.close(barrier=True)
        _PEER_WINDOWS.clear()


def _grad_scatter(x_f16, y_f16, x_stats, y_stats, scale, diag_shift: int, n_global: int,
                  w_row: float, w_col: float, want_dscale: bool, window: PeerWindow):
    """dX = grad(...) with row r stored into the window of rank r // window.rows (this rank's
    slot).  Returns (parity, dscale); finish with ``_scatter_finish`` after a collective."""
    lib = _lib.load()
    n_rows, d = x_f16.shape
    n_cols = y_f16.shape[0]
    dev = x_f16.device
    assert n_rows == window.rows * window.world and d == window.d
    sc = as_scale_tensor(scale, dev)
    ds = torch.zeros(1, dtype=torch.float32, device=dev) if want_dscale else None
    nbytes = lib.vlpclip_grad_workspace_bytes(n_rows, n_cols, d)
    ws = _ws(nbytes, dev)
    parity = window.next_parity()
    owners = window.owner_rows(parity)
    rc = lib.vlpclip_grad_scatter(
        x_f16.data_ptr(), x_f16.stride(0), y_f16.data_ptr(), y_f16.stride(0),
        x_stats[0].data_ptr(), x_stats[1].data_ptr(), x_stats[2].data_ptr(),
        y_stats[0].data_ptr(), y_stats[1].data_ptr(), y_stats[2].data_ptr(),
        n_rows, n_cols, d, sc.data_ptr(), int(diag_shift), int(n_global), float(w_row), float(w_col),
        owners, window.world, window.rows, ds.data_ptr() if want_dscale else None, ws.data_ptr(),
        nbytes, _stream())
    _lib.check(rc, "grad_scatter")
    return parity, ds


def _scatter_finish(window: PeerWindow, parity: int, out_mul, out_dtype, device):
    lib = _lib.load()
    out = torch.empty(window.rows, window.d, dtype=out_dtype, device=device)
    rc = lib.vlpclip_slot_sum(window.slots(parity), window.world, window.rows * window.d,
                              out_mul.data_ptr() if out_mul is not None else None,
                              1 if out_dtype == torch.bfloat16 else 0, out.data_ptr(), _stream())
    _lib.check(rc, "slot_sum")
    return out


class CudaOps:
    """The production ``ops`` of sharded.forward_plan / backward_plan: the sm_100a kernels."""

    @staticmethod
    def peer_window(group, world, rank, rows, d, device):
        return peer_window(group, world, rank, rows, d, device)

    @staticmethod
    def push_gather(i_loc, t_loc, group, world, rank):
        return _push_gather(i_loc, t_loc, group, world, rank)

    @staticmethod
    def grad_scatter(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale,
                     window):
        return _grad_scatter(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col,
                             want_dscale, window)

    @staticmethod
    def scatter_finish(window, parity, out_mul, out_dtype, device):
        return _scatter_finish(window, parity, out_mul, out_dtype, device)

    @staticmethod
    def lse_stats(x, y, scale, diag_shift):
        return lse_stats(x, y, scale, diag_shift)

    @staticmethod
    def lse_stats_fused(x, y, scale, diag_shift, row_ids=None, col_ids=None):
        return lse_stats_fused(x, y, scale, diag_shift, row_ids, col_ids)

    @staticmethod
    def merge_stats(part_max, part_l, diag, scale):
        return merge_stats(part_max, part_l, diag, scale)[:4]

    @staticmethod
    def loss_sums(row_loss, col_loss):
        return _loss_sums(row_loss, col_loss)

    @staticmethod
    def loss_finish(sums, n_global):
        return _loss_finish(sums, n_global)

    @staticmethod
    def grad(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale,
             out_mul=None, out_dtype=torch.float32):
        return _grad(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale,
                     out_mul, out_dtype)

    @staticmethod
    def grad_both(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col, want_dscale,
                  out_mul=None, out_dtypes=(torch.float32, torch.float32), window=None, row_ids=None,
                  col_ids=None):
        return _grad_both(x, y, x_stats, y_stats, scale, diag_shift, n_global, w_row, w_col,
                          want_dscale, out_mul, out_dtypes, window, row_ids, col_ids)

    @staticmethod
    def to_backward_operand(x_bf16):
        return cast_bf16_to_f16(x_bf16)


# ----------------------------------------------------------------------------------------------
# embedding-level fused loss
# ----------------------------------------------------------------------------------------------
class _FusedClipLoss(torch.autograd.Function):
    """loss, image_loss, text_loss = CLIP loss of (I, T, logit_scale); grads by tile recompute.

    Sharded mode (``group`` given): every rank passes its local rows; negatives are the global
    batch (gather of T -- fused into the fp16 operand cast over NVLink peer windows, else NCCL --,
    all-gather + merge of the per-column partial statistics, all-reduce of the loss sums; backward:
    reduce-scatter of dT fused into the kernel epilogue over the same windows, else NCCL;
    all-reduce of d logit_scale).
    """

    @staticmethod
    def forward(ctx, image_embeddings, text_embeddings, logit_scale, i_bf16, t_bf16, i_f16, t_f16,
                group, grad_scale, caption_ids=None):
        ctx.set_materialize_grads(False)
        _require_cuda(image_embeddings, "image_embeddings")
        _require_cuda(text_embeddings, "text_embeddings")
        if image_embeddings.shape != text_embeddings.shape or image_embeddings.dim() != 2:
            raise ValueError("image/text embeddings must both be [batch, dim] with equal shapes, got "
                             f"{tuple(image_embeddings.shape)} and {tuple(text_embeddings.shape)}")
        if logit_scale.numel() != 1:
            raise ValueError("logit_scale must have exactly one element (reference shape [1])")
        n_loc, d = image_embeddings.shape
        if n_loc == 0:
            raise ValueError("empty batch")

        # reference :456-457 -- exp + clamp(max=100), evaluated ON THE DEVICE (no host sync): the
        # kernels read s through a pointer
        # (d/dl clamp(e^l, max=100) = e^l if e^l <= 100 else 0)
        scale, dscale_dls = scale_from_logit_scale(logit_scale)

        # operand copies: fp32 embeddings take one fused pass (bf16 + fp16 together, single GPU)
        needs_grad_any = any(ctx.needs_input_grad[:3])
        single = sharded.group_info(group)[0] == 1

        def _operands(e, b16, f16):
            if b16 is None:
                if e.dtype == torch.float32 and single and e.numel() % 4 == 0:
                    b16, f16_new = cast_f32_operands(e.detach().contiguous(), want_f16=needs_grad_any and f16 is None)
                    f16 = f16 if f16 is not None else f16_new
                else:
                    b16 = e.detach().to(torch.bfloat16).contiguous()
            return b16, f16
        i_bf16, i_f16 = _operands(image_embeddings, i_bf16, i_f16)
        t_bf16, t_f16 = _operands(text_embeddings, t_bf16, t_f16)

        # fp16 operand copies for the backward: produced on a side stream so that the two small
        # casts run underneath the forward sweep instead of in front of the backward
        ctx.cast_event = None
        needs_grad = any(ctx.needs_input_grad[:3])
        if needs_grad and sharded.group_info(group)[0] == 1 and (i_f16 is None or t_f16 is None):
            main = torch.cuda.current_stream()
            side = sharded._side_stream(i_bf16.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if i_f16 is None:
                    i_f16 = cast_bf16_to_f16(i_bf16)
                if t_f16 is None:
                    t_f16 = cast_bf16_to_f16(t_bf16)
                ctx.cast_event = side.record_event()
            for t_ in (i_f16, t_f16, i_bf16, t_bf16):
                t_.record_stream(side)
                t_.record_stream(main)

        if caption_ids is not None and EXACT_COLUMNS:
            raise NotImplementedError("caption_ids need the fused column statistics (unset VLP_B200_EXACT_COLUMNS)")
        plan = sharded.forward_plan(CudaOps, i_bf16, t_bf16, scale, group,
                                    exact_columns=EXACT_COLUMNS, ids_loc=caption_ids)
        world = plan["world"]
        ctx.ids = plan.get("ids")
        ctx.group = group
        ctx.world, ctx.rank = world, plan["rank"]
        ctx.n_loc, ctx.n_glob = n_loc, plan["n_glob"]
        ctx.scale, ctx.dscale_dls = scale, dscale_dls
        ctx.grad_scale = float(grad_scale)
        ctx.in_dtypes = (image_embeddings.dtype, text_embeddings.dtype, logit_scale.dtype)
        ctx.ls_shape = logit_scale.shape
        ctx.save_for_backward(i_bf16, plan["t_all"], *plan["r_stats"], *plan["c_stats"])
        ctx.f16 = (i_f16, t_f16 if world == 1 else None)
        ctx.tail_barrier = plan["bwd_operands"] is not None
        if plan["bwd_operands"] is not None:      # the forward already swept the fp16 copies
            ctx.f16 = plan["bwd_operands"]
        return plan["loss"], plan["image_loss"], plan["text_loss"]

    @staticmethod
    def backward(ctx, g_loss, g_il, g_tl):
        with _device_guard(ctx.saved_tensors[0]):
            return _FusedClipLoss._backward(ctx, g_loss, g_il, g_tl)

    @staticmethod
    def _backward(ctx, g_loss, g_il, g_tl):
        i_bf16, t_all_bf16, r_max, r_lg, r_q, c_max, c_lg, c_q = ctx.saved_tensors
        r_stats, c_stats = (r_max, r_lg, r_q), (c_max, c_lg, c_q)
        need_i, need_t, need_ls = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_i or need_t or need_ls) or (g_loss is None and g_il is None and g_tl is None):
            return (None,) * 10
        # direction weights: d/dS = (w_r P_row + w_c P_col - (w_r + w_c) delta) / (2N)
        mul = None
        if g_il is None and g_tl is None:
            w_r = w_c = 1.0
            mul = g_loss                      # device scalar, applied after the kernels
        else:
            gl = 0.0 if g_loss is None else float(g_loss.item())
            w_r = gl + 2.0 * (0.0 if g_il is None else float(g_il.item()))
            w_c = gl + 2.0 * (0.0 if g_tl is None else float(g_tl.item()))
            sign = 1.0
            if w_r < 0 or w_c < 0:
                if w_r > 0 or w_c > 0:
                    raise NotImplementedError("mixed-sign upstream gradients for image/text loss")
                w_r, w_c, sign = -w_r, -w_c, -1.0
            if w_r == 0.0 and w_c == 0.0:
                return (None,) * 10
            mul = torch.tensor(sign, dtype=torch.float32, device=i_bf16.device)

        i_f16, t_f16 = ctx.f16
        if ctx.cast_event is not None:
            torch.cuda.current_stream().wait_event(ctx.cast_event)
        if i_f16 is None:
            i_f16 = CudaOps.to_backward_operand(i_bf16)
        t_all_f16 = t_f16 if t_f16 is not None else CudaOps.to_backward_operand(t_all_bf16)
        world = ctx.world
        gs = ctx.grad_scale
        # grad_scale (world size under DDP averaging) applies to the ROW-SHARDED gradients only:
        # d logit_scale is all-reduced to the global total on every rank, so averaging identical
        # values over ranks already leaves the true gradient
        mul_ls = mul.detach().float().reshape(1)
        mul = (mul_ls * gs).contiguous()   # device scalar, no host sync
        kdt = lambda dt: dt if dt in (torch.float32, torch.bfloat16) else torch.float32  # noqa: E731
        d_i, d_t, ds = sharded.backward_plan(
            CudaOps, i_f16, t_all_f16, r_stats, c_stats, ctx.scale, ctx.n_loc, ctx.n_glob, ctx.rank,
            world, ctx.group, w_r, w_c, need_i, need_t, need_ls, out_mul=mul,
            out_dtypes=(kdt(ctx.in_dtypes[0]), kdt(ctx.in_dtypes[1])),
            tail_barrier=ctx.tail_barrier, single_sweep=SINGLE_SWEEP, ids=ctx.ids)
        d_ls = None
        if need_ls:
            d_ls = ds * ctx.dscale_dls                      # chain rule through exp + clamp (:456-457)
            d_ls = (d_ls * mul_ls).to(ctx.in_dtypes[2]).reshape(ctx.ls_shape)
        if need_i and d_i.dtype != ctx.in_dtypes[0]:
            d_i = d_i.to(ctx.in_dtypes[0])
        if need_t and d_t.dtype != ctx.in_dtypes[1]:
            d_t = d_t.to(ctx.in_dtypes[1])
        return d_i, d_t, d_ls, None, None, None, None, None, None, None



# ----------------------------------------------------------------------------------------------
# CUDA-graph variant: the forward half and the backward half of the head captured as two graphs
# ----------------------------------------------------------------------------------------------
# A sharded step at 8 GPUs is ~40 short launches + a handful of collectives for ~0.6 ms of tensor work:
# the host cannot enqueue that fast.  With the temperature on the device nothing in the step needs the
# host, so each half (kernels, NCCL collectives, NVLink peer stores) is captured once per
# (shape, dtypes, group) and replayed; the upstream gradient is a static device scalar that the
# backward half folds into its kernel epilogues.
_GRAPH_MODE = os.environ.get("VLP_B200_CUDA_GRAPH", "auto")   # "auto" | "1" | "0"
_GRAPHS = {}
GRAPH_REPLAYED_LAUNCHES = 0    # kernels of this library launched through graph replays


def _kernel_dtype(dt):
    return dt if dt in (torch.float32, torch.bfloat16) else torch.float32


def _graph_forward(i_bf16, t_bf16, logit_scale, group, needs):
    """Forward half of a graphed step: the three losses plus everything the backward half reads."""
    scale, dscale_dls = scale_from_logit_scale(logit_scale)
    plan = sharded.forward_plan(CudaOps, i_bf16, t_bf16, scale, group, exact_columns=EXACT_COLUMNS)
    out = {"losses": plan["losses"], "plan": plan, "scale": scale, "dscale_dls": dscale_dls}
    if plan["bwd_operands"] is not None:      # the forward already swept the fp16 copies
        out["i_f16"], out["t_all_f16"] = plan["bwd_operands"]
    elif any(needs):
        out["i_f16"] = CudaOps.to_backward_operand(i_bf16)
        out["t_all_f16"] = CudaOps.to_backward_operand(plan["t_all"])
    return out


def _graph_backward(fo, g, g_ls, group, needs, in_dtypes, ls_shape):
    """Backward half: final gradients (upstream gradient ``g`` = g_loss * grad_scale folded into the
    kernel epilogues; ``g_ls`` = g_loss alone multiplies d logit_scale, see _FusedClipLoss.backward)."""
    plan = fo["plan"]
    need_i, need_t, need_ls = needs
    d_i, d_t, ds = sharded.backward_plan(
        CudaOps, fo["i_f16"], fo["t_all_f16"], plan["r_stats"], plan["c_stats"], fo["scale"],
        plan["n_loc"], plan["n_glob"], plan["rank"], plan["world"], group, 1.0, 1.0, need_i, need_t,
        need_ls, out_mul=g, out_dtypes=(_kernel_dtype(in_dtypes[0]), _kernel_dtype(in_dtypes[1])),
        tail_barrier=plan["bwd_operands"] is not None, single_sweep=SINGLE_SWEEP)
    d_ls = None
    if need_ls:
        d_ls = (ds * fo["dscale_dls"] * g_ls).to(in_dtypes[2]).reshape(ls_shape)
    if d_i is not None and d_i.dtype != in_dtypes[0]:
        d_i = d_i.to(in_dtypes[0])
    if d_t is not None and d_t.dtype != in_dtypes[1]:
        d_t = d_t.to(in_dtypes[1])
    return d_i, d_t, d_ls


class _GraphEntry:
    """Static buffers + the two captured graphs (forward half / backward half) of one step shape.

    Call 1 runs both halves eagerly (warm-up: workspaces, peer windows, NCCL channels); call 2
    captures each half the first time it runs; later calls replay."""

    def __init__(self, n_loc, d, device, in_dtypes, ls_shape, group, needs):
        self.i = torch.zeros(n_loc, d, dtype=torch.bfloat16, device=device)
        self.t = torch.zeros(n_loc, d, dtype=torch.bfloat16, device=device)
        self.ls = torch.zeros(1, dtype=in_dtypes[2] if in_dtypes[2] == torch.float64
                              else torch.float32, device=device)
        self.g = torch.ones(1, dtype=torch.float32, device=device)   # upstream gradient * grad_scale
        self.g_ls = torch.ones(1, dtype=torch.float32, device=device)   # upstream gradient alone
        self.group, self.needs = group, needs
        self.in_dtypes, self.ls_shape = in_dtypes, ls_shape
        self.fwd_graph = self.bwd_graph = None
        self.fwd_out = self.bwd_out = None
        self.launches = [0, 0]   # kernels of this library per replay (forward, backward)
        self.calls = 0
        self.generation = 0      # bumped by every forward: guards the shared static buffers

    def _capture(self, fn, pool=None):
        lib = _lib.load()
        before = lib.vlpclip_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            out = fn()
        return g, out, int(lib.vlpclip_launch_count() - before)

    def forward(self):
        global GRAPH_REPLAYED_LAUNCHES
        self.calls += 1
        run = lambda: _graph_forward(self.i, self.t, self.ls, self.group, self.needs)  # noqa: E731
        if self.fwd_graph is None and self.calls >= 2:
            self.fwd_graph, self.fwd_out, self.launches[0] = self._capture(run)
        if self.fwd_graph is not None:
            self.fwd_graph.replay()
            GRAPH_REPLAYED_LAUNCHES += self.launches[0]
        else:
            self.fwd_out = run()
        return self.fwd_out["losses"]

    def backward(self):
        global GRAPH_REPLAYED_LAUNCHES
        run = lambda: _graph_backward(self.fwd_out, self.g, self.g_ls, self.group, self.needs,  # noqa: E731
                                      self.in_dtypes, self.ls_shape)
        if self.bwd_graph is None and self.fwd_graph is not None:
            # (its inputs are the static outputs of the captured forward half)
            self.bwd_graph, self.bwd_out, self.launches[1] = self._capture(run, self.fwd_graph.pool())
        if self.bwd_graph is not None:
            self.bwd_graph.replay()
            GRAPH_REPLAYED_LAUNCHES += self.launches[1]
        else:
            self.bwd_out = run()
        return self.bwd_out

    def drop(self):
        self.fwd_graph = self.bwd_graph = None
        self.fwd_out = self.bwd_out = None


def release_graphs() -> None:
    """Drop every captured graph (call before ``destroy_process_group``: tearing down an NCCL
    communicator while graphs that captured its collectives are alive can hang)."""
    if _GRAPHS:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for entry in _GRAPHS.values():      # autograd contexts may still reference the entry
            entry.drop()
        _GRAPHS.clear()
        import gc
        gc.collect()
    release_peer_windows()


import atexit as _atexit  # noqa: E402

_atexit.register(release_graphs)


def _use_graph(world: int, grad_enabled: bool) -> bool:
    if not grad_enabled or _GRAPH_MODE == "0":
        return False
    if _GRAPH_MODE == "1":
        return True
    return world > 1          # auto: latency-bound sharded steps


class _FusedClipLossGraphed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_embeddings, text_embeddings, logit_scale, group, grad_scale):
        ctx.set_materialize_grads(False)
        n_loc, d = image_embeddings.shape
        dev = image_embeddings.device
        needs = (image_embeddings.requires_grad, text_embeddings.requires_grad,
                 logit_scale.requires_grad)
        in_dtypes = (image_embeddings.dtype, text_embeddings.dtype, logit_scale.dtype)
        key = (n_loc, d, dev.index, in_dtypes, tuple(logit_scale.shape), id(group), needs)
        entry = _GRAPHS.get(key)
        if entry is None:
            entry = _GRAPHS[key] = _GraphEntry(n_loc, d, dev, in_dtypes, logit_scale.shape, group,
                                               needs)
        entry.i.copy_(image_embeddings.detach())
        entry.t.copy_(text_embeddings.detach())
        entry.ls.copy_(logit_scale.detach().reshape(1))
        # the static loss buffer is overwritten by the next replay: hand out a copy
        losses = entry.forward().clone()
        entry.generation += 1
        ctx.entry, ctx.generation = entry, entry.generation
        ctx.grad_scale = float(grad_scale)
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_loss, g_il, g_tl):
        if g_il is not None or g_tl is not None:
            raise NotImplementedError(
                "back-propagating image_loss / text_loss separately needs the eager path "
                "(VLP_B200_CUDA_GRAPH=0): the graph computes the gradients of `loss`")
        if g_loss is None:
            return (None,) * 5
        entry = ctx.entry
        if entry.generation != ctx.generation or entry.fwd_out is None:
            raise RuntimeError(
                "fused CLIP loss (CUDA-graph mode): another forward of the same shape ran before this "
                "backward and overwrote the saved statistics; set VLP_B200_CUDA_GRAPH=0 for such "
                "schedules")
        entry.g_ls.copy_(g_loss.detach().reshape(1).to(torch.float32))
        torch.mul(entry.g_ls, ctx.grad_scale, out=entry.g)
        # gradients live in static buffers that the next backward of this shape overwrites
        d_i, d_t, d_ls = entry.backward()
        return (d_i if ctx.needs_input_grad[0] else None, d_t if ctx.needs_input_grad[1] else None,
                d_ls if ctx.needs_input_grad[2] else None, None, None)


def fused_clip_loss_from_embeddings(image_embeddings: torch.Tensor, text_embeddings: torch.Tensor,
                                    logit_scale: torch.Tensor, *, group=None,
                                    grad_scale: float = 1.0, caption_ids: Optional[torch.Tensor] = None,
                                    _operands=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Symmetric InfoNCE loss of already L2-normalised embeddings.

    Equivalent to the reference's ``logits = (I @ T.T) * clamp(exp(logit_scale), max=100)``
    followed by ``_compute_loss(logits)`` (VisionLanguageModule.py:456-459, 533-552), returning
    ``(loss, image_loss, text_loss)``.  The embeddings are consumed as bf16 (values that are
    bf16-representable are used exactly); accumulation is fp32.  ``group``: torch.distributed
    process group for the sharded global-batch variant (each rank passes its local rows);
    ``grad_scale`` multiplies the embedding gradients (use ``world_size`` under DDP gradient averaging).
    ``caption_ids`` (int tensor [batch], ids >= 0, globally consistent across ranks): duplicate-caption
    mask -- pairs (i, j != i) with the same caption id are excluded from both cross-entropies instead of
    being treated as negatives (the mask of the reference's ``_get_mask``, lines 506-530).
    """
    with _device_guard(image_embeddings, text_embeddings, logit_scale):
        return _fused_clip_loss_from_embeddings(image_embeddings, text_embeddings, logit_scale, group,
                                                grad_scale, _operands, caption_ids)


def _fused_clip_loss_from_embeddings(image_embeddings, text_embeddings, logit_scale, group, grad_scale,
                                     _operands, caption_ids=None):
    i_bf16 = t_bf16 = i_f16 = t_f16 = None
    if _operands is not None:
        i_bf16, t_bf16, i_f16, t_f16 = _operands
    world = sharded.group_info(group)[0]
    if world > 1 and (PEER_RS_MODE == "0" or world > MAX_PEER_RANKS):
        _reserve_sms_for_collectives()      # NCCL reduce-scatter overlaps with the dI kernel
    needs_grad = torch.is_grad_enabled() and (image_embeddings.requires_grad or
                                              text_embeddings.requires_grad or
                                              logit_scale.requires_grad)
    if caption_ids is not None:
        if caption_ids.numel() != image_embeddings.shape[0]:
            raise ValueError(f"caption_ids: expected {image_embeddings.shape[0]} ids, got {caption_ids.numel()}")
        caption_ids = caption_ids.detach().to(device=image_embeddings.device, dtype=torch.int32).contiguous()
    if caption_ids is None and _use_graph(world, needs_grad) and image_embeddings.is_cuda \
            and image_embeddings.dim() == 2 \
            and image_embeddings.shape == text_embeddings.shape and logit_scale.numel() == 1:
        return _FusedClipLossGraphed.apply(image_embeddings, text_embeddings, logit_scale, group,
                                           grad_scale)
    return _FusedClipLoss.apply(image_embeddings, text_embeddings, logit_scale, i_bf16, t_bf16,
                                i_f16, t_f16, group, grad_scale, caption_ids)


# ----------------------------------------------------------------------------------------------
# prologue: projection + L2 normalise
# ----------------------------------------------------------------------------------------------
def _gemm_tf32(a, b, m, n, k, trans_a, trans_b):
    lib = _lib.load()
    c = torch.empty(m, n, dtype=torch.float32, device=a.device)
    nbytes = lib.vlpclip_gemm_workspace_bytes(m, n, k)
    ws = _ws(nbytes, a.device)
    rc = lib.vlpclip_gemm_tf32(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, int(trans_a),
                               int(trans_b), ws.data_ptr(), nbytes, _stream())
    _lib.check(rc, "gemm_tf32")
    return c


class _ProjectNormalize(torch.autograd.Function):
    """emb = F.normalize(features @ W)  (reference :448-449, :452-453)."""

    @staticmethod
    def forward(ctx, features, weight, out_bf16=None):
        _require_cuda(features, "features")
        _require_cuda(weight, "projection")
        if features.dim() != 2 or weight.dim() != 2 or features.shape[1] != weight.shape[0]:
            raise ValueError(f"shape mismatch: features {tuple(features.shape)} @ projection "
                             f"{tuple(weight.shape)}")
        f32 = features.detach().float().contiguous()
        w32 = weight.detach().float().contiguous()
        n, f = f32.shape
        d = w32.shape[1]
        lib = _lib.load()
        dev = f32.device
        emb = torch.empty(n, d, dtype=torch.float32, device=dev)
        if out_bf16 is None:
            emb_bf16 = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
        else:   # caller-provided destination of the bf16 operand copy (rows of the epoch cache)
            if (out_bf16.dtype != torch.bfloat16 or tuple(out_bf16.shape) != (n, d) or not out_bf16.is_contiguous()
                    or out_bf16.device != dev or out_bf16.data_ptr() % 16 != 0):
                raise ValueError("out_bf16 must be a contiguous, 16-byte aligned bf16 [batch, dim] tensor on the "
                                 "features' device")
            emb_bf16 = out_bf16
        emb_f16 = torch.empty(n, d, dtype=torch.float16, device=dev)
        inv_norm = torch.empty(n, dtype=torch.float32, device=dev)
        nbytes = lib.vlpclip_project_workspace_bytes(n, f, d)
        ws = _ws(nbytes, dev)
        rc = lib.vlpclip_project_normalize_fwd(f32.data_ptr(), w32.data_ptr(), n, f, d,
                                               emb.data_ptr(), emb_bf16.data_ptr(),
                                               emb_f16.data_ptr(), inv_norm.data_ptr(),
                                               ws.data_ptr(), nbytes, _stream())
        _lib.check(rc, "project_normalize_fwd")
        ctx.save_for_backward(f32, w32, emb, inv_norm)
        ctx.in_dtypes = (features.dtype, weight.dtype)
        ctx.mark_non_differentiable(emb_bf16, emb_f16)
        ctx.set_materialize_grads(False)   # no 2 x 32 MB zero "gradients" for the operand copies
        return emb, emb_bf16, emb_f16

    @staticmethod
    def backward(ctx, d_emb, _g1, _g2):
        with _device_guard(ctx.saved_tensors[0]):
            return _ProjectNormalize._backward(ctx, d_emb)

    @staticmethod
    def _backward(ctx, d_emb):
        f32, w32, emb, inv_norm = ctx.saved_tensors
        if d_emb is None:
            return None, None, None
        n, f = f32.shape
        d = w32.shape[1]
        lib = _lib.load()
        d_emb = d_emb.float().contiguous()
        du = torch.empty_like(emb)
        _lib.check(lib.vlpclip_normalize_bwd(emb.data_ptr(), d_emb.data_ptr(), inv_norm.data_ptr(),
                                             n, d, du.data_ptr(), _stream()), "normalize_bwd")
        d_feat = d_w = None
        if ctx.needs_input_grad[0]:
            d_feat = _gemm_tf32(du, w32, n, f, d, 0, 1).to(ctx.in_dtypes[0])      # du @ W^T
        if ctx.needs_input_grad[1]:
            d_w = _gemm_tf32(f32, du, f, d, n, 1, 0).to(ctx.in_dtypes[1])          # feat^T @ du
        return d_feat, d_w, None


def project_normalize(features: torch.Tensor, projection: torch.Tensor, out_bf16: Optional[torch.Tensor] = None):
    """(emb_fp32, emb_bf16, emb_f16) = normalize(features @ projection).  ``out_bf16``: optional
    destination of the bf16 copy (e.g. rows of ``cache.EpochEmbeddingCache``), written by the kernel."""
    with _device_guard(features, projection):
        return _ProjectNormalize.apply(features, projection, out_bf16)


def fused_clip_loss(image_features: torch.Tensor, text_features: torch.Tensor,
                    image_projection: torch.Tensor, text_projection: torch.Tensor,
                    logit_scale: torch.Tensor, *, group=None, grad_scale: float = 1.0,
                    caption_ids: Optional[torch.Tensor] = None):
    """Full head: returns (loss, image_loss, text_loss, image_embeddings, text_embeddings).

    Same results as the reference ``forward`` (:448-459) + ``_compute_loss`` (:533-552) on the
    encoder outputs, with the loss evaluated on the bf16-rounded embeddings.
    """
    i_emb, i_bf16, i_f16 = project_normalize(image_features, image_projection)
    t_emb, t_bf16, t_f16 = project_normalize(text_features, text_projection)
    loss, il, tl = fused_clip_loss_from_embeddings(
        i_emb, t_emb, logit_scale, group=group, grad_scale=grad_scale, caption_ids=caption_ids,
        _operands=(i_bf16, t_bf16, i_f16, t_f16))
    return loss, il, tl, i_emb, t_emb


def clip_lse_stats(image_embeddings: torch.Tensor, text_embeddings: torch.Tensor,
                   logit_scale_value: float):
    """Forward-only helper: (row_lse, col_lse, diag_logit) in natural-log units (no autograd)."""
    with _device_guard(image_embeddings, text_embeddings):
        return _clip_lse_stats(image_embeddings, text_embeddings, logit_scale_value)


def _clip_lse_stats(image_embeddings, text_embeddings, logit_scale_value):
    s = min(math.exp(float(logit_scale_value)), LOGIT_SCALE_MAX)
    s = as_scale_tensor(s, image_embeddings.device)
    ib = image_embeddings.detach().to(torch.bfloat16).contiguous()
    tb = text_embeddings.detach().to(torch.bfloat16).contiguous()
    rm, rl, rdiag = lse_stats(ib, tb, s, 0)
    cm, cl, cdiag = lse_stats(tb, ib, s, 0)
    row_lse = merge_stats(rm, rl, rdiag, s, want_lse=True)[4]
    col_lse = merge_stats(cm, cl, cdiag, s, want_lse=True)[4]
    return row_lse, col_lse, rdiag * s
