"""Retrieval metrics of the reference (``VisionLanguageModule.py:364-439``) without the M x M matrix.

Called once per epoch on ALL cached embeddings (reference ``:647-656, 688-694``); the reference
materialises the M x M similarity matrix (~5 GiB fp32 at MURA+LERA scale) and runs one ``topk`` per k.
Here both metrics come out of ONE sweep of the fused forward kernel's tile mainloop (tcgen05, the
similarity tile lives in TMEM) with a ranking epilogue (``csrc/lse_fwd.cu``, MODE_RANK / MODE_TOPK):

* recall@k: the paired caption is among the k best iff fewer than k captions rank before it, so one
  compare-and-count per similarity gives the rank of the positive pair and with it every recall@k;
* precision@k: a streaming top-16 (value, index) per row yields the neighbours' identities.

Similarities are evaluated on the bf16-rounded normalised embeddings (fp32 accumulate) -- the operand
format of the loss kernels.  Ties are ordered like a stable descending sort (ascending index); the
reference's ``topk`` leaves the order of exact ties unspecified.  There is no CPU path: tensors must
live on a B200.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

from . import _lib

MAX_K = 16      # entries of the kernel's streaming top-k (the reference's ks go up to 15, + self)


def _operand(x: torch.Tensor) -> torch.Tensor:
    """normalise (reference :385 / :423-424), round to bf16, pad the dim to a multiple of 8"""
    if not isinstance(x, torch.Tensor) or x.dim() != 2:
        raise ValueError("embeddings must be a 2-D tensor [batch, dim]")
    if not x.is_cuda:
        raise RuntimeError(f"embeddings live on {x.device}: the fused retrieval metrics only run on a "
                           "CUDA sm_100 device (no CPU fallback)")
    e = F.normalize(x.detach().float())
    pad = (-e.shape[1]) % 8
    if pad:
        e = F.pad(e, (0, pad))
    if e.shape[1] > 768:
        raise ValueError(f"embedding dim {x.shape[1]} unsupported (<= 768)")
    return e.to(torch.bfloat16).contiguous()


def retrieval_ranks(queries: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
    """rank[i] = number of keys ranked before key i for query i (int32 [n]); keys[i] is query i's pair."""
    q, k = _operand(queries), _operand(keys)
    if q.shape[1] != k.shape[1] or q.device != k.device:
        raise ValueError("queries and keys must share dim and device")
    lib = _lib.load()
    with torch.cuda.device(q.device):
        n_rows, d = q.shape
        n_cols = k.shape[0]
        rank = torch.empty(n_rows, dtype=torch.int32, device=q.device)
        nbytes = lib.vlpclip_retrieval_workspace_bytes(n_rows, n_cols, d, 0)
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=q.device)
        rc = lib.vlpclip_retrieval_ranks(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), n_rows, n_cols, d,
                                         rank.data_ptr(), ws.data_ptr(), nbytes,
                                         torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "retrieval_ranks")
    return rank


def retrieval_topk(queries: torch.Tensor, keys: torch.Tensor, k: int, return_values: bool = False):
    """idx[i, :k] = the k best keys of query i (int32 [n, k]), ties by ascending index."""
    if not 1 <= k <= MAX_K:
        raise ValueError(f"k = {k} unsupported: the kernel keeps the {MAX_K} best columns of a row")
    q, kk = _operand(queries), _operand(keys)
    if q.shape[1] != kk.shape[1] or q.device != kk.device:
        raise ValueError("queries and keys must share dim and device")
    lib = _lib.load()
    with torch.cuda.device(q.device):
        n_rows, d = q.shape
        n_cols = kk.shape[0]
        idx = torch.empty(n_rows, k, dtype=torch.int32, device=q.device)
        val = torch.empty(n_rows, k, dtype=torch.float32, device=q.device) if return_values else None
        nbytes = lib.vlpclip_retrieval_workspace_bytes(n_rows, n_cols, d, k)
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=q.device)
        rc = lib.vlpclip_retrieval_topk(q.data_ptr(), q.stride(0), kk.data_ptr(), kk.stride(0), n_rows, n_cols, d,
                                        k, idx.data_ptr(), val.data_ptr() if val is not None else None,
                                        ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "retrieval_topk")
    return (idx, val) if return_values else idx


def precision_at_k_on_image_embeddings(image_embeddings: torch.Tensor, labels: torch.Tensor,
                                       ks: Sequence[int]) -> Dict[int, float]:
    """Fraction of the k nearest images (cosine; the best match, normally the image itself, is
    dropped exactly as reference :391-393 does) sharing the query's label."""
    assert all(k + 1 <= image_embeddings.shape[0] for k in ks), \
        "k+1 must be less than or equal to the batch size"                      # reference :382
    kmax = max(ks) + 1
    top = retrieval_topk(image_embeddings, image_embeddings, kmax).long()       # column 0: the query itself
    labels = labels.to(top.device)
    hits = labels.unsqueeze(1) == labels[top[:, 1:]]
    return {k: (hits[:, :k].sum(dim=1).float() / k).mean().item() for k in ks}


def recall_at_k_on_image_text_retrieval(image_embeddings: torch.Tensor,
                                        text_embeddings: torch.Tensor,
                                        ks: Sequence[int]) -> Dict[int, float]:
    """Fraction of images whose paired caption is among the k most similar captions."""
    n = image_embeddings.shape[0]
    rank = retrieval_ranks(image_embeddings, text_embeddings)
    return {k: int((rank < k).sum().item()) / n for k in ks}
