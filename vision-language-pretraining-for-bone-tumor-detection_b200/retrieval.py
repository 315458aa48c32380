"""Retrieval metrics of the reference (``VisionLanguageModule.py:364-439``) without the M x M matrix.

These run once per epoch on the cached embeddings and are NOT part of the fused hot path
(SURVEY.md section 8 marks them "next", row f1): they stay on stock torch ops, but evaluate the
similarity matrix in row chunks so that at most ``chunk x M`` similarities exist at a time
(the reference materialises all M x M of them, ~5 GiB fp32 at MURA+LERA scale).
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

_CHUNK = 4096


def _topk_indices(queries: torch.Tensor, keys: torch.Tensor, k: int) -> torch.Tensor:
    out = []
    for lo in range(0, queries.shape[0], _CHUNK):
        sim = queries[lo:lo + _CHUNK] @ keys.T
        out.append(sim.topk(k=k, dim=1).indices)
    return torch.cat(out, dim=0)


def precision_at_k_on_image_embeddings(image_embeddings: torch.Tensor, labels: torch.Tensor,
                                       ks: Sequence[int]) -> Dict[int, float]:
    """Fraction of the k nearest images (cosine, self excluded) sharing the query's label."""
    assert all(k + 1 <= image_embeddings.shape[0] for k in ks), \
        "k+1 must be less than or equal to the batch size"
    emb = F.normalize(image_embeddings.detach().float())
    kmax = max(ks) + 1
    top = _topk_indices(emb, emb, kmax)          # column 0 is the query itself
    hits = labels.unsqueeze(1) == labels[top[:, 1:]]
    result = {}
    for k in ks:
        result[k] = (hits[:, :k].sum(dim=1).float() / k).mean().item()
    return result


def recall_at_k_on_image_text_retrieval(image_embeddings: torch.Tensor,
                                        text_embeddings: torch.Tensor,
                                        ks: Sequence[int]) -> Dict[int, float]:
    """Fraction of images whose paired caption is among the k most similar captions."""
    img = F.normalize(image_embeddings.detach().float())
    txt = F.normalize(text_embeddings.detach().float())
    n = img.shape[0]
    kmax = min(max(ks), n)
    top = _topk_indices(img, txt, kmax)
    target = torch.arange(n, device=top.device).unsqueeze(1)
    hit = top == target
    result = {}
    for k in ks:
        result[k] = hit[:, :min(k, kmax)].any(dim=1).sum().item() / n
    return result
