"""CPU oracle for the CLIP-style contrastive head -- TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the hot path of the reference
(`/root/reference/src/models/pretrain/VisionLanguageModule.py`):

* ``forward``        lines 441-461  (projection, L2-normalise, clamp(exp(logit_scale)), N x N logits)
* ``_compute_loss``  lines 532-554  (row CE + column CE, averaged)
* parameters         lines 102-111  (projection init N(0, F^-0.5), fp64 ``logit_scale`` of shape [1])
* retrieval metrics  lines 364-439  (precision@k / recall@k; the "next" row f1 of SURVEY.md section 8)

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product package never does: it fails loudly when the CUDA
extension is missing instead of falling back to this code.

Pinning.  The reference ships NO tests, golden vectors or known answers for this path
(SURVEY.md section 8c), and the module itself cannot be imported in this image (lightning, timm, monai,
torchmetrics, hydra are absent).  The oracle is therefore pinned in two ways:

1. ``tests/golden/make_golden.py`` extracts the *source text* of the reference's ``forward`` and
   ``_compute_loss`` methods with ``ast`` from ``/root/reference`` and executes that exact code
   (bound to a stub ``self``) to produce ``tests/golden/*.npz`` -- i.e. outputs of the reference
   itself, run in this container.  ``tests/test_oracle.py`` checks this restatement against those
   fixtures (they travel to the GPU box; ``/root/reference`` does not).
2. closed-form gradients are checked against autograd in fp64, plus analytic
   known answers (identical rows => loss = ln N; orthonormal I = T => loss = ln(e^s + N - 1) - s;
   clamp edge => d logit_scale = 0).

Arithmetic lives in PyTorch itself (reference pins torch 2.7.0, this image has 2.11): ``@``,
``F.normalize``, ``exp``/``clamp``, ``F.cross_entropy`` and autograd.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LOGIT_SCALE_INIT = float(np.log(1 / 0.07))  # VisionLanguageModule.py:111
LOGIT_SCALE_MAX = 100.0                     # VisionLanguageModule.py:457


# --------------------------------------------------------------------------------------
# literal restatement (torch ops, same order as the reference)
# --------------------------------------------------------------------------------------
def reference_forward(image_features: torch.Tensor, text_features: torch.Tensor,
                      image_projection: torch.Tensor, text_projection: torch.Tensor,
                      logit_scale: torch.Tensor):
    """VisionLanguageModule.forward, lines 447-461 (after the encoders)."""
    image_embeddings = image_features @ image_projection            # :448
    text_embeddings = text_features @ text_projection               # :449
    image_embeddings = F.normalize(image_embeddings)                # :452
    text_embeddings = F.normalize(text_embeddings)                  # :453
    scale = logit_scale.exp()                                       # :456
    scale = torch.clamp(scale, max=LOGIT_SCALE_MAX)                 # :457
    logits = (image_embeddings @ text_embeddings.T) * scale         # :459
    return logits, image_embeddings, text_embeddings                # :461


def reference_compute_loss(logits: torch.Tensor, deduplicate: bool = False, masked: bool = False,
                           captions=None):
    """VisionLanguageModule._compute_loss, lines 532-554 (flag behaviour kept: both raise)."""
    labels = torch.arange(len(logits), device=logits.device)        # :533
    if deduplicate:                                                 # :535-538
        raise DeprecationWarning(
            "Deduplication loss was made obsolete by generating diverse captions and the custom batch sampler")
    if masked:                                                      # :541-544
        raise DeprecationWarning(
            "Masked loss was made obsolete by generating diverse captions and the custom batch sampler")
    image_loss = F.cross_entropy(logits, labels, reduction="mean")  # :550
    text_loss = F.cross_entropy(logits.T, labels, reduction="mean")  # :551
    loss = (image_loss + text_loss) / 2                             # :552
    return loss, image_loss, text_loss                              # :554


def loss_from_embeddings(image_embeddings: torch.Tensor, text_embeddings: torch.Tensor,
                         logit_scale: torch.Tensor):
    """Embedding-level entry: lines 456-459 + 533-552 on already normalised embeddings."""
    scale = torch.clamp(logit_scale.exp(), max=LOGIT_SCALE_MAX)
    logits = (image_embeddings @ text_embeddings.T) * scale
    return reference_compute_loss(logits)


def loss_and_grads_from_embeddings(I: torch.Tensor, T: torch.Tensor, logit_scale: torch.Tensor,
                                   dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Autograd of the literal restatement w.r.t. the embeddings and logit_scale."""
    I = I.detach().to(dtype).clone().requires_grad_(True)
    T = T.detach().to(dtype).clone().requires_grad_(True)
    ls = logit_scale.detach().to(dtype).clone().requires_grad_(True)
    loss, il, tl = loss_from_embeddings(I, T, ls)
    loss.backward()
    return {"loss": loss.detach(), "image_loss": il.detach(), "text_loss": tl.detach(),
            "dI": I.grad, "dT": T.grad, "dlogit_scale": ls.grad}


def head_loss_and_grads(image_features, text_features, image_projection, text_projection,
                        logit_scale, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Autograd of the full head (projection -> normalise -> logits -> symmetric CE)."""
    f_i = image_features.detach().to(dtype).clone().requires_grad_(True)
    f_t = text_features.detach().to(dtype).clone().requires_grad_(True)
    w_i = image_projection.detach().to(dtype).clone().requires_grad_(True)
    w_t = text_projection.detach().to(dtype).clone().requires_grad_(True)
    ls = logit_scale.detach().to(dtype).clone().requires_grad_(True)
    logits, I, T = reference_forward(f_i, f_t, w_i, w_t, ls)
    loss, il, tl = reference_compute_loss(logits)
    loss.backward()
    return {"loss": loss.detach(), "image_loss": il.detach(), "text_loss": tl.detach(),
            "image_embeddings": I.detach(), "text_embeddings": T.detach(),
            "d_image_features": f_i.grad, "d_text_features": f_t.grad,
            "d_image_projection": w_i.grad, "d_text_projection": w_t.grad,
            "dlogit_scale": ls.grad}


# --------------------------------------------------------------------------------------
# closed form (what the CUDA kernels compute), fp64 numpy
# --------------------------------------------------------------------------------------
def closed_form(I: np.ndarray, T: np.ndarray, logit_scale: float) -> Dict[str, np.ndarray]:
    """Loss and gradients in closed form, float64.

    G = dloss/dS = (P_row + P_col - 2 Id) / (2N);  dI = s G T;  dT = s G^T I;
    ds = sum(G * I T^T);  d logit_scale = ds * e^l * [e^l <= 100].
    """
    I = np.asarray(I, dtype=np.float64)
    T = np.asarray(T, dtype=np.float64)
    n = I.shape[0]
    e = math.exp(float(logit_scale))
    s = min(e, LOGIT_SCALE_MAX)
    C = I @ T.T
    S = s * C
    row_max = S.max(axis=1, keepdims=True)
    row_lse = row_max[:, 0] + np.log(np.exp(S - row_max).sum(axis=1))
    col_max = S.max(axis=0, keepdims=True)
    col_lse = col_max[0] + np.log(np.exp(S - col_max).sum(axis=0))
    diag = np.diag(S)
    image_loss = float(np.mean(row_lse - diag))
    text_loss = float(np.mean(col_lse - diag))
    P_row = np.exp(S - row_lse[:, None])
    P_col = np.exp(S - col_lse[None, :])
    G = (P_row + P_col - 2.0 * np.eye(n)) / (2.0 * n)
    dI = s * (G @ T)
    dT = s * (G.T @ I)
    ds = float((G * C).sum())
    dl = ds * e if e <= LOGIT_SCALE_MAX else 0.0
    return {"loss": 0.5 * (image_loss + text_loss), "image_loss": image_loss,
            "text_loss": text_loss, "row_lse": row_lse, "col_lse": col_lse, "diag": diag,
            "dI": dI, "dT": dT, "dscale": ds, "dlogit_scale": dl, "scale": s}


def normalize_backward(u: np.ndarray, d_emb: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """Backward of F.normalize (dim=1): du = (dE - E * rowsum(E * dE)) / max(||u||, eps)."""
    u = np.asarray(u, dtype=np.float64)
    d_emb = np.asarray(d_emb, dtype=np.float64)
    nrm = np.maximum(np.linalg.norm(u, axis=1, keepdims=True), eps)
    E = u / nrm
    return (d_emb - E * (E * d_emb).sum(axis=1, keepdims=True)) / nrm


# --------------------------------------------------------------------------------------
# sharded decomposition, collectives emulated by python sums
# --------------------------------------------------------------------------------------
def sharded_closed_form(I: np.ndarray, T: np.ndarray, logit_scale: float, world: int):
    """Row-sharded evaluation: rank r owns rows [r*b, (r+1)*b) of I and T.

    fwd: all_gather(T); local block S_r = s I_r T_all^T -> row lse local; per-column
    (max, sum-exp) partials -> all_reduce(max) then all_reduce(sum) -> col lse.
    bwd: dI_r local; dT_all partial -> reduce_scatter; ds partial -> all_reduce.
    Returns the same dict as `closed_form` (assembled from the per-rank pieces).
    """
    I = np.asarray(I, dtype=np.float64)
    T = np.asarray(T, dtype=np.float64)
    n = I.shape[0]
    assert n % world == 0
    b = n // world
    e = math.exp(float(logit_scale))
    s = min(e, LOGIT_SCALE_MAX)
    T_all = T  # all_gather
    blocks, row_lse, col_max_p, diag = [], [], [], []
    for r in range(world):
        S = s * (I[r * b:(r + 1) * b] @ T_all.T)
        blocks.append(S)
        m = S.max(axis=1, keepdims=True)
        row_lse.append(m[:, 0] + np.log(np.exp(S - m).sum(axis=1)))
        col_max_p.append(S.max(axis=0))
        diag.append(S[np.arange(b), r * b + np.arange(b)])
    col_max = np.max(np.stack(col_max_p), axis=0)                      # all_reduce(max)
    col_sum = sum(np.exp(S - col_max[None, :]).sum(axis=0) for S in blocks)  # all_reduce(sum)
    col_lse = col_max + np.log(col_sum)
    row_lse = np.concatenate(row_lse)
    diag = np.concatenate(diag)
    image_loss = float(np.mean(row_lse - diag))
    text_loss = float(np.mean(col_lse - diag))
    dI = np.zeros_like(I)
    dT = np.zeros_like(T)
    ds = 0.0
    for r in range(world):
        S = blocks[r]
        G = (np.exp(S - row_lse[r * b:(r + 1) * b, None]) + np.exp(S - col_lse[None, :]))
        G[np.arange(b), r * b + np.arange(b)] -= 2.0
        G /= (2.0 * n)
        dI[r * b:(r + 1) * b] = s * (G @ T_all)
        dT += s * (G.T @ I[r * b:(r + 1) * b])                         # reduce_scatter
        ds += float((G * S).sum()) / s                                 # all_reduce
    dl = ds * e if e <= LOGIT_SCALE_MAX else 0.0
    return {"loss": 0.5 * (image_loss + text_loss), "image_loss": image_loss,
            "text_loss": text_loss, "row_lse": row_lse, "col_lse": col_lse, "diag": diag,
            "dI": dI, "dT": dT, "dscale": ds, "dlogit_scale": dl, "scale": s}


# --------------------------------------------------------------------------------------
# duplicate-caption mask (reference lines 506-530) -- "next" row f3
# --------------------------------------------------------------------------------------
def reference_get_mask(caption_ids: torch.Tensor) -> torch.Tensor:
    """VisionLanguageModule._get_mask, lines 506-530, on caption ids instead of strings (the
    reference maps the strings to ids first, :520-521): 0.0 where two DIFFERENT samples carry the
    same caption, 1.0 elsewhere."""
    eq = caption_ids.unsqueeze(0) == caption_ids.unsqueeze(1)                    # :524
    mask = torch.ones_like(eq, dtype=torch.float)                               # :527
    mask[eq & ~torch.eye(len(caption_ids), dtype=torch.bool, device=eq.device)] = 0.0   # :528
    return mask


def masked_loss_and_grads_from_embeddings(I: torch.Tensor, T: torch.Tensor, logit_scale: torch.Tensor,
                                          caption_ids: torch.Tensor, dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """Symmetric cross-entropy of lines 456-459 + 550-552 with the masked entries of
    ``reference_get_mask`` EXCLUDED from both soft-maxes (logit -> -inf), and its autograd gradients.
    The reference's own application of the mask sits behind a DeprecationWarning (:541-544) and is
    not in the source any more; "duplicates are not negatives" is the definition the kernels
    implement (vlpclip_lse_fwd_fused_masked / vlpclip_grad_both_masked)."""
    I = I.detach().to(dtype).clone().requires_grad_(True)
    T = T.detach().to(dtype).clone().requires_grad_(True)
    ls = logit_scale.detach().to(dtype).clone().requires_grad_(True)
    scale = torch.clamp(ls.exp(), max=LOGIT_SCALE_MAX)                          # :456-457
    logits = (I @ T.T) * scale                                                  # :459
    logits = logits.masked_fill(reference_get_mask(caption_ids) == 0, float("-inf"))
    labels = torch.arange(len(logits))                                          # :533
    image_loss = F.cross_entropy(logits, labels)                                # :550
    text_loss = F.cross_entropy(logits.T, labels)                               # :551
    loss = (image_loss + text_loss) / 2                                         # :552
    loss.backward()
    return {"loss": loss.detach(), "image_loss": image_loss.detach(), "text_loss": text_loss.detach(),
            "dI": I.grad, "dT": T.grad, "dlogit_scale": ls.grad}


# --------------------------------------------------------------------------------------
# retrieval metrics (reference lines 364-439) -- "next" row f1
# --------------------------------------------------------------------------------------
def precision_at_k_on_image_embeddings(image_embeddings: torch.Tensor, labels: torch.Tensor,
                                       ks: Sequence[int]) -> Dict[int, float]:
    """VisionLanguageModule.precision_at_k_on_image_embeddings, lines 364-400."""
    assert all(k + 1 <= image_embeddings.shape[0] for k in ks), \
        "k+1 must be less than or equal to the batch size"                       # :382
    image_embeddings = F.normalize(image_embeddings)                             # :385
    similarity_matrix = image_embeddings @ image_embeddings.T                    # :386
    out = {}
    for k in ks:
        top = similarity_matrix.topk(k=k + 1, dim=1).indices[:, 1:]              # :391-393
        correct = (labels.unsqueeze(1) == labels[top]).sum(dim=1)                # :395
        out[k] = (correct.float() / k).mean().item()                             # :397-398
    return out


def recall_at_k_on_image_text_retrieval(image_embeddings: torch.Tensor,
                                        text_embeddings: torch.Tensor,
                                        ks: Sequence[int]) -> Dict[int, float]:
    """VisionLanguageModule.recall_at_k_on_image_text_retreival, lines 402-439."""
    image_embeddings = F.normalize(image_embeddings)                             # :423
    text_embeddings = F.normalize(text_embeddings)                               # :424
    similarity_matrix = image_embeddings @ text_embeddings.T                     # :425
    out = {}
    n = image_embeddings.shape[0]
    for k in ks:
        top = similarity_matrix.topk(k=k, dim=1).indices                         # :429
        targets = torch.arange(n, device=top.device)
        hit = (top == targets.unsqueeze(1)).any(dim=1)                           # :431
        out[k] = hit.sum().item() / n                                            # :433-434
    return out


def retrieval_ranks(queries: torch.Tensor, keys: torch.Tensor, chunk: int = 2048) -> torch.Tensor:
    """rank[i] = number of keys ranked before key i for query i under (similarity desc, index asc),
    i.e. the position of the paired key in a STABLE descending sort of row i of Q K^T.  With it
    recall@k of reference :425-434 is mean(rank < k): `similarity_matrix.topk(k)` contains column i
    iff fewer than k columns rank before it (exact ties excepted: topk leaves their order open).
    fp64, evaluated in row chunks; works on any device."""
    q, k = queries.double(), keys.double()
    n = q.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=q.device)
    cols = torch.arange(k.shape[0], device=q.device)
    for lo in range(0, n, chunk):
        sim = q[lo:lo + chunk] @ k.T
        rows = torch.arange(lo, min(n, lo + chunk), device=q.device)
        diag = sim[rows - lo, rows]
        before = (sim > diag[:, None]) | ((sim == diag[:, None]) & (cols[None, :] < rows[:, None]))
        out[lo:lo + chunk] = before.sum(dim=1)
    return out


def retrieval_topk(queries: torch.Tensor, keys: torch.Tensor, k: int, chunk: int = 2048) -> torch.Tensor:
    """indices of the k best keys of every query under (similarity desc, index asc) -- the first k
    entries of a stable descending sort of each row of Q K^T (reference :386-391 uses topk, which
    agrees wherever the similarities are distinct).  fp64, row chunks, any device."""
    q, kk = queries.double(), keys.double()
    out = []
    for lo in range(0, q.shape[0], chunk):
        sim = q[lo:lo + chunk] @ kk.T
        out.append(torch.sort(sim, dim=1, descending=True, stable=True).indices[:, :k])
    return torch.cat(out)


# --------------------------------------------------------------------------------------
# seeded synthetic inputs
# --------------------------------------------------------------------------------------
def make_embeddings(n: int, d: int, rho: float = 0.35, seed: int = 42,
                    round_bf16: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """a,b ~ N(0,1); b <- rho a + sqrt(1-rho^2) b; L2-normalise; round to bf16 (returned as fp32)."""
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, d, generator=g, dtype=torch.float32)
    b = torch.randn(n, d, generator=g, dtype=torch.float32)
    b = rho * a + math.sqrt(max(0.0, 1.0 - rho * rho)) * b
    I = F.normalize(a)
    T = F.normalize(b)
    if round_bf16:
        I = I.to(torch.bfloat16).to(torch.float32)
        T = T.to(torch.bfloat16).to(torch.float32)
    return I, T


def make_features(n: int, f_img: int = 512, f_txt: int = 312, d: int = 512, seed: int = 42):
    """f_img = relu(N(0,1)), f_txt ~ N(0,1); projections N(0, F^-0.5) as reference lines 102-109."""
    g = torch.Generator().manual_seed(seed)
    fi = torch.relu(torch.randn(n, f_img, generator=g))
    ft = torch.randn(n, f_txt, generator=g)
    wi = torch.randn(f_img, d, generator=g) * f_img ** -0.5
    wt = torch.randn(f_txt, d, generator=g) * f_txt ** -0.5
    return fi, ft, wi, wt


def rel_err(a, b) -> float:
    """normwise relative error ||a-b|| / ||b|| (scalar: |a-b|/|b|)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    if den == 0.0:
        return float(num)
    return float(num / den)
